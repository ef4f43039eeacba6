"""ctypes bindings for the CHECKER libraries (test infrastructure only).

* ``liboracle.so``        -- oracle/trpo_oracle.c, the in-memory C restatement
* ``_ref/libtrpo_ref.so`` -- the unmodified reference sources (TRPO_FVP.c, TRPO_CG.c, TRPO_Update.c, TRPO_Util.c)
                             compiled by oracle/Makefile; file-based API with ``TRPOparam`` by value
                             (/root/reference/src/include/TRPO.h:6-49,88-104).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_double_p = C.POINTER(C.c_double)
c_size_p = C.POINTER(C.c_size_t)


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


class OracleNet(C.Structure):
    _fields_ = [("NumLayers", C.c_size_t), ("AcFunc", C.c_char_p), ("LayerSize", c_size_p)]


class OracleUpdateInfo(C.Structure):
    _fields_ = [("cg_iters", C.c_int), ("cg_rdotr", C.c_double * 16), ("cg_xnorm", C.c_double * 16),
                ("shs", C.c_double), ("lm", C.c_double), ("gnorm", C.c_double), ("fval", C.c_double),
                ("ls_steps", C.c_int), ("ls_accepted", C.c_int),
                ("ls_actual", C.c_double * 16), ("ls_expected", C.c_double * 16), ("ls_ratio", C.c_double * 16)]


class TRPOparam(C.Structure):
    """/root/reference/src/include/TRPO.h:6-49 (11 x 8 bytes, passed by value)."""
    _fields_ = [("ModelFile", C.c_char_p), ("BaselineFile", C.c_char_p), ("ResultFile", C.c_char_p),
                ("DataFile", C.c_char_p), ("NumLayers", C.c_size_t), ("AcFunc", C.c_char_p),
                ("LayerSize", c_size_p), ("NumSamples", C.c_size_t), ("CG_Damping", C.c_double),
                ("PaddedLayerSize", c_size_p), ("NumBlocks", c_size_p)]


def _dp(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def num_params(layers):
    return sum(layers[i] * layers[i + 1] + layers[i + 1] for i in range(len(layers) - 1)) + layers[-1]


class _Net:
    def __init__(self, layers, acfunc):
        self.layers = (C.c_size_t * len(layers))(*layers)
        self.ac = C.c_char_p(acfunc.encode() if isinstance(acfunc, str) else bytes(acfunc))
        self.net = OracleNet(len(layers), self.ac, C.cast(self.layers, c_size_p))


class Oracle:
    """The C restatement. ``fast=True`` loads the -O3 build used as the timed CPU baseline."""

    def __init__(self, fast=False):
        path = os.path.join(ORACLE_DIR, "liboracle_fast.so" if fast else "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = C.CDLL(path)
        L = self.lib
        L.oracle_num_params.restype = C.c_size_t
        for f in ("oracle_load_model", "oracle_load_data", "oracle_forward", "oracle_fvp_fast", "oracle_fvp_4pass",
                  "oracle_cg", "oracle_policy_gradient", "oracle_update"):
            getattr(L, f).restype = C.c_int

    def load_model(self, path, layers, acfunc):
        n = _Net(layers, acfunc)
        theta = np.zeros(num_params(layers))
        rc = self.lib.oracle_load_model(path.encode(), C.byref(n.net), _dp(theta))
        if rc:
            raise IOError(path)
        return theta

    def load_data(self, path, layers, acfunc, N):
        n = _Net(layers, acfunc)
        O, A = layers[0], layers[-1]
        d = dict(Mean=np.zeros((N, A)), Std=np.zeros(A), Observ=np.zeros((N, O)), Action=np.zeros((N, A)),
                 Advantage=np.zeros(N))
        rc = self.lib.oracle_load_data(path.encode(), C.byref(n.net), C.c_size_t(N), _dp(d["Mean"]), _dp(d["Std"]),
                                       _dp(d["Observ"]), _dp(d["Action"]), _dp(d["Advantage"]))
        if rc:
            raise IOError(path)
        return d

    def forward(self, layers, acfunc, theta, observ):
        n = _Net(layers, acfunc)
        N = observ.shape[0]
        mean = np.zeros((N, layers[-1]))
        assert self.lib.oracle_forward(C.byref(n.net), _dp(theta), _dp(observ), C.c_size_t(N), _dp(mean)) == 0
        return mean

    def fvp(self, layers, acfunc, theta, std, observ, damping, v, four_pass=False):
        n = _Net(layers, acfunc)
        N = observ.shape[0]
        out = np.zeros(num_params(layers))
        fn = self.lib.oracle_fvp_4pass if four_pass else self.lib.oracle_fvp_fast
        rc = fn(C.byref(n.net), _dp(theta), _dp(std), _dp(observ), C.c_size_t(N), C.c_double(damping), _dp(v), _dp(out))
        if rc:
            raise ValueError("oracle fvp failed")
        return out

    def cg(self, layers, acfunc, theta, std, observ, damping, b, max_iter=10, residual_th=1e-10):
        n = _Net(layers, acfunc)
        N = observ.shape[0]
        out = np.zeros(num_params(layers))
        rd = np.zeros(max_iter + 1)
        xn = np.zeros(max_iter + 1)
        nf = self.lib.oracle_cg(C.byref(n.net), _dp(theta), _dp(std), _dp(observ), C.c_size_t(N), C.c_double(damping),
                                _dp(b), C.c_size_t(max_iter), C.c_double(residual_th), _dp(out), _dp(rd), _dp(xn))
        if nf < 0:
            raise ValueError("oracle cg failed")
        return out, nf, rd[:nf + 1], xn[:nf + 1]

    def policy_gradient(self, layers, acfunc, theta, observ, mean, action, advantage):
        n = _Net(layers, acfunc)
        N = observ.shape[0]
        out = np.zeros(num_params(layers))
        assert self.lib.oracle_policy_gradient(C.byref(n.net), _dp(theta), _dp(observ), _dp(mean), _dp(action),
                                               _dp(advantage), C.c_size_t(N), _dp(out)) == 0
        return out

    def update(self, layers, acfunc, theta, std, observ, mean, action, advantage, damping):
        n = _Net(layers, acfunc)
        N = observ.shape[0]
        out = np.zeros(num_params(layers))
        info = OracleUpdateInfo()
        assert self.lib.oracle_update(C.byref(n.net), _dp(theta), _dp(std), _dp(observ), _dp(mean), _dp(action),
                                      _dp(advantage), C.c_size_t(N), C.c_double(damping), _dp(out), C.byref(info)) == 0
        return out, info


class Reference:
    """The unmodified reference, compiled from /root/reference by oracle/Makefile (file-based API)."""

    def __init__(self, fast=False):
        path = os.path.join(ORACLE_DIR, "_ref", "libtrpo_ref_fast.so" if fast else "libtrpo_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        L = self.lib
        L.FVP.restype = C.c_double
        L.FVP.argtypes = [TRPOparam, c_double_p, c_double_p]
        L.FVPFast.restype = C.c_double
        L.FVPFast.argtypes = [TRPOparam, c_double_p, c_double_p, C.c_size_t]
        L.CG.restype = C.c_double
        L.CG.argtypes = [TRPOparam, c_double_p, c_double_p, C.c_size_t, C.c_double, C.c_size_t]
        L.TRPO_Update.restype = C.c_double
        L.TRPO_Update.argtypes = [TRPOparam, c_double_p, C.c_size_t]

    @staticmethod
    def available():
        return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libtrpo_ref.so"))

    @staticmethod
    def param(model_file, data_file, layers, acfunc, N, damping, keep):
        ls = (C.c_size_t * len(layers))(*layers)
        ac = C.c_char_p(acfunc.encode())
        keep.extend([ls, ac])
        p = TRPOparam()
        p.ModelFile = model_file.encode()
        p.DataFile = data_file.encode()
        p.NumLayers = len(layers)
        p.AcFunc = ac
        p.LayerSize = C.cast(ls, c_size_p)
        p.NumSamples = N
        p.CG_Damping = damping
        return p

    def fvp_fast(self, model_file, data_file, layers, acfunc, N, damping, v, threads=1):
        keep = []
        p = self.param(model_file, data_file, layers, acfunc, N, damping, keep)
        out = np.zeros(num_params(layers))
        t = self.lib.FVPFast(p, _dp(out), _dp(np.ascontiguousarray(v)), threads)
        return out, t

    def fvp(self, model_file, data_file, layers, acfunc, N, damping, v):
        keep = []
        p = self.param(model_file, data_file, layers, acfunc, N, damping, keep)
        out = np.zeros(num_params(layers))
        t = self.lib.FVP(p, _dp(out), _dp(np.ascontiguousarray(v)))
        return out, t

    def cg(self, model_file, data_file, layers, acfunc, N, damping, b, max_iter=10, residual_th=1e-10, threads=1):
        keep = []
        p = self.param(model_file, data_file, layers, acfunc, N, damping, keep)
        out = np.zeros(num_params(layers))
        t = self.lib.CG(p, _dp(out), _dp(np.ascontiguousarray(b)), max_iter, residual_th, threads)
        return out, t

    def update(self, model_file, data_file, layers, acfunc, N, damping, threads=1):
        keep = []
        p = self.param(model_file, data_file, layers, acfunc, N, damping, keep)
        out = np.zeros(num_params(layers))
        t = self.lib.TRPO_Update(p, _dp(out), threads)
        return out, t
