"""ctypes bindings for the CHECKER libraries (test infrastructure only).

* ``liboracle.so``        -- oracle/trpo_oracle.c, the in-memory C restatement
* ``_ref/libtrpo_ref.so`` -- the unmodified reference sources (TRPO_FVP.c, TRPO_CG.c, TRPO_Update.c, TRPO_Util.c,
                             TRPO_Baseline.c, TRPO_Lightweight.c and the vendored lbfgs.c)
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


class TRPOBaselineParam(C.Structure):
    """/root/reference/src/include/TRPO.h:51-77."""
    _fields_ = [("NumLayers", C.c_size_t), ("ObservSpaceDim", C.c_size_t), ("NumEpBatch", C.c_size_t),
                ("EpLen", C.c_size_t), ("NumSamples", C.c_size_t), ("NumParams", C.c_size_t),
                ("PaddedParams", C.c_int), ("AcFunc", C.c_char_p), ("LayerSizeBase", c_size_p),
                ("WBase", C.POINTER(c_double_p)), ("BBase", C.POINTER(c_double_p)), ("LayerBase", C.POINTER(c_double_p)),
                ("GWBase", C.POINTER(c_double_p)), ("GBBase", C.POINTER(c_double_p)),
                ("GLayerBase", C.POINTER(c_double_p)),
                ("Observ", c_double_p), ("Target", c_double_p), ("Predict", c_double_p)]


class LbfgsParameter(C.Structure):
    """lbfgs_parameter_t of libLBFGS 1.10 (/root/reference/src/include/lbfgs.h:198-358), LBFGS_FLOAT = 64."""
    _fields_ = [("m", C.c_int), ("epsilon", C.c_double), ("past", C.c_int), ("delta", C.c_double),
                ("max_iterations", C.c_int), ("linesearch", C.c_int), ("max_linesearch", C.c_int),
                ("min_step", C.c_double), ("max_step", C.c_double), ("ftol", C.c_double), ("wolfe", C.c_double),
                ("gtol", C.c_double), ("xtol", C.c_double), ("orthantwise_c", C.c_double),
                ("orthantwise_start", C.c_int), ("orthantwise_end", C.c_int)]


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

    # ---- rows f-3 / f-4 -------------------------------------------------------------------------------------------
    def vf_predict(self, vf_layers, acfunc, x, observ, num_ep, ep_len):
        n = _Net(vf_layers, acfunc)
        out = np.zeros(num_ep * ep_len)
        assert self.lib.oracle_vf_predict(C.byref(n.net), C.c_size_t(num_ep), C.c_size_t(ep_len), _dp(observ), _dp(x),
                                          _dp(out)) == 0
        return out

    def vf_evaluate(self, vf_layers, acfunc, x, observ, target, num_ep, ep_len, n_padded=None):
        """(fx, g, Predict) of the baseline objective; x has num_params(vf_layers) - 1 entries (+ padding)."""
        n = _Net(vf_layers, acfunc)
        n_padded = n_padded or len(x)
        g = np.zeros(n_padded)
        pred = np.zeros(num_ep * ep_len)
        self.lib.oracle_vf_evaluate.restype = C.c_double
        fx = self.lib.oracle_vf_evaluate(C.byref(n.net), C.c_size_t(num_ep), C.c_size_t(ep_len), _dp(observ),
                                         _dp(target), _dp(x), _dp(g), C.c_int(n_padded), _dp(pred))
        return fx, g, pred

    def reward_stats(self, reward, num_ep, ep_len):
        m, s = C.c_double(), C.c_double()
        self.lib.oracle_reward_stats.restype = None
        self.lib.oracle_reward_stats(C.c_size_t(num_ep), C.c_size_t(ep_len), _dp(reward), C.byref(m), C.byref(s))
        return m.value, s.value

    def gae(self, reward, baseline, num_ep, ep_len, gamma, lam):
        """Returns (Return, standardised Advantage); ``reward`` is left untouched (the C routine works on a copy)."""
        r = np.array(reward, dtype=np.float64)
        ret = np.zeros(num_ep * ep_len)
        adv = np.zeros(num_ep * ep_len)
        assert self.lib.oracle_gae(C.c_size_t(num_ep), C.c_size_t(ep_len), C.c_double(gamma), C.c_double(lam), _dp(r),
                                   _dp(baseline), _dp(ret), _dp(adv)) == 0
        return ret, adv

    def arm_rollout(self, layers, acfunc, theta, num_ep, ep_len):
        """One batch from the lightweight arm simulator; consumes the C library's rand() stream."""
        n = _Net(layers, acfunc)
        N, O, A = num_ep * ep_len, layers[0], layers[-1]
        d = dict(Observ=np.zeros((N, O)), Mean=np.zeros((N, A)), Std=np.zeros(A), Action=np.zeros((N, A)),
                 Reward=np.zeros(N))
        assert self.lib.oracle_arm_rollout(C.byref(n.net), _dp(theta), C.c_size_t(num_ep), C.c_size_t(ep_len),
                                           _dp(d["Observ"]), _dp(d["Mean"]), _dp(d["Std"]), _dp(d["Action"]),
                                           _dp(d["Reward"])) == 0
        return d


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


    # ---- rows f-3 / f-4 -------------------------------------------------------------------------------------------
    def vf_evaluate(self, vf_layers, acfunc, x, observ, target, num_ep, ep_len, n_padded=None):
        """The reference's libLBFGS callback ``evaluate`` (TRPO_Baseline.c:29) on in-memory data."""
        K = len(vf_layers) - 1
        n_padded = n_padded or len(x)
        N = num_ep * ep_len
        keep = []

        def rows(sizes):
            arrs = [np.zeros(max(1, sz)) for sz in sizes]
            ptrs = (c_double_p * len(arrs))(*[_dp(a) for a in arrs])
            keep.extend(arrs)
            return ptrs

        bp = TRPOBaselineParam()
        ls = (C.c_size_t * len(vf_layers))(*vf_layers)
        ac = C.c_char_p(acfunc.encode())
        bp.NumLayers, bp.ObservSpaceDim, bp.NumEpBatch, bp.EpLen = len(vf_layers), vf_layers[0] - 1, num_ep, ep_len
        bp.NumSamples, bp.NumParams, bp.PaddedParams = N, num_params(vf_layers) - 1, n_padded
        bp.AcFunc, bp.LayerSizeBase = ac, C.cast(ls, c_size_p)
        bp.WBase = rows([vf_layers[i] * vf_layers[i + 1] for i in range(K)])
        bp.BBase = rows([vf_layers[i + 1] for i in range(K)])
        bp.GWBase = rows([vf_layers[i] * vf_layers[i + 1] for i in range(K)])
        bp.GBBase = rows([vf_layers[i + 1] for i in range(K)])
        bp.LayerBase = rows(list(vf_layers))
        bp.GLayerBase = rows(list(vf_layers))
        pred = np.zeros(N)
        obs = np.ascontiguousarray(observ)
        tgt = np.ascontiguousarray(target)
        bp.Observ, bp.Target, bp.Predict = _dp(obs), _dp(tgt), _dp(pred)
        g = np.zeros(n_padded)
        xx = np.ascontiguousarray(x)
        self.lib.evaluate.restype = C.c_double
        self.lib.evaluate.argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_int, C.c_double]
        fx = self.lib.evaluate(C.cast(C.byref(bp), C.c_void_p), _dp(xx), _dp(g), n_padded, 0.0)
        return fx, g, pred

    def lbfgs(self, x0, evaluate, max_iterations=25):
        """Minimise with the libLBFGS the reference vendors (src/lbfgs.c), exactly as TRPO_Lightweight.c:325-327,675
        drives it: default parameters, max_iterations = 25. ``evaluate(x, g) -> fx`` works on numpy views of the
        solver's own buffers; alternatively pass ``(fnptr, instance)`` to hand libLBFGS a native callback."""
        L = self.lib
        n = len(x0)
        L.lbfgs_malloc.restype = c_double_p
        L.lbfgs_malloc.argtypes = [C.c_int]
        L.lbfgs_free.argtypes = [c_double_p]
        L.lbfgs_parameter_init.argtypes = [C.POINTER(LbfgsParameter)]
        cb_t = C.CFUNCTYPE(C.c_double, C.c_void_p, c_double_p, c_double_p, C.c_int, C.c_double)
        L.lbfgs.restype = C.c_int
        L.lbfgs.argtypes = [C.c_int, c_double_p, c_double_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.POINTER(LbfgsParameter)]
        xbuf = L.lbfgs_malloc(n)
        xv = np.ctypeslib.as_array(xbuf, shape=(n,))
        xv[:] = x0
        prm = LbfgsParameter()
        L.lbfgs_parameter_init(C.byref(prm))
        prm.max_iterations = max_iterations
        fx = C.c_double()
        if isinstance(evaluate, tuple):
            fnptr, instance = evaluate
            cb = None
        else:
            def _cb(_inst, xp, gp, nn, _step):
                return float(evaluate(np.ctypeslib.as_array(xp, shape=(nn,)), np.ctypeslib.as_array(gp, shape=(nn,))))
            cb = cb_t(_cb)
            fnptr, instance = C.cast(cb, C.c_void_p), None
        rc = L.lbfgs(n, xbuf, C.byref(fx), fnptr, None, instance, C.byref(prm))
        out = np.array(xv)
        L.lbfgs_free(xbuf)
        return out, fx.value, rc

    def lightweight(self, model_file, baseline_file, result_prefix, layers, acfunc, damping, iters, threads=1):
        """TRPO_Lightweight (TRPO_Lightweight.c:12): the whole training loop, srand(0), 20 episodes x 150 steps; writes
        ``<result_prefix>%03d.txt`` for the last iteration (the prefix must stay under 23 characters, :1469-1473)."""
        keep = []
        p = self.param(model_file, "", layers, acfunc, 3000, damping, keep)
        p.BaselineFile = baseline_file.encode()
        p.ResultFile = result_prefix.encode()
        self.lib.TRPO_Lightweight.restype = C.c_double
        self.lib.TRPO_Lightweight.argtypes = [TRPOparam, C.c_int, C.c_size_t]
        return self.lib.TRPO_Lightweight(p, iters, threads)
