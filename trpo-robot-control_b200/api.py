"""ctypes bindings of ``libtrpo_b200.so`` (C-ABI declared in ``include/trpo_b200.h``).

Mirrors the reference's operator interface for the path:
  FVP_FPGA / CG_FPGA  (/root/reference/src/include/TRPO.h:98,101)  ->  FVP_GPU / CG_GPU
  TRPO_Update         (/root/reference/src/include/TRPO.h:104)      ->  TRPO_Update_GPU
plus the persistent context (``Context``) that stages a rollout batch once.
Fails loudly when the library is missing: there is no fallback path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
c_double_p = C.POINTER(C.c_double)
c_size_p = C.POINTER(C.c_size_t)

PATH_AUTO, PATH_GEMM_CHAIN, PATH_FUSED = 0, 1, 2
COMM_NCCL, COMM_P2P = 0, 1
PRECISION_FP64, PRECISION_FP32 = 0, 1


class TRPOparam(C.Structure):
    """Same field order as /root/reference/src/include/TRPO.h:6-49; passed by value."""
    _fields_ = [("ModelFile", C.c_char_p), ("BaselineFile", C.c_char_p), ("ResultFile", C.c_char_p),
                ("DataFile", C.c_char_p), ("NumLayers", C.c_size_t), ("AcFunc", C.c_char_p),
                ("LayerSize", c_size_p), ("NumSamples", C.c_size_t), ("CG_Damping", C.c_double),
                ("PaddedLayerSize", c_size_p), ("NumBlocks", c_size_p)]


class TrpoInfo(C.Structure):
    _fields_ = [("cg_iters", C.c_int), ("cg_rdotr", C.c_double * 34), ("cg_xnorm", C.c_double * 34),
                ("shs", C.c_double), ("lm", C.c_double), ("gnorm", C.c_double), ("fval", C.c_double),
                ("ls_steps", C.c_int), ("ls_accepted", C.c_int),
                ("ls_actual", C.c_double * 16), ("ls_expected", C.c_double * 16), ("ls_ratio", C.c_double * 16)]


def library_path(dropin=False):
    return os.path.join(PKG_DIR, "libtrpo_b200_dropin.so" if dropin else "libtrpo_b200.so")


def build_library():
    """Compile every CUDA/C source for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-s", "-C", PKG_DIR, "-j8"], check=True)


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        L = C.CDLL(path)
        L.trpo_last_error.restype = C.c_char_p
        L.trpo_num_params.restype = C.c_size_t
        L.trpo_num_params.argtypes = [c_size_p, C.c_size_t]
        L.trpo_ctx_create.restype = C.c_void_p
        L.trpo_ctx_create.argtypes = [c_size_p, C.c_char_p, C.c_size_t, C.c_int, C.c_int]
        L.trpo_ctx_destroy.argtypes = [C.c_void_p]
        L.trpo_ctx_num_params.restype = C.c_size_t
        L.trpo_ctx_num_params.argtypes = [C.c_void_p]
        L.trpo_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.trpo_ctx_get_stream.restype = C.c_void_p
        L.trpo_ctx_get_stream.argtypes = [C.c_void_p]
        L.trpo_ctx_set_path.argtypes = [C.c_void_p, C.c_int]
        L.trpo_ctx_get_path.argtypes = [C.c_void_p]
        L.trpo_ctx_set_chunk.argtypes = [C.c_void_p, C.c_size_t]
        L.trpo_ctx_get_chunk.restype = C.c_size_t
        L.trpo_ctx_get_chunk.argtypes = [C.c_void_p]
        L.trpo_ctx_get_cg_trace.argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_size_t]
        L.trpo_host_alloc_pinned.restype = C.c_void_p
        L.trpo_host_alloc_pinned.argtypes = [C.c_size_t]
        L.trpo_host_free_pinned.argtypes = [C.c_void_p]
        L.trpo_vf_failed.argtypes = [C.c_void_p]
        L.trpo_ctx_sync.argtypes = [C.c_void_p]
        L.trpo_ctx_solve_kernel_used.argtypes = [C.c_void_p]
        L.trpo_ctx_solve_timeline.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_ulonglong)]
        L.trpo_ctx_launch_count.restype = C.c_longlong
        L.trpo_ctx_launch_count.argtypes = [C.c_void_p]
        L.trpo_ctx_kernel_timing.argtypes = [C.c_void_p, C.c_int]
        L.trpo_ctx_kernel_time_ms.restype = C.c_double
        L.trpo_ctx_kernel_time_ms.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.trpo_ctx_set_model.argtypes = [C.c_void_p, c_double_p]
        L.trpo_ctx_set_batch.argtypes = [C.c_void_p, C.c_size_t] + [c_double_p] * 5
        L.trpo_ctx_set_batch_device.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, c_double_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.trpo_ctx_fvp.argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_double]
        L.trpo_ctx_cg.argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_size_t, C.c_double, C.c_double]
        L.trpo_ctx_policy_gradient.argtypes = [C.c_void_p, c_double_p]
        L.trpo_ctx_forward.argtypes = [C.c_void_p, c_double_p]
        L.trpo_ctx_update.argtypes = [C.c_void_p, c_double_p, C.c_double]
        L.trpo_ctx_get_info.argtypes = [C.c_void_p, C.POINTER(TrpoInfo)]
        L.trpo_ctx_fvp_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        L.trpo_ctx_cg_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_double]
        L.trpo_ctx_set_rollout.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t] + [c_double_p] * 5
        L.trpo_ctx_rollout_arm.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_int), C.c_ulonglong]
        L.trpo_ctx_get_rollout.argtypes = [C.c_void_p] + [c_double_p] * 4
        L.trpo_vf_create.restype = C.c_void_p
        L.trpo_vf_create.argtypes = [C.c_void_p, c_size_p, C.c_char_p, C.c_size_t]
        L.trpo_vf_destroy.argtypes = [C.c_void_p]
        L.trpo_vf_num_params.restype = C.c_size_t
        L.trpo_vf_num_params.argtypes = [C.c_void_p]
        L.trpo_vf_bind_batch.argtypes = [C.c_void_p, C.c_size_t]
        L.trpo_vf_set_target.argtypes = [C.c_void_p, c_double_p]
        L.trpo_vf_predict.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.trpo_vf_advantage.argtypes = [C.c_void_p, c_double_p, C.c_double, C.c_double, c_double_p, c_double_p]
        L.trpo_vf_evaluate.restype = C.c_double
        L.trpo_vf_evaluate.argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_int, C.c_double]
        L.trpo_batch_file_write.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t] + [c_double_p] * 5
        L.trpo_batch_file_from_text.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t]
        L.trpo_batch_file_read.argtypes = [C.c_char_p, C.c_size_t] + [c_double_p] * 5
        L.trpo_ctx_set_batch_file.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.trpo_nccl_unique_id.argtypes = [C.c_char_p]
        L.trpo_ctx_init_comm.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        L.trpo_ctx_p2p_export.argtypes = [C.c_void_p, C.c_char_p]
        L.trpo_ctx_p2p_attach.argtypes = [C.c_void_p, C.c_char_p]
        L.trpo_ctx_set_comm_mode.argtypes = [C.c_void_p, C.c_int]
        L.trpo_ctx_comm_error.argtypes = [C.c_void_p]
        for name in ("trpo_probe_fp64_peak_tflops", "trpo_probe_tf32_mma_sync_tflops"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_int]
        L.trpo_ctx_global_samples.restype = C.c_size_t
        L.trpo_ctx_global_samples.argtypes = [C.c_void_p]
        for name in ("FVP_GPU", "CG_GPU", "TRPO_Update_GPU", "TRPO_Lightweight_GPU", "TRPO_Lightweight_GPU_ex"):
            getattr(L, name).restype = C.c_double
        L.TRPO_Lightweight_GPU.argtypes = [TRPOparam, C.c_int, C.c_size_t]
        L.TRPO_Lightweight_GPU_ex.argtypes = [TRPOparam, C.c_int, C.c_size_t, C.c_size_t, C.c_double, C.c_double]
        L.FVP_GPU.argtypes = [TRPOparam, c_double_p, c_double_p]
        L.CG_GPU.argtypes = [TRPOparam, c_double_p, c_double_p, C.c_size_t, C.c_double, C.c_size_t]
        L.TRPO_Update_GPU.argtypes = [TRPOparam, c_double_p, C.c_size_t]
        _LIB = L
    return _LIB


def last_error():
    return lib().trpo_last_error().decode()


def _dp(a):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need contiguous float64"
    return a.ctypes.data_as(c_double_p)


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"libtrpo_b200: {last_error()}")


def num_params(layers):
    arr = (C.c_size_t * len(layers))(*layers)
    return lib().trpo_num_params(arr, len(layers))


def make_param(model_file, data_file, layers, acfunc, num_samples, damping, keep):
    ls = (C.c_size_t * len(layers))(*layers)
    ac = C.c_char_p(acfunc.encode())
    keep.extend([ls, ac])
    p = TRPOparam()
    p.ModelFile = model_file.encode()
    p.DataFile = data_file.encode()
    p.NumLayers = len(layers)
    p.AcFunc = ac
    p.LayerSize = C.cast(ls, c_size_p)
    p.NumSamples = num_samples
    p.CG_Damping = damping
    return p


def FVP_GPU(model_file, data_file, layers, acfunc, num_samples, damping, v):
    """File-based drop-in for FVP_FPGA (/root/reference/src/TRPO_FVP_FPGA.c:13). Returns (result, seconds)."""
    keep = []
    p = make_param(model_file, data_file, layers, acfunc, num_samples, damping, keep)
    out = np.zeros(num_params(layers))
    t = lib().FVP_GPU(p, _dp(out), _dp(np.ascontiguousarray(v, dtype=np.float64)))
    return out, t


def CG_GPU(model_file, data_file, layers, acfunc, num_samples, damping, b, max_iter=10, residual_th=1e-10, threads=1):
    """File-based drop-in for CG_FPGA (/root/reference/src/TRPO_CG_FPGA.c:13). Returns (result, seconds)."""
    keep = []
    p = make_param(model_file, data_file, layers, acfunc, num_samples, damping, keep)
    out = np.zeros(num_params(layers))
    t = lib().CG_GPU(p, _dp(out), _dp(np.ascontiguousarray(b, dtype=np.float64)), max_iter, residual_th, threads)
    return out, t


def TRPO_Update_GPU(model_file, data_file, layers, acfunc, num_samples, damping, threads=1):
    """GPU counterpart of TRPO_Update (/root/reference/src/TRPO_Update.c:10). Returns (result, seconds)."""
    keep = []
    p = make_param(model_file, data_file, layers, acfunc, num_samples, damping, keep)
    out = np.zeros(num_params(layers))
    t = lib().TRPO_Update_GPU(p, _dp(out), threads)
    return out, t


def TRPO_Lightweight_GPU(model_file, baseline_file, result_prefix, layers, acfunc, damping, iters, threads=1,
                         num_ep=None, ep_len=None, gamma=0.995, lam=0.98):
    """The reference's all-in-one training loop on the GPU; num_ep / ep_len select the _ex entry point."""
    keep = []
    p = make_param(model_file, "", layers, acfunc, 0, damping, keep)
    p.BaselineFile = baseline_file.encode()
    p.ResultFile = result_prefix.encode()
    if num_ep is None:
        return lib().TRPO_Lightweight_GPU(p, iters, threads)
    return lib().TRPO_Lightweight_GPU_ex(p, iters, num_ep, ep_len, gamma, lam)


class Context:
    """Persistent device context: model + rollout batch staged once, CG state resident on the GPU."""

    def __init__(self, layers, acfunc, device=-1, precision=0):
        self.layers = list(layers)
        self.acfunc = acfunc
        ls = (C.c_size_t * len(layers))(*layers)
        self.h = lib().trpo_ctx_create(ls, acfunc.encode(), len(layers), device, precision)
        if not self.h:
            raise RuntimeError(f"libtrpo_b200: {last_error()}")
        self.P = lib().trpo_ctx_num_params(self.h)
        self._keep = []

    def close(self):
        if self.h:
            lib().trpo_ctx_destroy(self.h)          # synchronises: no copy can still be reading the kept source
            self.h = None
        self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_ptr):
        _check(lib().trpo_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_path(self, path):
        _check(lib().trpo_ctx_set_path(self.h, path))

    def path_used(self):
        return lib().trpo_ctx_get_path(self.h)

    def set_chunk(self, chunk_samples):
        """GEMM-chain path: samples per pass over the kernel chain (0 = automatic)."""
        _check(lib().trpo_ctx_set_chunk(self.h, chunk_samples))

    def chunk_used(self):
        return lib().trpo_ctx_get_chunk(self.h)

    def cg_trace(self):
        """(rdotr, xnorm) of every iteration of the last CG -- the values TRPO_CG.c:56 prints."""
        n = self.info().cg_iters + 1
        rd, xn = np.zeros(n), np.zeros(n)
        got = lib().trpo_ctx_get_cg_trace(self.h, _dp(rd), _dp(xn), n)
        return rd[:got], xn[:got]

    def sync(self):
        _check(lib().trpo_ctx_sync(self.h))

    def solve_timeline(self, max_iters, read=False):
        """Enable (read=False) or read (read=True -> array [max_iters, 8] of ns stamps) the solve kernel's phase timeline."""
        if not read:
            _check(lib().trpo_ctx_solve_timeline(self.h, max_iters, None))
            return None
        out = np.zeros((max_iters, 8), dtype=np.uint64)
        _check(lib().trpo_ctx_solve_timeline(self.h, max_iters, out.ctypes.data_as(C.POINTER(C.c_ulonglong))))
        return out

    def solve_kernel_used(self):
        """True if the last CG ran as the single persistent cooperative kernel."""
        return bool(lib().trpo_ctx_solve_kernel_used(self.h))

    def launch_count(self):
        return lib().trpo_ctx_launch_count(self.h)

    def kernel_timing(self, enable=True):
        _check(lib().trpo_ctx_kernel_timing(self.h, int(enable)))

    def kernel_time_ms(self):
        n = C.c_int(0)
        ms = lib().trpo_ctx_kernel_time_ms(self.h, C.byref(n))
        return ms, n.value

    def set_model(self, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        assert theta.size == self.P
        _check(lib().trpo_ctx_set_model(self.h, _dp(theta)))

    def set_batch(self, observ, std, mean=None, action=None, advantage=None):
        observ = np.ascontiguousarray(observ, dtype=np.float64)
        std = np.ascontiguousarray(std, dtype=np.float64)
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (mean, action, advantage)]
        # a pinned source is copied asynchronously (the first FVP overlaps the DMA): keep it alive until the next batch
        self._keep = [observ]
        _check(lib().trpo_ctx_set_batch(self.h, observ.shape[0], _dp(observ), _dp(std), *[_dp(a) for a in arrs]))

    def set_batch_device(self, num_samples, d_observ, std, d_mean=0, d_action=0, d_advantage=0):
        std = np.ascontiguousarray(std, dtype=np.float64)
        _check(lib().trpo_ctx_set_batch_device(self.h, num_samples, C.c_void_p(d_observ), _dp(std),
                                               C.c_void_p(d_mean or None), C.c_void_p(d_action or None),
                                               C.c_void_p(d_advantage or None)))

    def set_rollout(self, num_ep, ep_len, observ, std, mean, action, reward):
        """Stage one batch of rollouts (row = ep * ep_len + step); the advantage comes from ValueFunction.advantage."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (observ, std, mean, action, reward)]
        assert arrs[0].shape[0] == num_ep * ep_len
        self._keep = [arrs[0]]
        _check(lib().trpo_ctx_set_rollout(self.h, num_ep, ep_len, *[_dp(a) for a in arrs]))

    def rollout_arm(self, num_ep, ep_len, rand_draws=None, seed=0):
        """Produce one batch of rollouts of the lightweight arm simulator on the device with the current model.
        rand_draws: int32 array of raw rand() values in the reference's order (num_ep * (3 + 6 * ep_len)), or None."""
        ptr = None
        if rand_draws is not None:
            rand_draws = np.ascontiguousarray(rand_draws, dtype=np.int32)
            assert rand_draws.size == num_ep * (3 + 2 * self.layers[-1] * ep_len)
            ptr = rand_draws.ctypes.data_as(C.POINTER(C.c_int))
        _check(lib().trpo_ctx_rollout_arm(self.h, num_ep, ep_len, ptr, seed))

    def get_rollout(self, num_samples):
        O, A = self.layers[0], self.layers[-1]
        d = dict(Observ=np.zeros((num_samples, O)), Mean=np.zeros((num_samples, A)), Action=np.zeros((num_samples, A)),
                 Reward=np.zeros(num_samples))
        _check(lib().trpo_ctx_get_rollout(self.h, _dp(d["Observ"]), _dp(d["Mean"]), _dp(d["Action"]), _dp(d["Reward"])))
        return d

    def set_batch_file(self, path, num_samples=0):
        _check(lib().trpo_ctx_set_batch_file(self.h, path.encode(), num_samples))

    def fvp(self, v, damping):
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.zeros(self.P)
        _check(lib().trpo_ctx_fvp(self.h, _dp(v), _dp(out), damping))
        return out

    def cg(self, b, max_iter=10, residual_th=1e-10, damping=0.1):
        b = np.ascontiguousarray(b, dtype=np.float64)
        out = np.zeros(self.P)
        _check(lib().trpo_ctx_cg(self.h, _dp(b), _dp(out), max_iter, residual_th, damping))
        return out, self.info()

    def forward(self, num_samples):
        out = np.zeros((num_samples, self.layers[-1]))
        _check(lib().trpo_ctx_forward(self.h, _dp(out)))
        return out

    def policy_gradient(self):
        out = np.zeros(self.P)
        _check(lib().trpo_ctx_policy_gradient(self.h, _dp(out)))
        return out

    def update(self, damping=0.1):
        out = np.zeros(self.P)
        _check(lib().trpo_ctx_update(self.h, _dp(out), damping))
        return out, self.info()

    def info(self):
        info = TrpoInfo()
        _check(lib().trpo_ctx_get_info(self.h, C.byref(info)))
        return info

    def fvp_device(self, d_in, d_out, damping):
        _check(lib().trpo_ctx_fvp_device(self.h, C.c_void_p(d_in), C.c_void_p(d_out), damping))

    def cg_device(self, d_b, d_out, max_iter=10, residual_th=1e-10, damping=0.1):
        _check(lib().trpo_ctx_cg_device(self.h, C.c_void_p(d_b), C.c_void_p(d_out), max_iter, residual_th, damping))

    def init_comm(self, unique_id, rank, world):
        _check(lib().trpo_ctx_init_comm(self.h, unique_id, rank, world))

    def p2p_export(self):
        buf = C.create_string_buffer(64)
        _check(lib().trpo_ctx_p2p_export(self.h, buf))
        return buf.raw

    def p2p_attach(self, handles):
        _check(lib().trpo_ctx_p2p_attach(self.h, handles))

    def set_comm_mode(self, mode):
        _check(lib().trpo_ctx_set_comm_mode(self.h, mode))

    def comm_error(self):
        return lib().trpo_ctx_comm_error(self.h)

    def global_samples(self):
        return lib().trpo_ctx_global_samples(self.h)


class ValueFunction:
    """Baseline network bound to a policy Context (TRPO_Baseline.c / TRPO_Lightweight.c:565-694 on the GPU)."""

    def __init__(self, policy, vf_layers, acfunc):
        self.policy = policy
        self.layers = list(vf_layers)
        ls = (C.c_size_t * len(vf_layers))(*vf_layers)
        self.h = lib().trpo_vf_create(policy.h, ls, acfunc.encode(), len(vf_layers))
        if not self.h:
            raise RuntimeError(f"libtrpo_b200: {last_error()}")
        self.num_params = lib().trpo_vf_num_params(self.h)

    def close(self):
        if self.h:
            lib().trpo_vf_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def bind_batch(self, ep_len):
        _check(lib().trpo_vf_bind_batch(self.h, ep_len))

    def set_target(self, target):
        target = np.ascontiguousarray(target, dtype=np.float64)
        _check(lib().trpo_vf_set_target(self.h, _dp(target)))

    def predict(self, x, num_samples):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(num_samples)
        _check(lib().trpo_vf_predict(self.h, _dp(x), _dp(out)))
        return out

    def advantage(self, x, num_samples, gamma, lam, fetch=True):
        """(Return, standardised Advantage); also installs them as the baseline target / the policy batch's Advantage.
        fetch=False leaves both on the device (nothing is copied back) and returns None."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        if not fetch:
            _check(lib().trpo_vf_advantage(self.h, _dp(x), gamma, lam, None, None))
            return None
        ret, adv = np.zeros(num_samples), np.zeros(num_samples)
        _check(lib().trpo_vf_advantage(self.h, _dp(x), gamma, lam, _dp(ret), _dp(adv)))
        return ret, adv

    def evaluate(self, x, n_padded=None):
        """(fx, g) of the baseline objective -- the libLBFGS callback called from Python."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        g = np.zeros(n_padded or x.size)
        fx = lib().trpo_vf_evaluate(self.h, _dp(x), _dp(g), g.size, 0.0)
        if lib().trpo_vf_failed(self.h):     # a failed evaluation returns +inf to L-BFGS and is recorded on the network
            raise RuntimeError(f"libtrpo_b200: {last_error()}")
        return fx, g

    def callback_pointer(self):
        """(function pointer, instance) to hand to a native libLBFGS: lbfgs(n, x, &fx, fn, NULL, instance, &param)."""
        return C.cast(lib().trpo_vf_evaluate, C.c_void_p), C.c_void_p(self.h)


def batch_file_write(path, observ, std, mean=None, action=None, advantage=None):
    observ = np.ascontiguousarray(observ, dtype=np.float64)
    std = np.ascontiguousarray(std, dtype=np.float64)
    arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (mean, action, advantage)]
    _check(lib().trpo_batch_file_write(path.encode(), observ.shape[0], observ.shape[1], std.size, _dp(observ), _dp(std),
                                       *[_dp(a) for a in arrs]))


def batch_file_from_text(text_path, bin_path, num_samples, obs_dim, act_dim):
    _check(lib().trpo_batch_file_from_text(text_path.encode(), bin_path.encode(), num_samples, obs_dim, act_dim))


def batch_file_read(path, num_samples, obs_dim, act_dim):
    d = dict(Observ=np.zeros((num_samples, obs_dim)), Std=np.zeros(act_dim), Mean=np.zeros((num_samples, act_dim)),
             Action=np.zeros((num_samples, act_dim)), Advantage=np.zeros(num_samples))
    _check(lib().trpo_batch_file_read(path.encode(), num_samples, _dp(d["Observ"]), _dp(d["Std"]), _dp(d["Mean"]),
                                      _dp(d["Action"]), _dp(d["Advantage"])))
    return d


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().trpo_nccl_unique_id(buf))
    return buf.raw
