"""The reference's training-loop body (TRPO_Lightweight.c:349-1461) with pluggable back-ends -- test infrastructure.

One iteration = rollouts from the lightweight arm simulator (the producer; always the oracle's CPU restatement, it
consumes the C library's rand() stream like the reference) -> baseline prediction -> return / GAE advantage ->
baseline fit by the reference's vendored libLBFGS -> TRPO update. ``backend`` supplies the four compute steps so that
the same loop runs on the oracle (pinning it against the compiled reference's result file) and on the GPU library.
"""
import ctypes as C

import numpy as np

ARM_LAYERS = [15, 16, 16, 3]
ARM_VF_LAYERS = [16, 16, 16, 1]          # TRPO_Lightweight.c:36
ARM_ACFUNC = "lttl"
NUM_EP, EP_LEN = 20, 150                 # TRPO_Lightweight.c:52-55
GAMMA, LAM = 0.995, 0.98                 # TRPO_Lightweight.c:32-33
PADDED = 576                             # ceil(561 / 16) * 16, TRPO_Lightweight.c:45


class OracleBackend:
    def __init__(self, oracle):
        self.o = oracle

    def advantage(self, batch, x_base):
        base = self.o.vf_predict(ARM_VF_LAYERS, ARM_ACFUNC, x_base, batch["Observ"], NUM_EP, EP_LEN)
        return self.o.gae(batch["Reward"], base, NUM_EP, EP_LEN, GAMMA, LAM)

    def vf_callback(self, batch, target):
        def ev(x, g):
            fx, gg, _ = self.o.vf_evaluate(ARM_VF_LAYERS, ARM_ACFUNC, np.ascontiguousarray(x), batch["Observ"], target,
                                           NUM_EP, EP_LEN, n_padded=len(x))
            g[:] = gg
            return fx
        return ev

    def update(self, theta, batch, adv, damping):
        out, _ = self.o.update(ARM_LAYERS, ARM_ACFUNC, theta, batch["Std"], batch["Observ"], batch["Mean"],
                               batch["Action"], adv, damping)
        return out


def libc_rand_draws(n):
    """The next n values of the C library's rand() stream (what the reference's simulator would consume)."""
    libc = C.CDLL(None)
    libc.rand.restype = C.c_int
    return np.fromiter((libc.rand() for _ in range(n)), dtype=np.int32, count=n)


def run(backend, oracle, reference, theta0, x_base0, iters, damping=0.1, seed=0, trace=None):
    """Returns the policy parameters after ``iters`` iterations (what TRPO_Lightweight writes to its result file)."""
    C.CDLL(None).srand(seed)                                           # TRPO_Lightweight.c:66
    theta = np.array(theta0, dtype=np.float64)
    x = np.zeros(PADDED)
    x[:len(x_base0)] = x_base0
    for it in range(iters):
        if hasattr(backend, "rollout"):                                # device-side producer fed with the same rand() stream
            batch = backend.rollout(theta)
        else:
            batch = oracle.arm_rollout(ARM_LAYERS, ARM_ACFUNC, theta, NUM_EP, EP_LEN)
        ret, adv = backend.advantage(batch, x)
        x, fx, rc = reference.lbfgs(x, backend.vf_callback(batch, ret), max_iterations=25)
        theta = backend.update(theta, batch, adv, damping)
        if trace is not None:
            trace.append(dict(batch=batch, ret=ret, adv=adv, x_base=x.copy(), fx=fx, lbfgs_rc=rc, theta=theta.copy()))
    return theta
