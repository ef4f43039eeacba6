"""Synthetic model / rollout-batch generator for the shapes BASELINE.json names.

W ~ U(-g, g) with g = sqrt(6/(fan_in+fan_out)) (the ArmTest range), B ~ U(-0.1, 0.1), LogStd ~ U(-0.5, 0),
Observ ~ N(0, 1), Action = Mean + Std*N(0,1), Advantage standardised, CG right-hand side b ~ 0.01*N(0,1),
FVP direction v ~ U(0, 1). Seeds: 0x5EEDB200 + config index.
"""
import numpy as np

SEED_BASE = 0x5EEDB200

# name -> (layers, acfunc, default N)
SHAPES = {
    "arm": ([15, 16, 16, 3], "lttl", 50_000),          # armDOF_0, /root/reference/src/TRPOCpuCode.c:142-160
    "mlp64": ([17, 64, 64, 6], "lttl", 1_000_000),     # 64x64 tanh Gaussian policy
    "pendulum64": ([4, 64, 64, 1], "lttl", 1_000_000),  # /root/reference/src/include/TRPO.h:30
    "humanoid64": ([376, 64, 64, 17], "lttl", 1_000_000),  # /root/reference/src/include/TRPO.h:31
    "humanoid256": ([376, 256, 256, 17], "lttl", 1_000_000),
}


def num_params(layers):
    return sum(layers[i] * layers[i + 1] + layers[i + 1] for i in range(len(layers) - 1)) + layers[-1]


def layout(layers):
    """Offsets (w_off[i], b_off[i]) and logstd_off of the flat vector (/root/reference/src/TRPO_FVP.c:704-725)."""
    w, b, pos = [], [], 0
    for i in range(len(layers) - 1):
        w.append(pos); pos += layers[i] * layers[i + 1]
        b.append(pos); pos += layers[i + 1]
    return w, b, pos


def make_model(layers, seed):
    rng = np.random.default_rng(seed)
    theta = np.zeros(num_params(layers))
    w, b, ls = layout(layers)
    for i in range(len(layers) - 1):
        g = np.sqrt(6.0 / (layers[i] + layers[i + 1]))
        theta[w[i]:w[i] + layers[i] * layers[i + 1]] = rng.uniform(-g, g, layers[i] * layers[i + 1])
        theta[b[i]:b[i] + layers[i + 1]] = rng.uniform(-0.1, 0.1, layers[i + 1])
    theta[ls:] = rng.uniform(-0.5, 0.0, layers[-1])
    return theta


def forward_numpy(layers, acfunc, theta, observ):
    """Policy mean (plain numpy; used only to fill the Mean column of synthetic batches)."""
    w, b, _ = layout(layers)
    y = observ
    for i in range(len(layers) - 1):
        W = theta[w[i]:w[i] + layers[i] * layers[i + 1]].reshape(layers[i], layers[i + 1])
        x = y @ W + theta[b[i]:b[i] + layers[i + 1]]
        a = acfunc[i + 1]
        if a == "t":
            y = np.tanh(x)
        elif a == "o":
            y = 0.1 * x
        elif a == "s":
            y = 1.0 / (1.0 + np.exp(-x))
        else:
            y = x
    return y


def make_batch(layers, acfunc, theta, N, seed, obs_dist="normal"):
    rng = np.random.default_rng(seed + 1)
    O, A = layers[0], layers[-1]
    if obs_dist == "arm":
        observ = rng.uniform(-0.17, 0.19, (N, O))
    else:
        observ = rng.standard_normal((N, O))
    std = np.exp(theta[-A:])
    mean = forward_numpy(layers, acfunc, theta, observ)
    action = mean + std * rng.standard_normal((N, A))
    adv = rng.standard_normal(N)
    adv = (adv - adv.mean()) / adv.std()
    return dict(Mean=np.ascontiguousarray(mean), Std=std, Observ=np.ascontiguousarray(observ),
                Action=np.ascontiguousarray(action), Advantage=np.ascontiguousarray(adv))


def make_vectors(layers, seed):
    rng = np.random.default_rng(seed + 2)
    P = num_params(layers)
    return dict(v=rng.uniform(0.0, 1.0, P), b=0.01 * rng.standard_normal(P))
