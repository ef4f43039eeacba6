"""Reference text formats for the FVP/CG path (whitespace-separated decimal text).

* model file: for each weight layer ``W[i]`` row-major ``[L_i][L_{i+1}]`` then ``B[i]``; finally ``LogStd[A]``
  (/root/reference/src/TRPO_FVP.c:677-696)
* data file: one row per sample ``Mean[A] Std[A] Observ[O] Action[A] Advantage``
  (/root/reference/src/TRPO_FVP.c:740-759)
* vector files (ArmTestFVP.txt / ArmTestCG.txt): ``input expected`` per row (/root/reference/src/TRPOCpuCode.c:49-51)

Numbers are written with ``repr`` (shortest round-trip), so a write -> ``fscanf("%lf")`` cycle is exact.
"""
import numpy as np


def write_model(path, theta):
    with open(path, "w") as f:
        f.write("\n".join(repr(float(x)) for x in np.asarray(theta, dtype=np.float64)))
        f.write("\n")


def write_data(path, mean, std, observ, action, advantage):
    mean = np.asarray(mean, dtype=np.float64)
    N, A = mean.shape
    std_rows = np.broadcast_to(np.asarray(std, dtype=np.float64).reshape(1, A), (N, A))
    rows = np.concatenate([mean, std_rows, np.asarray(observ, dtype=np.float64),
                           np.asarray(action, dtype=np.float64),
                           np.asarray(advantage, dtype=np.float64).reshape(N, 1)], axis=1)
    with open(path, "w") as f:
        for r in rows:
            f.write(" ".join(repr(float(x)) for x in r))
            f.write("\n")


def read_vector_pairs(path):
    a = np.loadtxt(path, dtype=np.float64)
    return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1])


def read_model(path, num_params):
    a = np.loadtxt(path, dtype=np.float64).reshape(-1)
    return np.ascontiguousarray(a[:num_params])


def read_data(path, layers, num_samples):
    O, A = layers[0], layers[-1]
    a = np.loadtxt(path, dtype=np.float64, max_rows=num_samples).reshape(num_samples, 3 * A + O + 1)
    return dict(Mean=np.ascontiguousarray(a[:, :A]), Std=np.ascontiguousarray(a[-1, A:2 * A]),
                Observ=np.ascontiguousarray(a[:, 2 * A:2 * A + O]),
                Action=np.ascontiguousarray(a[:, 2 * A + O:3 * A + O]),
                Advantage=np.ascontiguousarray(a[:, 3 * A + O]))
