import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def net_params(d, L=5):
    return [d[f"W{l}"] for l in range(L)], [d[f"b{l}"] for l in range(L)]


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm((a - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-30))


def flat_params(d, L=5, prefix=""):
    return np.concatenate([np.concatenate([d[f"{prefix}W{l}"].ravel(), d[f"{prefix}b{l}"].ravel()]) for l in range(L)]).astype(np.float32)


def split_flat(flat, d, L=5, prefix=""):
    out_w, out_b, off = [], [], 0
    for l in range(L):
        w, b = d[f"{prefix}W{l}"], d[f"{prefix}b{l}"]
        out_w.append(flat[off:off + w.size].reshape(w.shape)); off += w.size
        out_b.append(flat[off:off + b.size].reshape(b.shape)); off += b.size
    return out_w, out_b
