"""TEST INFRASTRUCTURE ONLY: CPU restatement of the bottle-neck conditioner chain (BASELINE.json configs[4]).

PARITY UNPINNED: the reference tree holds no source for this variant (run_sampleneck.sh:2 switches to a git branch that is
not vendored) and no test or fixture of it.  The restatement follows the thesis (doc/Barbany_report.pdf 3.2.1, Fig. 3.4):
k = 1 Conv1d layers cond_dim -> 40 -> 30 -> 20 -> ind_cond_dim, a ReLU after each; weight-norm as torch.nn.utils.weight_norm
(g * v / ||v|| per output row).  It checks the CUDA kernel's arithmetic, not the absent reference code.
"""
import numpy as np


def fold(layer):
    """state_dict entries of one layer -> (dout, din) float32 weight, (dout,) bias."""
    if "weight" in layer:
        w = np.asarray(layer["weight"], np.float32)[..., 0]
    else:
        v = np.asarray(layer["weight_v"], np.float32)[..., 0]
        g = np.asarray(layer["weight_g"], np.float32).reshape(-1, 1)
        w = g * v / np.sqrt((v.astype(np.float64) ** 2).sum(1, keepdims=True)).astype(np.float32)
    return w.astype(np.float32), np.asarray(layer["bias"], np.float32)


def chain_forward(layers, cond):
    x = np.asarray(cond, np.float32)
    for layer in layers:
        w, b = fold(layer)
        x = np.maximum(x @ w.T + b, 0).astype(np.float32)
    return x
