"""Deterministic synthetic training stream shared by make_loss_curve.py (reference run) and the GPU overlay test
(SURVEY.md 8d "teacher-forced samples (ii)": mu-law-quantised sine sweeps + noise; conditioners in [0, 1])."""
import numpy as np

CONFIG = dict(frame_sizes=[20, 4], n_rnn=2, dim=64, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True,
              cond_dim=86, spk_dim=6)
B, T, STEPS, LOOKBACK = 4, 80, 1000, 80
LR = 1e-3                                               # train.py:56 default


def audio():
    """(B, LOOKBACK + STEPS*T) float64 in (-1, 1): per-row chirps whose pitch follows the conditioner."""
    n = LOOKBACK + STEPS * T
    t = np.arange(n, dtype=np.float64)
    rows = []
    rs = np.random.RandomState(20261018)
    for b in range(B):
        f = 180.0 * (b + 1) * (1.0 + 0.3 * np.sin(2 * np.pi * t / 16000.0 * (0.7 + 0.2 * b)))
        phase = 2 * np.pi * np.cumsum(f) / 16000.0
        rows.append(0.5 * np.sin(phase) + 0.2 * np.sin(2.0 * phase + b) + 0.02 * rs.randn(n))
    return np.clip(np.stack(rows), -0.999, 0.999)


def conditioners():
    """(B, STEPS + 1, cond_dim) float64 in [0, 1]."""
    f = np.arange(STEPS + 1, dtype=np.float64)[None, :, None]
    k = np.arange(CONFIG["cond_dim"], dtype=np.float64)[None, None, :]
    b = np.arange(B, dtype=np.float64)[:, None, None]
    return 0.5 + 0.5 * np.sin(0.37 * f * (1 + 0.1 * b) + 1.3 * k + b)


def speakers():
    return (np.arange(B, dtype=np.int64) % CONFIG["spk_dim"]).reshape(B, 1)


def chunk(data, cond, i):
    """dataset.py:241-266 slicing of training step i (one conditioner frame per step: T == lookback)."""
    s = i * T
    return (data[:, s: s + LOOKBACK + T - 1], data[:, s + LOOKBACK: s + LOOKBACK + T], cond[:, i + 1: i + 2])
