"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's generation front-end (generate.py:146-196), loop form.

`interpolation` follows interpolate.py:34-72 statement by statement (python loop); pinned against outputs of the
reference's own function in tests/golden/frontend.npz (make_frontend_golden.py).  The feature concatenation /
normalisation / look-ahead lines are inline code of `generate.main` (not importable: the module needs librosa), restated
here with their line numbers."""
import numpy as np


def interpolation(signal, unvoiced_symbol):
    tb, fb = [None, None], [None, None]                     # interpolate.py:46-47
    prev = signal[0]
    out = np.copy(signal)
    uv = np.ones(signal.shape, dtype=np.int8)
    for t in range(1, signal.shape[0]):                      # :51
        if signal[t] > unvoiced_symbol and prev <= unvoiced_symbol and tb == [None, None]:
            out[:t] = signal[t]                              # :53-56
            uv[:t] = 0
        elif signal[t] <= unvoiced_symbol and prev > unvoiced_symbol:
            tb[0], fb[0] = t - 1, prev                       # :57-59
        elif signal[t] > unvoiced_symbol and prev <= unvoiced_symbol:
            tb[1], fb[1] = t, signal[t]                      # :60-64
            for k in range(tb[0], tb[1]):                    # linear_interpolation, :34-42
                out[k] = fb[0] + (k - tb[0]) * ((fb[1] - fb[0]) / (tb[1] - tb[0]))
            uv[tb[0]:tb[1]] = 0
            tb, fb = [None, None], [None, None]              # :66-67
        prev = signal[t]
    if tb[0] is not None:                                    # :69-71
        out[tb[0]:] = fb[0]
        uv[tb[0]:] = 0
    return out, uv


def conditioner(cc, lf0, gv, speaker, min_cond, max_cond, norm_ind, look_ahead):
    f0, _ = interpolation(np.asarray(lf0, dtype=np.float64), -10000000000)       # generate.py:151-153
    f0 = f0.reshape(f0.shape[0], 1)
    fv, uv = interpolation(np.asarray(gv, dtype=np.float64), 1e3)                # :156-160
    uv = uv.reshape(fv.shape[0], 1)
    fv = fv.reshape(fv.shape[0], 1)
    cond = np.concatenate((cc, f0), axis=1)                                      # :165-167
    cond = np.concatenate((cond, fv), axis=1)
    cond = np.concatenate((cond, uv), axis=1)
    if norm_ind:                                                                 # :174-179
        cond = (cond - min_cond[speaker]) / (max_cond[speaker] - min_cond[speaker])
    else:
        cond = (cond - min_cond) / (max_cond - min_cond)
    if look_ahead:                                                               # :182-185
        delayed = np.copy(cond)
        delayed[:-1, :] = delayed[1:, :]
        cond = np.concatenate((cond, delayed), axis=1)
    return cond
