#!/usr/bin/env python
"""Golden vectors of the Ahocoder feature interpolation from the UNMODIFIED reference `interpolate.interpolation`
(build container only).  Writes tests/golden/frontend.npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SRNN_REFERENCE", "/root/reference")


def main():
    sys.path.insert(0, REF)
    import interpolate as ref
    sys.path.pop(0)
    rs = np.random.RandomState(11)
    out, k = {}, 0
    for frac in (0.0, 0.1, 0.4, 0.8, 1.0):
        for n in (1, 2, 7, 40, 200):
            for sym, lo, hi in ((-10000000000, 4.0, 6.0), (1e3, 2e3, 6e3)):
                sig = rs.uniform(lo, hi, size=n)
                sig[rs.rand(n) < frac] = sym if sym < 0 else 0.0      # lf0 marks unvoiced with -1e10, gv with 0
                o, uv = ref.interpolation(sig.copy(), sym)
                out[f"{k}/signal"], out[f"{k}/sym"], out[f"{k}/out"], out[f"{k}/uv"] = sig, np.float64(sym), o, uv
                k += 1
    out["n"] = np.array(k)
    path = os.path.join(HERE, "frontend.npz")
    np.savez_compressed(path, **out)
    print(path, k, "cases", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
