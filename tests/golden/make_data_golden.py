#!/usr/bin/env python
"""Golden vectors of the training data path from the UNMODIFIED reference (build container only).

`dataset.py` imports librosa (absent here) at module level, so empty stand-in modules are registered for the import only;
`FolderDataset.__getitem__` (dataset.py:238-289) and `utils.uquantize/linear_quantize` then run as shipped on a dataset
object whose arrays are set directly (the `__init__` that scans wav folders is the out-of-scope part).
Writes tests/golden/data_path.npz."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SRNN_REFERENCE", "/root/reference")


def main():
    for name in ("librosa", "librosa.core"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["librosa.core"].load = lambda *a, **k: None
    sys.path.insert(0, REF)
    import dataset as ref_dataset
    import utils as ref_utils
    sys.path.pop(0)
    rs = np.random.RandomState(7)
    bs, N, seq_len, overlap, cond_len, cond_dim = 3, 2000, 160, 80, 80, 5
    t = np.arange(N)
    data = np.stack([0.6 * np.sin(0.01 * (r + 1) * t) + 0.3 * rs.randn(N) * 0.2 for r in range(bs)]).astype(np.float32)
    data = np.clip(data, -1, 1)
    data[0, 5], data[1, 7], data[2, 9] = 1.0, -1.0, 0.0                     # edge values (1.0 -> index 256 in the reference)
    n_frames = N // cond_len + 2
    cond = rs.rand(bs, n_frames, cond_dim)
    spk = rs.randint(0, 6, size=(bs, n_frames))
    out = {"data": data, "cond": cond, "global_spk": spk,
           "meta": np.array([bs, N, seq_len, overlap, cond_len, cond_dim, 256])}
    for ulaw in (True, False):
        ds = object.__new__(ref_dataset.FolderDataset)
        ds.overlap_len, ds.q_levels, ds.ulaw, ds.seq_len, ds.batch_size = overlap, 256, ulaw, seq_len, bs
        ds.quantize = ref_utils.uquantize if ulaw else ref_utils.linear_quantize
        ds.cond_len, ds.cond_dim = cond_len, cond_dim
        # the non-ulaw branch expects the stored data to be integer already (dataset.py:249-251)
        # (utils.linear_quantize only accepts 1-D tensors under current torch: `min(dim=-1)[0].expand_as` -> row by row)
        ds.data = data if ulaw else np.stack([ref_utils.linear_quantize(torch.from_numpy(r), 256).numpy() for r in data])
        ds.cond, ds.global_spk = cond, spk
        ds.length = int(np.prod(data.shape)) // seq_len
        tag = "ulaw" if ulaw else "lin"
        n_items = (N - overlap) // seq_len * bs
        for idx in range(n_items):
            d, reset, tg, c, s = ds[idx]
            out[f"{tag}/{idx}/data"], out[f"{tag}/{idx}/target"] = d.numpy(), tg.numpy()
            out[f"{tag}/{idx}/cond"], out[f"{tag}/{idx}/spk"], out[f"{tag}/{idx}/reset"] = c.numpy(), s.numpy(), np.array(reset)
        out[f"{tag}/n_items"] = np.array(n_items)
    x = torch.from_numpy(np.concatenate([np.linspace(-1, 1, 20001), rs.uniform(-1, 1, 50000)]).astype(np.float32))
    out["q/x"] = x.numpy()
    out["q/ulaw"] = ref_utils.uquantize(x, 256).numpy()
    rows = torch.from_numpy(rs.uniform(-0.8, 0.9, size=(4, 999)).astype(np.float32))
    out["q/rows"] = rows.numpy()
    out["q/linear"] = np.stack([ref_utils.linear_quantize(r, 256).numpy() for r in rows])
    path = os.path.join(HERE, "data_path.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB", "max ulaw index", int(out["q/ulaw"].max()))


if __name__ == "__main__":
    main()
