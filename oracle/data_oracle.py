"""TEST INFRASTRUCTURE ONLY (see oracle/srnn_oracle.py): CPU restatement of the reference's training data path.

  * quantisers: utils.py:9-15 (linear), utils.py:33-36,48-51,58-59 (mu-law + midrise)
  * TBPTT item slicing: dataset.py:238-289 (`FolderDataset.__getitem__`), collated over one batch like
    torch's DataLoader with shuffle=False (item index = chunk * batch_size + row)

Pinned against outputs of the UNMODIFIED reference (tests/golden/make_data_golden.py -> data_path.npz).
"""
import numpy as np

LOG_MU1 = 5.5451774444795623          # utils.py:29


def uquantize(x, q_levels=256):
    """utils.uquantize in float32, operation by operation (utils.py:33-36: ulaw; 48-51: midrise)."""
    x = np.asarray(x, dtype=np.float32)
    v = np.float32(255.0)
    y = np.sign(x) * np.log(v * np.abs(x) + np.float32(1.0)) / np.float32(LOG_MU1)
    y = y.astype(np.float32)
    t = np.float32(0.5) * (y + np.float32(1.0))
    t = t * np.float32(q_levels - 1e-6)
    return t.astype(np.int64)


def linear_quantize(x, q_levels=256):
    """utils.linear_quantize (utils.py:9-15): min/max over the last dimension."""
    s = np.array(x, dtype=np.float32, copy=True)
    s = s - s.min(axis=-1, keepdims=True)
    s = s / s.max(axis=-1, keepdims=True)
    s = s * np.float32(q_levels - 1e-2)
    s = s + np.float32(1e-2 / 2)
    return s.astype(np.int64)


def get_item(data, cond, global_spk, index, overlap_len, q_levels, ulaw, seq_len, batch_size, cond_len):
    """dataset.py:238-289 -> (data, reset, target, cond, spk)."""
    n_batch, row = divmod(index, batch_size)                               # :242
    start_data = n_batch * seq_len                                         # :245
    start_target = start_data + overlap_len
    end_target = start_target + seq_len
    q = uquantize if ulaw else linear_quantize
    if not ulaw:                                                           # :249-251 (stored data already integer)
        d = np.asarray(data[row][start_data:end_target - 1]).astype(np.int64)
        t = np.asarray(data[row][start_target:end_target]).astype(np.int64)
    else:                                                                  # :252-253
        d = q(data[row][start_data:end_target - 1], q_levels)
        t = q(data[row][start_target:end_target], q_levels)
    cond_in_seq = seq_len // cond_len                                      # :256
    reset = n_batch == 0                                                   # :258-263
    from_cond = n_batch * cond_in_seq + 1
    to_cond = from_cond + cond_in_seq
    c = cond[row][from_cond:to_cond]                                       # :274
    spk = int(np.argmax(np.bincount(np.asarray(global_spk[row][from_cond:to_cond]).astype(int))))   # :277-281
    return d, reset, t, c, np.array([spk])


def get_batch(data, cond, global_spk, n_batch, **kw):
    items = [get_item(data, cond, global_spk, n_batch * kw["batch_size"] + r, **kw) for r in range(kw["batch_size"])]
    return (np.stack([i[0] for i in items]), items[0][1], np.stack([i[2] for i in items]), np.stack([i[3] for i in items]),
            np.stack([i[4] for i in items]))
