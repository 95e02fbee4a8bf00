"""Training data path on the device: batched mirror of ``FolderDataset.__getitem__`` (dataset.py:238-289) + the DataLoader
collation of one batch (train.py builds ``DataLoader(dataset, batch_size, shuffle=False)``), with the quantiser
(utils.uquantize / utils.linear_quantize) running as a CUDA kernel (``srnn_quantize``).

The reference keeps three arrays per partition (dataset.py:203-231): ``data`` (batch_size, N) float audio in [-1, 1],
``cond`` (batch_size, n_frames, cond_dim) normalised conditioners (float64) and ``global_spk`` (batch_size, n_frames)
speaker ids; item ``index`` is row ``index % batch_size`` of TBPTT chunk ``index // batch_size``.  ``TBPTTBatcher.batch(k)``
returns what the reference's loader yields for chunk k -- ``(input_sequences, reset, target_sequences, cond, spk)`` -- as
device tensors ready for ``Predictor.forward`` / ``sequence_nll_loss_bits``.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def quantize(samples, q_levels, ulaw=True):
    """utils.uquantize / utils.linear_quantize on the GPU: float32 (rows, cols) CUDA tensor -> int64 (rows, cols)."""
    if samples.device.type != "cuda":
        raise L.SrnnError("quantize: the B200 path has no CPU fallback, pass a CUDA tensor")
    x = samples.to(torch.float32)
    if x.dim() == 1:
        x = x.unsqueeze(0)
    if ulaw or x.stride(-1) != 1:
        x = x.contiguous()
    rows, cols = x.shape
    q = torch.empty(rows, cols, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().srnn_quantize(x.data_ptr(), rows, cols, x.stride(0), int(q_levels), int(bool(ulaw)), q.data_ptr(),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return q.reshape(samples.shape) if samples.dim() != 1 else q[0]


class TBPTTBatcher:
    """dataset.py:238-289 for a whole batch at a time.  ``overlap_len`` = model.lookback, ``cond_len`` = 80 samples per
    conditioner frame (train.py:41), ``seq_len`` a multiple of ``cond_len``."""

    def __init__(self, data, cond, global_spk, overlap_len, q_levels, ulaw, seq_len, batch_size, cond_len, device="cuda"):
        data = np.asarray(data)
        if data.shape[0] != batch_size:
            raise ValueError("data must be (batch_size, N) as stored by the reference (dataset.py:203)")
        self.overlap_len, self.q_levels, self.ulaw = int(overlap_len), int(q_levels), bool(ulaw)
        self.seq_len, self.batch_size, self.cond_len = int(seq_len), int(batch_size), int(cond_len)
        self.length = int(np.prod(data.shape)) // self.seq_len                        # dataset.py:225 (items)
        self.device = torch.device(device)
        # ulaw: float audio in [-1, 1], quantised here.  Linear: the reference quantises every FILE once when the dataset is
        # created (dataset.py:129-130) and __getitem__ only casts the stored levels with .long() (dataset.py:249-251).
        self.data = (torch.as_tensor(data, dtype=torch.float32) if self.ulaw else
                     torch.as_tensor(np.asarray(data).astype(np.int64))).to(self.device)
        self.cond = torch.as_tensor(np.asarray(cond)).to(self.device)                 # float64 as stored (dataset.py:274)
        self.cond_in_seq = self.seq_len // self.cond_len                              # dataset.py:256
        # majority-vote speaker of every (chunk, row): dataset.py:277-281; tiny, done once on the host like the reference
        spk = np.asarray(global_spk).astype(int)
        n_chunks = len(self)
        votes = np.zeros((n_chunks, self.batch_size), dtype=np.int64)
        for k in range(n_chunks):
            lo = k * self.cond_in_seq + 1
            for r in range(self.batch_size):
                votes[k, r] = np.argmax(np.bincount(spk[r][lo:lo + self.cond_in_seq]))
        self.spk = torch.from_numpy(votes).to(self.device)
        # mu-law quantisation is element-wise: the whole stream is quantised once (one kernel launch); linear-quantised
        # data is stored as levels already (use ``quantize(file, q_levels, False)`` per file when building such a dataset)
        self.qdata = quantize(self.data, self.q_levels, True) if self.ulaw else self.data

    def __len__(self):
        """number of TBPTT chunks (batches); the reference's len() counts items = chunks * batch_size."""
        return self.length // self.batch_size

    def batch(self, n_batch):
        if not (0 <= n_batch < len(self)):
            raise IndexError(n_batch)
        start_data = n_batch * self.seq_len                                           # dataset.py:245-247
        start_target = start_data + self.overlap_len
        end_target = start_target + self.seq_len
        data = self.qdata[:, start_data:end_target - 1].contiguous()              # dataset.py:249-253
        target = self.qdata[:, start_target:end_target].contiguous()
        reset = n_batch == 0                                                          # dataset.py:258-263
        from_cond = n_batch * self.cond_in_seq + 1
        cond = self.cond[:, from_cond:from_cond + self.cond_in_seq].contiguous()
        spk = self.spk[n_batch].reshape(self.batch_size, 1)
        return data, reset, target, cond, spk

    def __iter__(self):
        for k in range(len(self)):
            yield self.batch(k)
