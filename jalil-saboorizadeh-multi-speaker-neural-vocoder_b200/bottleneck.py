"""Bottle-neck voice-conversion variant (BASELINE.json configs[4]; ``run_sampleneck.sh`` ``--ind_cond_dim 30``).

PARITY UNPINNED.  The variant's source lives on the reference's ``bottle-neck`` git branch, which is not part of the
reference tree (``run_sampleneck.sh:2`` does ``git reset --hard origin/bottle-neck``).  What is built here follows the only
description available, the thesis (doc/Barbany_report.pdf section 3.2.1, Fig. 3.4): the top tier's single
``cond_expand`` Conv1d(cond_dim -> dim) becomes a chain of k = 1 Conv1d layers
``cond_dim -> 40 -> 30 -> 20 -> ind_cond_dim -> dim`` with a ReLU after every chain layer (look-ahead off, weight-norm on).
Everything else -- GRU tiers, upsampling, sample-level MLP -- is the configs[1] model.

The chain up to ``ind_cond_dim`` runs as one CUDA kernel (``srnn_cond_chain_fwd``); its last layer IS the ``cond_expand`` of a
``SampleRNN`` built with ``cond_dim = ind_cond_dim``, so generation and the teacher-forced forward pass reuse the whole hot
path unchanged.  Training the chain itself (gradients w.r.t. its weights) is not provided: the library does not return
dL/dcond.
"""
import ctypes as C
import math

import torch
from torch import nn as tnn

from . import _lib as L
from .model import Generator, Predictor, SampleRNN, _Conv, _kaiming_uniform_, _stream

CHAIN_HIDDEN = (40, 30, 20)            # thesis Fig. 3.4


class BottleneckConditioner(tnn.Module):
    """k = 1 Conv1d chain cond_dim -> 40 -> 30 -> 20 -> ind_cond_dim, ReLU after each layer."""

    def __init__(self, cond_dim, ind_cond_dim, weight_norm=True, hidden=CHAIN_HIDDEN):
        super().__init__()
        self.dims = [int(cond_dim)] + [int(h) for h in hidden] + [int(ind_cond_dim)]
        if len(self.dims) - 1 > L.MAX_CHAIN or max(self.dims) > 128:
            raise ValueError("conditioner chain too long / too wide for srnn_cond_chain_fwd")
        self.layers = tnn.ModuleList([_Conv((o, i, 1), _kaiming_uniform_, weight_norm, o)
                                      for i, o in zip(self.dims[:-1], self.dims[1:])])

    def forward(self, cond):
        """cond (..., cond_dim) float -> (..., ind_cond_dim) float32 on the parameters' CUDA device."""
        dev = self.layers[0].bias.device
        if dev.type != "cuda":
            raise L.SrnnError("BottleneckConditioner parameters are on %s: the B200 path has no CPU fallback" % dev)
        x = torch.as_tensor(cond).to(device=dev, dtype=torch.float32).contiguous()
        if x.shape[-1] != self.dims[0]:
            raise ValueError("cond has width %d, the chain expects %d" % (x.shape[-1], self.dims[0]))
        rows = x.numel() // self.dims[0]
        chain = L.CondChain()
        chain.n_layers = len(self.layers)
        for i, d in enumerate(self.dims):
            chain.dims[i] = d
        for i, layer in enumerate(self.layers):
            chain.layers[i] = layer.c_params()
        out = torch.empty(*x.shape[:-1], self.dims[-1], device=dev, dtype=torch.float32)
        scratch = torch.empty(sum(i * o for i, o in zip(self.dims[:-1], self.dims[1:])), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.check(L.load().srnn_cond_chain_fwd(C.byref(chain), x.data_ptr(), rows, out.data_ptr(), scratch.data_ptr(), _stream()))
        return out


class BottleneckSampleRNN(tnn.Module):
    """``SampleRNN`` of the bottle-neck branch: ``.conditioner`` (the chain) in front of ``.core`` (a SampleRNN whose
    ``cond_dim`` is ``ind_cond_dim``; its top tier's ``cond_expand`` is the chain's last layer)."""

    def __init__(self, frame_sizes, n_rnn, dim, learn_h0, q_levels, ulaw, weight_norm, cond_dim, spk_dim, ind_cond_dim=30):
        super().__init__()
        self.conditioner = BottleneckConditioner(cond_dim, ind_cond_dim, weight_norm)
        self.core = SampleRNN(frame_sizes, n_rnn, dim, learn_h0, q_levels, ulaw, weight_norm, ind_cond_dim, spk_dim)
        self.cond_dim, self.ind_cond_dim = cond_dim, ind_cond_dim

    @property
    def lookback(self):
        return self.core.lookback


class BottleneckPredictor(tnn.Module):
    """Teacher-forced forward pass (model.py:357-436) of the bottle-neck model; the chain is applied to ``cond`` first."""

    def __init__(self, model, mode=L.MODE_FP32):
        super().__init__()
        self.model = model
        self.core = Predictor(model.core, mode=mode)

    def forward(self, input_sequences, reset, cond, spk, writer=None, iterations=None):
        with torch.no_grad():
            c = self.model.conditioner(cond)
        return self.core(input_sequences, reset, c, spk, writer, iterations)


class BottleneckGenerator:
    """``Generator`` (model.py:439-520) of the bottle-neck model: same call signature, chain applied to ``cond`` first."""

    def __init__(self, model, cuda=False, mode=L.MODE_FP32):
        self.model = model
        self.core = Generator(model.core, cuda=cuda, mode=mode)

    @torch.no_grad()
    def __call__(self, n_seqs, seq_len, cond, spk, **kw):
        return self.core(n_seqs, seq_len, self.model.conditioner(cond), spk, **kw)
