"""CPU oracle for the SampleRNN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``jalil-saboorizadeh-multi-speaker-neural-vocoder_b200``) never does; it fails loudly when the CUDA
library is missing.

It is a plain restatement (explicit tensor algebra on CPU, no ``nn.GRU`` / ``Conv1d`` /
``weight_norm`` modules) of the reference algorithm, each function citing the reference
``file:line`` it follows (paths relative to the reference repo root).

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the *unmodified* reference
``model.py`` / ``nn.py`` / ``utils.py`` / ``optim.py`` in the build container, runs it on seeded
inputs and commits the input/output vectors under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function here against those vectors.

The one place where the reference is implementation-defined is ``Tensor.multinomial`` (model.py:517).
The golden run replaces it by the *defined sampler* below (``sample_rows``), which is also what the
CUDA kernels implement, so "same uniforms -> same indices" is a meaningful bit-exact statement.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

MU = 255.0                      # utils.py:29
LOG_MU1 = 5.5451774444795623    # utils.py:30  log(1 + MU)


# ----------------------------------------------------------------------------------------------
# configuration / weights
# ----------------------------------------------------------------------------------------------
@dataclass
class Config:
    """Constructor arguments of the reference ``SampleRNN`` (model.py:20)."""
    frame_sizes: Sequence[int]
    n_rnn: int
    dim: int
    learn_h0: bool = True
    q_levels: int = 256
    ulaw: bool = True
    weight_norm: bool = True
    cond_dim: int = 43
    spk_dim: int = 6

    @property
    def ns_frame_samples(self) -> List[int]:          # model.py:34
        return [int(x) for x in np.cumprod(list(self.frame_sizes))]

    @property
    def lookback(self) -> int:                        # model.py:60-62
        return self.ns_frame_samples[-1]


def _wn(sd: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """Effective weight of a (possibly weight-normalised) layer.

    torch ``weight_norm(dim=0)`` as applied at model.py:119-121,130-131,177-178,303-306:
    ``w = g * v / ||v||`` with the norm taken over every dim except 0 (per output channel for
    Conv1d, per *input* channel for ConvTranspose1d whose weight is (in, out, k))."""
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"]
    g, v = sd[prefix + ".weight_g"], sd[prefix + ".weight_v"]
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
    return v * (g / norm)


@dataclass
class TierWeights:
    frame_size: int
    n_frame_samples: int
    h0: torch.Tensor                     # (n_rnn, H)
    w_in: torch.Tensor                   # (H, n)
    b_in: torch.Tensor                   # (H,)
    w_ih: List[torch.Tensor]             # n_rnn x (3H, H)   gate row blocks r, z, n
    w_hh: List[torch.Tensor]
    b_ih: List[torch.Tensor]
    b_hh: List[torch.Tensor]
    w_up: torch.Tensor                   # (H_in, H_out, k)
    b_up: torch.Tensor                   # (H_out, k)
    w_cond: Optional[torch.Tensor] = None   # (H, cond_dim)
    b_cond: Optional[torch.Tensor] = None
    spk_emb: Optional[torch.Tensor] = None  # (spk_dim, spk_dim)
    w_spk: Optional[torch.Tensor] = None    # (H, spk_dim)
    b_spk: Optional[torch.Tensor] = None
    w_up_mat: Optional[torch.Tensor] = None  # cache, see learned_upsampling


@dataclass
class Weights:
    cfg: Config
    tiers: List[TierWeights]
    emb: torch.Tensor                    # (Q, Q)
    w_mlp_in: torch.Tensor               # (H, Q, FS0)
    w_mlp_hid: torch.Tensor              # (H, H)
    b_mlp_hid: torch.Tensor
    w_mlp_out: torch.Tensor              # (Q, H)
    b_mlp_out: torch.Tensor
    w_mlp_in_taps: Optional[torch.Tensor] = None   # cache (FS0, Q, H): tap j = W_in[:, :, j]^T, contiguous


def unpack_state_dict(sd: Dict[str, torch.Tensor], cfg: Config, dtype=torch.float32) -> Weights:
    """``Predictor.state_dict()`` (keys prefixed ``model.``) -> plain weights.

    Works on tensors that require grad (the weight-norm fold is differentiable)."""
    sd = {k[len("model."):] if k.startswith("model.") else k: v.to(dtype) for k, v in sd.items()}
    tiers = []
    n_tiers = len(cfg.frame_sizes)
    for i, (fs, n) in enumerate(zip(cfg.frame_sizes, cfg.ns_frame_samples)):
        p = f"frame_level_rnns.{i}"
        tw = TierWeights(
            frame_size=int(fs), n_frame_samples=int(n),
            h0=sd[p + ".h0"],
            w_in=_wn(sd, p + ".input_expand").squeeze(-1), b_in=sd[p + ".input_expand.bias"],
            w_ih=[sd[f"{p}.rnn.weight_ih_l{l}"] for l in range(cfg.n_rnn)],
            w_hh=[sd[f"{p}.rnn.weight_hh_l{l}"] for l in range(cfg.n_rnn)],
            b_ih=[sd[f"{p}.rnn.bias_ih_l{l}"] for l in range(cfg.n_rnn)],
            b_hh=[sd[f"{p}.rnn.bias_hh_l{l}"] for l in range(cfg.n_rnn)],
            w_up=_wn(sd, p + ".upsampling.conv_t"), b_up=sd[p + ".upsampling.bias"],
        )
        if i == n_tiers - 1:                                  # model.py:46-47 only the top tier is conditioned
            tw.w_cond = _wn(sd, p + ".cond_expand").squeeze(-1)
            tw.b_cond = sd[p + ".cond_expand.bias"]
            tw.spk_emb = sd[p + ".spk_embedding.weight"]
            tw.w_spk = _wn(sd, p + ".spk_expand").squeeze(-1)
            tw.b_spk = sd[p + ".spk_expand.bias"]
        tiers.append(tw)
    m = "sample_level_mlp"
    return Weights(
        cfg=cfg, tiers=tiers,
        emb=sd[m + ".embedding.weight"],
        w_mlp_in=_wn(sd, m + ".input"),
        w_mlp_hid=_wn(sd, m + ".hidden").squeeze(-1), b_mlp_hid=sd[m + ".hidden.bias"],
        w_mlp_out=_wn(sd, m + ".output").squeeze(-1), b_mlp_out=sd[m + ".output.bias"],
    )


# ----------------------------------------------------------------------------------------------
# element-wise pieces
# ----------------------------------------------------------------------------------------------
def dequantize(q: torch.Tensor, q_levels: int, ulaw: bool, dtype=torch.float32) -> torch.Tensor:
    """utils.py:18-19 (linear) / utils.py:54-55,39-42,62-63 (mu-law): int -> [-1, 1)."""
    x = q.to(dtype)
    if not ulaw:
        return x / (q_levels / 2) - 1
    c = x * 2.0 / q_levels - 1.0                       # imidrise
    return torch.sign(c) * (torch.exp(torch.abs(c) * LOG_MU1) - 1) / MU   # iulaw (ignores its mu arg)


def dequant_lut(q_levels: int, ulaw: bool) -> np.ndarray:
    """The q_levels-entry table ``2 * dequantize(q)`` that model.py:385,471 feed to the frame tiers."""
    q = torch.arange(q_levels)
    return (2 * dequantize(q, q_levels, ulaw)).numpy().astype(np.float32)


def q_zero(q_levels: int) -> int:                      # utils.py:22-23
    return q_levels // 2


# ----------------------------------------------------------------------------------------------
# tiers
# ----------------------------------------------------------------------------------------------
def gru_forward(x: torch.Tensor, h: torch.Tensor, tw: TierWeights):
    """``nn.GRU(H, H, n_rnn, batch_first=True)`` at model.py:154-159,244.

    x (B, F, H), h (n_rnn, B, H) -> (out (B, F, H), h_new (n_rnn, B, H)).
    r = s(gi_r+gh_r); z = s(gi_z+gh_z); n = tanh(gi_n + r*gh_n); h' = (1-z)*n + z*h."""
    H = x.shape[-1]
    new_h = []
    layer_in = x
    for l in range(len(tw.w_ih)):
        gi_all = layer_in @ tw.w_ih[l].t() + tw.b_ih[l]          # time-batched input projection
        hl = h[l]
        outs = []
        for t in range(x.shape[1]):
            gi = gi_all[:, t]
            gh = hl @ tw.w_hh[l].t() + tw.b_hh[l]
            r = torch.sigmoid(gi[:, :H] + gh[:, :H])
            z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
            n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
            hl = (1 - z) * n + z * hl
            outs.append(hl)
        layer_in = torch.stack(outs, dim=1)
        new_h.append(hl)
    return layer_in, torch.stack(new_h, dim=0)


def learned_upsampling(x: torch.Tensor, tw: TierWeights) -> torch.Tensor:
    """nn.py:33-43: ConvTranspose1d(stride=kernel=k, no bias) + per-(channel, phase) bias.

    x (B, F, H) -> (B, F*k, H):  out[b, t*k+j, o] = sum_c x[b,t,c] W[c,o,j] + bias[o,j]."""
    B, F, H = x.shape
    k = tw.frame_size
    if tw.w_up_mat is None:      # (H_in, k*H_out), column j*H+o = W[c,o,j]; cached contiguous copy (speed only)
        tw.w_up_mat = tw.w_up.permute(0, 2, 1).reshape(H, -1).contiguous()
    y = x.reshape(B * F, H) @ tw.w_up_mat + tw.b_up.t().reshape(1, -1)
    return y.reshape(B, F * k, -1)


def frame_level_forward(tw: TierWeights, prev: torch.Tensor, upper: Optional[torch.Tensor],
                        hidden: Optional[torch.Tensor], cond: Optional[torch.Tensor],
                        spk: Optional[torch.Tensor]):
    """``FrameLevelRNN.forward`` model.py:180-263.

    prev (B, F, n) float in [-2, 2); upper (B, F, H) or None (top tier); hidden (n_rnn, B, H) or None;
    cond (B, F, cond_dim); spk (B, 1) int.  Returns (out (B, F*fs, H), hidden)."""
    B = prev.shape[0]
    x = prev @ tw.w_in.t() + tw.b_in                                     # model.py:196-198
    if upper is not None:
        x = x + upper                                                    # model.py:199-200
    else:
        x = x + (cond.to(x.dtype) @ tw.w_cond.t() + tw.b_cond)           # model.py:202-203
        e = tw.spk_emb[spk.long().reshape(B)]                            # model.py:208  (B, spk_dim)
        x = x + (e @ tw.w_spk.t() + tw.b_spk).unsqueeze(1)               # model.py:217-218 broadcast over frames
    if hidden is None:                                                   # model.py:222-228
        hidden = tw.h0.unsqueeze(1).expand(-1, B, -1)
    out, hidden = gru_forward(x, hidden, tw)                             # model.py:244
    return learned_upsampling(out, tw), hidden                           # model.py:249-251


def mlp_logits(w: Weights, prev_q: torch.Tensor, upper: torch.Tensor) -> torch.Tensor:
    """``SampleLevelMLP.forward`` model.py:308-325 up to (not including) log_softmax.

    prev_q (B, T+FS-1) int in [0, Q); upper (B, T, H) -> logits (B, T, Q)."""
    FS = w.w_mlp_in.shape[-1]
    T = upper.shape[1]
    e = w.emb[prev_q.long()]                                             # (B, T+FS-1, Q)  model.py:311-315
    if w.w_mlp_in_taps is None:
        w.w_mlp_in_taps = w.w_mlp_in.permute(2, 1, 0).contiguous()
    x = upper
    for j in range(FS):                                                  # Conv1d(Q->H, k=FS, no bias) model.py:317
        x = x + e[:, j:j + T] @ w.w_mlp_in_taps[j]
    x = torch.relu(x)
    x = torch.relu(x @ w.w_mlp_hid.t() + w.b_mlp_hid)                    # model.py:321
    return x @ w.w_mlp_out.t() + w.b_mlp_out                             # model.py:322


def mlp_forward(w: Weights, prev_q: torch.Tensor, upper: torch.Tensor) -> torch.Tensor:
    return torch.log_softmax(mlp_logits(w, prev_q, upper), dim=-1)       # model.py:324-325 (implicit dim -> -1)


def folded_table(w: Weights) -> torch.Tensor:
    """Exact algebraic fold of embedding o conv: Tbl[j, q, h] = W_in[h, :, j] . E[q, :]."""
    return torch.einsum("hej,qe->jqh", w.w_mlp_in, w.emb)


# ----------------------------------------------------------------------------------------------
# Predictor (teacher forcing) -- model.py:352-436
# ----------------------------------------------------------------------------------------------
class Predictor:
    def __init__(self, w: Weights):
        self.w = w
        self.hidden: List[Optional[torch.Tensor]] = [None] * len(w.tiers)   # Runner.hidden_states model.py:335-336

    def reset_hidden_states(self):
        self.hidden = [None] * len(self.w.tiers)

    def forward(self, input_sequences: torch.Tensor, reset: bool, cond: torch.Tensor, spk: torch.Tensor,
                return_logits: bool = False) -> torch.Tensor:
        """input_sequences (B, lookback+T-1) int; cond (B, T/lookback, cond_dim); spk (B, 1) -> (B, T, Q) log-probs."""
        w, cfg = self.w, self.w.cfg
        if reset:
            self.reset_hidden_states()                                   # model.py:358-359
        B = input_sequences.shape[0]
        L = input_sequences.shape[1]
        upper = None
        for i in reversed(range(len(w.tiers))):                          # model.py:378 top tier first
            tw = w.tiers[i]
            n = tw.n_frame_samples
            sl = input_sequences[:, cfg.lookback - n: L - n + 1]         # model.py:379-380
            prev = 2 * dequantize(sl, cfg.q_levels, cfg.ulaw, w.emb.dtype)   # model.py:385-388
            prev = prev.reshape(B, -1, n)                                # model.py:406-408
            c, s = (cond, spk) if upper is None else (None, None)
            upper, h = frame_level_forward(tw, prev, upper, self.hidden[i], c, s)
            self.hidden[i] = h.detach()                                  # model.py:348 TBPTT carry
        fs0 = w.tiers[0].frame_size
        mlp_in = input_sequences[:, cfg.lookback - fs0:]                 # model.py:422-423
        if return_logits:
            return mlp_logits(w, mlp_in, upper)
        return mlp_forward(w, mlp_in, upper)                             # model.py:434-436


def nll_bits(logp: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """nn.py:66-70: mean NLL over B*T in bits."""
    Q = logp.shape[-1]
    picked = logp.reshape(-1, Q).gather(1, target.reshape(-1, 1).long())
    return -picked.mean() * math.log(math.e, 2)


# ----------------------------------------------------------------------------------------------
# the defined sampler (replaces Tensor.multinomial at model.py:517)
# ----------------------------------------------------------------------------------------------
def sample_rows(p: np.ndarray, u: np.ndarray) -> np.ndarray:
    """Inverse-CDF draw from unnormalised non-negative rows ``p`` (R, 256) fp32 with uniforms ``u`` (R,) fp32.

    Defined summation order (all arithmetic fp32, mirrored exactly by the CUDA sampler):
      * the row is split into 32 chunks of 8 consecutive entries; inside chunk l the running sums
        loc[l][i] = (((p[8l] + p[8l+1]) + ...) + p[8l+i]) are sequential;
      * chunk totals S[l] = loc[l][7] go through a 5-step Kogge-Stone inclusive scan
        (d = 1, 2, 4, 8, 16:  S[l] += S[l-d] for l >= d, all lanes at once);
      * cdf[8l+i] = (S[l-1] if l > 0 else 0) + loc[l][i];   total = S[31];
      * idx = min(255, #{k : cdf[k] <= u * total}).
    """
    p = np.ascontiguousarray(p, dtype=np.float32)
    R, Q = p.shape
    assert Q == 256
    c = p.reshape(R, 32, 8)
    loc = np.empty_like(c)
    acc = c[:, :, 0].copy()
    loc[:, :, 0] = acc
    for i in range(1, 8):
        acc = (acc + c[:, :, i]).astype(np.float32)
        loc[:, :, i] = acc
    S = loc[:, :, 7].copy()
    for d in (1, 2, 4, 8, 16):
        nxt = S.copy()
        nxt[:, d:] = (S[:, d:] + S[:, :-d]).astype(np.float32)
        S = nxt
    excl = np.zeros_like(S)
    excl[:, 1:] = S[:, :-1]
    cdf = (excl[:, :, None] + loc).astype(np.float32).reshape(R, Q)
    thr = (u.astype(np.float32) * S[:, 31]).astype(np.float32)
    idx = (cdf <= thr[:, None]).sum(axis=1)
    return np.minimum(idx, Q - 1).astype(np.int64)


# ----------------------------------------------------------------------------------------------
# Generator (autoregressive) -- model.py:439-520
# ----------------------------------------------------------------------------------------------
class Generator:
    def __init__(self, w: Weights):
        self.w = w

    @torch.no_grad()
    def __call__(self, n_seqs: int, cond, spk, uniforms: np.ndarray, return_logp: bool = False):
        """cond: (n_cond, cond_dim) shared by all sequences (reference form, model.py:484-487) or
        (n_seqs, n_cond, cond_dim) per utterance (extension).  spk: int or (n_seqs,) ints.
        uniforms: (n_cond*lookback, n_seqs) fp32, indexed [t, b].
        Returns int64 samples (n_seqs, n_cond*lookback) [and the per-step log-probs (B, T, Q)]."""
        w, cfg = self.w, self.w.cfg
        dt = w.emb.dtype
        cond = torch.as_tensor(np.asarray(cond))
        if cond.dim() == 2:
            cond = cond.unsqueeze(0).expand(n_seqs, -1, -1)
        spk = torch.as_tensor(np.asarray(spk)).reshape(-1)
        if spk.numel() == 1:
            spk = spk.expand(n_seqs)
        spk = spk.reshape(n_seqs, 1)
        n_cond = cond.shape[1]
        lookback = cfg.lookback
        seq_len = n_cond * lookback                                      # model.py:455 (caller's seq_len ignored)
        fs0 = w.tiers[0].n_frame_samples
        seq = torch.full((n_seqs, lookback + seq_len), q_zero(cfg.q_levels), dtype=torch.int64)   # model.py:459
        hidden: List[Optional[torch.Tensor]] = [None] * len(w.tiers)
        outs: List[Optional[torch.Tensor]] = [None] * len(w.tiers)
        logps = []
        for i in range(lookback, lookback + seq_len):                    # model.py:462
            for ti in reversed(range(len(w.tiers))):
                tw = w.tiers[ti]
                n = tw.n_frame_samples
                if i % n != 0:                                           # model.py:465
                    continue
                prev = (2 * dequantize(seq[:, i - n:i], cfg.q_levels, cfg.ulaw, dt)).unsqueeze(1)   # 470-476
                if ti == len(w.tiers) - 1:
                    j = i // lookback - 1                                # model.py:483
                    c, s, upper = cond[:, j:j + 1, :], spk, None
                else:
                    fi = (i // n) % w.tiers[ti + 1].frame_size           # model.py:491-492
                    c, s, upper = None, None, outs[ti + 1][:, fi:fi + 1, :]
                outs[ti], hidden[ti] = frame_level_forward(tw, prev, upper, hidden[ti], c, s)   # 500-502
            prev_q = seq[:, i - fs0:i]                                   # model.py:504-507
            upper = outs[0][:, i % fs0: i % fs0 + 1, :]                  # model.py:511-513
            logp = mlp_forward(w, prev_q, upper).squeeze(1)              # model.py:514-516
            p = torch.exp(logp).float().numpy()
            seq[:, i] = torch.from_numpy(sample_rows(p, uniforms[i - lookback]))   # model.py:517
            if return_logp:
                logps.append(logp)
        samples = seq[:, lookback:]
        if return_logp:
            return samples, torch.stack(logps, dim=1)
        return samples

    def audio(self, samples: torch.Tensor) -> torch.Tensor:
        """model.py:520: what ``Generator.__call__`` returns -- dequantised audio in [-1, 1)."""
        return dequantize(samples, self.w.cfg.q_levels, self.w.cfg.ulaw)


# ----------------------------------------------------------------------------------------------
# training step -- trainer/__init__.py:99-112, optim.py:4-21, torch.optim.Adam (train.py:238)
# ----------------------------------------------------------------------------------------------
@dataclass
class AdamState:
    step: int = 0
    m: Dict[str, torch.Tensor] = field(default_factory=dict)
    v: Dict[str, torch.Tensor] = field(default_factory=dict)


def loss_and_grads(sd: Dict[str, torch.Tensor], cfg: Config, hidden, input_sequences, reset, cond, spk, target,
                   dtype=torch.float32):
    """Forward + ``sequence_nll_loss_bits`` + backward w.r.t. every state_dict entry.

    Gradients of unused parameters (``h0`` on non-reset batches) are ZERO tensors, reproducing
    torch-0.4 ``zero_grad`` semantics that the reference was written against.
    Returns (loss, grads dict, new hidden list, log-probs)."""
    leaves = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
    w = unpack_state_dict(leaves, cfg, dtype)
    pr = Predictor(w)
    pr.hidden = list(hidden) if hidden is not None else [None] * len(w.tiers)
    logp = pr.forward(input_sequences, reset, cond, spk)
    loss = nll_bits(logp, target)
    names = list(leaves.keys())
    gs = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, gs)}
    return loss.detach(), grads, pr.hidden, logp.detach()


def clamp_adam_step(sd: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], st: AdamState, lr: float,
                    beta1=0.9, beta2=0.999, eps=1e-8, trainable=None):
    """optim.py:10-13 element-wise clamp to [-1, 1], then torch.optim.Adam (no weight decay, no amsgrad)."""
    st.step += 1
    bc1 = 1 - beta1 ** st.step
    bc2 = 1 - beta2 ** st.step
    out = {}
    for k, p in sd.items():
        if trainable is not None and k not in trainable:
            out[k] = p
            continue
        g = grads[k].clamp(-1, 1)
        m = st.m.get(k, torch.zeros_like(p)) * beta1 + (1 - beta1) * g
        v = st.v.get(k, torch.zeros_like(p)) * beta2 + (1 - beta2) * g * g
        st.m[k], st.v[k] = m, v
        denom = v.sqrt() / math.sqrt(bc2) + eps
        out[k] = p - (lr / bc1) * m / denom
    return out
