"""Host-side mirror of the reference ``model.py`` API over the B200 C-ABI library.

``SampleRNN`` / ``Predictor`` / ``Generator`` / ``Runner`` keep the reference's constructor and call signatures
and its ``state_dict`` layout, so ``from model import SampleRNN, Predictor, Generator``
(train.py:3, generate.py:1) can be pointed here.  All arithmetic happens in ``libsrnn_b200.so``
(include/srnn_b200.h); torch only owns device memory and streams.  There is no CPU fallback: parameters must
live on a CUDA device and the library must be built, otherwise calls raise.

Reference quirks deliberately NOT reproduced: the ``<spk>.txt`` debug dump
(model.py:209-214), the global cuDNN toggle (model.py:448-449,518), per-step prints (model.py:456-457,469).
"""
import ctypes as C
import math
import weakref

import numpy as np
import torch
from torch import nn as tnn

from . import _lib as L


# ------------------------------------------------------------------------------------------------------------------
# parameter holders with the reference's names / shapes / initialisers
# ------------------------------------------------------------------------------------------------------------------
def _kaiming_uniform_(w):                      # init.kaiming_uniform, model.py:98,115,127,285,292
    return tnn.init.kaiming_uniform_(w)


def _lecun_uniform_(w):                        # nn.py:46-48
    fan_in = tnn.init._calculate_correct_fan(w, "fan_in")
    return tnn.init.uniform_(w, -math.sqrt(3 / fan_in), math.sqrt(3 / fan_in))


def _concat_init_(w, inits):                   # nn.py:51-63
    with torch.no_grad():
        length, fan_out = w.shape
        fan_in = length // len(inits)
        for i, init in enumerate(inits):
            chunk = w.new_empty(fan_in, fan_out)
            init(chunk)
            w[i * fan_in:(i + 1) * fan_in] = chunk


class _Conv(tnn.Module):
    """Parameters of a Conv1d / ConvTranspose1d: ``weight`` or the weight_norm pair ``weight_g``/``weight_v``."""

    def __init__(self, shape, init, weight_norm, bias_len=None):
        super().__init__()
        w = torch.empty(*shape)
        init(w)
        if weight_norm:                        # torch weight_norm(dim=0): g = ||v|| over every dim but 0
            g = w.reshape(shape[0], -1).norm(dim=1).reshape(shape[0], *([1] * (len(shape) - 1)))
            self.weight_g = tnn.Parameter(g)
            self.weight_v = tnn.Parameter(w)
        else:
            self.weight = tnn.Parameter(w)
        if bias_len is not None:
            self.bias = tnn.Parameter(torch.zeros(bias_len))

    def c_params(self, bias=None, ptr=None):
        """``ptr`` maps a parameter tensor to the device pointer to bind (its data by default, its gradient buffer for
        the backward pass; None skips the entry)."""
        ptr = ptr or (lambda t: t.data_ptr())
        p = L.ConvParams()
        if hasattr(self, "weight"):
            p.weight = ptr(self.weight)
        else:
            p.weight_g, p.weight_v = ptr(self.weight_g), ptr(self.weight_v)
        b = bias if bias is not None else getattr(self, "bias", None)
        p.bias = ptr(b) if b is not None else None
        return p


class _Embedding(tnn.Module):
    def __init__(self, n, d):
        super().__init__()
        self.weight = tnn.Parameter(torch.randn(n, d))        # torch.nn.Embedding default N(0,1)


class _GRU(tnn.Module):
    """Parameter holder with ``nn.GRU(H, H, n_rnn)`` names; init per model.py:154-165."""

    def __init__(self, dim, n_rnn):
        super().__init__()
        for l in range(n_rnn):
            w_ih, w_hh = torch.empty(3 * dim, dim), torch.empty(3 * dim, dim)
            _concat_init_(w_ih, [_lecun_uniform_] * 3)
            _concat_init_(w_hh, [_lecun_uniform_, _lecun_uniform_, tnn.init.orthogonal_])
            setattr(self, f"weight_ih_l{l}", tnn.Parameter(w_ih))
            setattr(self, f"weight_hh_l{l}", tnn.Parameter(w_hh))
            setattr(self, f"bias_ih_l{l}", tnn.Parameter(torch.zeros(3 * dim)))
            setattr(self, f"bias_hh_l{l}", tnn.Parameter(torch.zeros(3 * dim)))


class LearnedUpsampling1d(tnn.Module):
    """nn.py:7-43.  ``conv_t`` is ALWAYS weight-normalised (model.py:177 tests the imported function)."""

    def __init__(self, dim, kernel_size):
        super().__init__()
        bound = math.sqrt(6 / dim)                             # model.py:172-175
        self.bias = tnn.Parameter(torch.zeros(dim, kernel_size))
        self.conv_t = _Conv((dim, dim, kernel_size), lambda w: tnn.init.uniform_(w, -bound, bound), True)


_OWNERS = {}          # id(FrameLevelRNN | SampleLevelMLP) -> (weakref to the owning SampleRNN, tier index or -1)


def _owner_of(module):
    ref, tier = _OWNERS.get(id(module), (None, -1))
    model = ref() if ref is not None else None
    if model is not None and not (module is model.sample_level_mlp or any(module is r for r in model.frame_level_rnns)):
        model = None                                   # a recycled id
    return model, tier


class FrameLevelRNN(tnn.Module):
    """model.py:65-178 (parameters); the forward lives in the CUDA library."""

    def __init__(self, frame_size, n_frame_samples, n_rnn, dim, learn_h0, is_cond, cond_dim, spk_dim, w_norm, qrnn):
        super().__init__()
        self.frame_size, self.n_frame_samples, self.dim = frame_size, n_frame_samples, dim
        self.cond_dim, self.spk_dim, self.weight_norm, self.qrnn = cond_dim, spk_dim, w_norm, qrnn
        h0 = torch.zeros(n_rnn, dim)
        if learn_h0:
            self.h0 = tnn.Parameter(h0)
        else:
            self.register_buffer("h0", h0)
        self.input_expand = _Conv((dim, n_frame_samples, 1), _kaiming_uniform_, w_norm, dim)
        if is_cond:
            self.cond_expand = _Conv((dim, cond_dim, 1), _kaiming_uniform_, w_norm, dim)
            self.spk_embedding = _Embedding(spk_dim, spk_dim)
            self.spk_expand = _Conv((dim, spk_dim, 1), _kaiming_uniform_, w_norm, dim)
        else:
            self.cond_expand = self.spk_expand = self.spk_embedding = None
        self.rnn = _GRU(dim, n_rnn)
        self.upsampling = LearnedUpsampling1d(dim, frame_size)

    def forward(self, prev_samples, upper_tier_conditioning, hidden, cond, spk, writer=None, iterations=None):
        """model.py:180-263 as one C-ABI call (``srnn_tier_fwd``): prev_samples (B, F, n_frame_samples) float =
        ``2 * dequantize(window)``, upper_tier_conditioning (B, F, dim) or None on the top tier (which takes cond
        (B, F, cond_dim) and spk (B, 1)); hidden (n_rnn, B, dim) or None (start from h0).  Returns (output (B, F*frame_size,
        dim), new hidden).  Inference form: Predictor.forward / Generator are the fused (and differentiable) paths; the mode
        is the owning SampleRNN's ``module_mode`` (fp32, or the split-bf16 tensor-core parity mode)."""
        model, tier = _owner_of(self)
        if model is None:
            raise L.SrnnError("FrameLevelRNN.forward needs the SampleRNN that owns the tier (packed weights live in its context)")
        h = model._ensure_packed()
        dev = model._ctx_device
        B, F, n = prev_samples.shape
        if n != self.n_frame_samples:
            raise ValueError("prev_samples must be (B, F, %d)" % self.n_frame_samples)
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        prev = f32(prev_samples)
        top = self.cond_expand is not None
        if top == (upper_tier_conditioning is not None):
            raise ValueError("the top tier takes cond + spk, the lower tiers the upper tier's conditioning (model.py:199-217)")
        up = cnd = sp = None
        if top:
            if tuple(cond.shape) != (B, F, self.cond_dim):
                raise ValueError("cond must be (B, F, %d), got %s" % (self.cond_dim, tuple(cond.shape)))
            cnd = f32(cond)
            sp = spk.detach().reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
            if sp.numel() != B or int(sp.min()) < 0 or int(sp.max()) >= self.spk_dim:
                raise ValueError("spk must hold one id in [0, %d) per utterance" % self.spk_dim)
        else:
            if tuple(upper_tier_conditioning.shape) != (B, F, self.dim):
                raise ValueError("upper_tier_conditioning must be (B, F, %d)" % self.dim)
            up = f32(upper_tier_conditioning)
        n_rnn = self.h0.shape[0]
        reset = hidden is None
        hid = torch.empty(n_rnn, B, self.dim, device=dev, dtype=torch.float32) if reset else f32(hidden).clone()
        out = torch.empty(B, F * self.frame_size, self.dim, device=dev, dtype=torch.float32)
        ptr = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(dev):
            L.check(L.load().srnn_tier_fwd(h, tier, B, F, ptr(prev), ptr(up), ptr(cnd), ptr(sp), hid.data_ptr(), int(reset),
                                           out.data_ptr(), model.module_mode, _stream()))
        return out, hid


class SampleLevelMLP(tnn.Module):
    """model.py:266-306 (parameters)."""

    def __init__(self, frame_size, dim, q_levels, wnorm):
        super().__init__()
        self.q_levels, self.weight_norm = q_levels, wnorm
        self.embedding = _Embedding(q_levels, q_levels)
        self.input = _Conv((dim, q_levels, frame_size), _kaiming_uniform_, wnorm)
        self.hidden = _Conv((dim, dim, 1), _kaiming_uniform_, wnorm, dim)
        self.output = _Conv((q_levels, dim, 1), _lecun_uniform_, wnorm, q_levels)

    def forward(self, prev_samples, upper_tier_conditioning):
        """model.py:308-325 as one C-ABI call (``srnn_mlp_fwd``): prev_samples (B, T + frame_size - 1) int64 sample indices,
        upper_tier_conditioning (B, T, dim) -> log-probabilities (B, T, q_levels).  Inference form (see FrameLevelRNN.forward)."""
        model, _ = _owner_of(self)
        if model is None:
            raise L.SrnnError("SampleLevelMLP.forward needs the SampleRNN that owns it (packed weights live in its context)")
        h = model._ensure_packed()
        dev = model._ctx_device
        B, T, H = upper_tier_conditioning.shape
        fs = model.frame_sizes[0]
        if tuple(prev_samples.shape) != (B, T + fs - 1) or H != model.dim:
            raise ValueError("prev_samples must be (B, T + %d), upper_tier_conditioning (B, T, %d)" % (fs - 1, model.dim))
        ps = prev_samples.detach().to(device=dev, dtype=torch.int64).contiguous()
        if int(ps.min()) < 0 or int(ps.max()) >= self.q_levels:
            raise ValueError("prev_samples holds indices outside [0, %d)" % self.q_levels)
        up = upper_tier_conditioning.detach().to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty(B, T, self.q_levels, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.check(L.load().srnn_mlp_fwd(h, B, T, ps.data_ptr(), up.data_ptr(), out.data_ptr(), model.module_mode, _stream()))
        return out


class SampleRNN(tnn.Module):
    """model.py:18-62."""

    def __init__(self, frame_sizes, n_rnn, dim, learn_h0, q_levels, ulaw, weight_norm, cond_dim, spk_dim, qrnn=False):
        super().__init__()
        if qrnn:
            raise NotImplementedError("qrnn=True is broken in the reference (model.py:133-153,245-247); not supported")
        self.dim, self.q_levels, self.ulaw, self.cond_dim, self.spk_dim = dim, q_levels, ulaw, cond_dim, spk_dim
        self.n_rnn = n_rnn
        self.frame_sizes = [int(f) for f in frame_sizes]
        ns = [int(x) for x in np.cumprod(self.frame_sizes)]
        top = len(self.frame_sizes) - 1
        self.frame_level_rnns = tnn.ModuleList([
            FrameLevelRNN(fs, n, n_rnn, dim, learn_h0, i == top, cond_dim, spk_dim, weight_norm, qrnn)
            for i, (fs, n) in enumerate(zip(self.frame_sizes, ns))])
        self.sample_level_mlp = SampleLevelMLP(self.frame_sizes[0], dim, q_levels, weight_norm)
        # per-module forward calls (FrameLevelRNN.forward / SampleLevelMLP.forward / Runner.run_rnn) run on this model's
        # packed weights: weak back-references (plain attributes, not submodules) and the arithmetic mode they use
        self._bind_modules()
        self.module_mode = L.MODE_FP32
        self._ctx = None
        self._ctx_device = None
        self._packed_key = None

    def _bind_modules(self):
        """Register this model as the owner of its tiers / MLP (weak, outside the modules: state_dict, pickling and deepcopy
        of the modules are unaffected; a copied model re-registers when its context is created)."""
        for i, rnn in enumerate(self.frame_level_rnns):
            _OWNERS[id(rnn)] = (weakref.ref(self), i)
        _OWNERS[id(self.sample_level_mlp)] = (weakref.ref(self), -1)

    # -- reference API ---------------------------------------------------------------------------------------------
    @property
    def lookback(self):
        return self.frame_level_rnns[-1].n_frame_samples

    def dequantize(self, samples, q_levels=None):
        """API parity with utils.udequantize / linear_dequantize; not on the hot path (kernels use the LUT)."""
        x = samples.float()
        q = q_levels or self.q_levels
        if not self.ulaw:
            return x / (q / 2) - 1
        c = x * 2.0 / q - 1.0
        return torch.sign(c) * (torch.exp(torch.abs(c) * 5.5451774444795623) - 1) / 255.0

    # -- device context ------------------------------------------------------------------------------------------------
    def _device(self):
        dev = self.sample_level_mlp.embedding.weight.device
        if dev.type != "cuda":
            raise L.SrnnError("SampleRNN parameters are on %s: the B200 path has no CPU fallback, call .cuda()" % dev)
        return dev

    def _context(self):
        self._bind_modules()
        dev = self._device()
        if self._ctx is None or self._ctx_device != dev:
            self._release()
            lib = L.load()
            cfg = L.Config()
            cfg.n_tiers = len(self.frame_sizes)
            for i, f in enumerate(self.frame_sizes):
                cfg.frame_sizes[i] = f
            cfg.n_rnn, cfg.dim, cfg.q_levels = self.n_rnn, self.dim, self.q_levels
            cfg.cond_dim, cfg.spk_dim, cfg.ulaw = self.cond_dim, self.spk_dim, int(bool(self.ulaw))
            h = C.c_void_p()
            with torch.cuda.device(dev):
                L.check(lib.srnn_create(C.byref(cfg), C.byref(h)))
            self._ctx, self._ctx_device, self._packed_key = h, dev, None
        return self._ctx

    def _release(self):
        if getattr(self, "_ctx", None) is not None:
            try:
                L.load().srnn_destroy(self._ctx)
            except Exception:
                pass
            try:
                object.__setattr__(self, "_ctx", None)      # plain attribute; also safe during interpreter shutdown
            except Exception:
                pass

    def __del__(self):
        self._release()

    def _tensors(self):
        return list(self.parameters()) + list(self.buffers())

    def _c_params(self, ptr=None):
        """srnn_params over the parameter tensors (``ptr`` = tensor -> device pointer or None, see _Conv.c_params)."""
        ptr = ptr or (lambda t: t.data_ptr())
        P = L.Params()
        for i, rnn in enumerate(self.frame_level_rnns):
            tp = P.tiers[i]
            tp.h0 = ptr(rnn.h0)
            tp.input_expand = rnn.input_expand.c_params(ptr=ptr)
            if rnn.cond_expand is not None:
                tp.cond_expand = rnn.cond_expand.c_params(ptr=ptr)
                tp.spk_embedding = ptr(rnn.spk_embedding.weight)
                tp.spk_expand = rnn.spk_expand.c_params(ptr=ptr)
            for l in range(self.n_rnn):
                tp.weight_ih[l] = ptr(getattr(rnn.rnn, f"weight_ih_l{l}"))
                tp.weight_hh[l] = ptr(getattr(rnn.rnn, f"weight_hh_l{l}"))
                tp.bias_ih[l] = ptr(getattr(rnn.rnn, f"bias_ih_l{l}"))
                tp.bias_hh[l] = ptr(getattr(rnn.rnn, f"bias_hh_l{l}"))
            tp.upsampling = rnn.upsampling.conv_t.c_params(bias=rnn.upsampling.bias, ptr=ptr)
        mlp = self.sample_level_mlp
        P.embedding = ptr(mlp.embedding.weight)
        P.mlp_input, P.mlp_hidden, P.mlp_output = (mlp.input.c_params(ptr=ptr), mlp.hidden.c_params(ptr=ptr),
                                                   mlp.output.c_params(ptr=ptr))
        return P

    def _ensure_packed(self):
        """Re-snapshot the parameters into kernel layouts when any of them changed (in-place or re-allocated)."""
        ctx = self._context()
        ts = self._tensors()
        for t in ts:
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise L.SrnnError("parameters must be contiguous fp32")
        key = tuple((t.data_ptr(), t._version) for t in ts)
        if key == self._packed_key:
            return ctx
        P = self._c_params()
        with torch.cuda.device(self._ctx_device):
            L.check(L.load().srnn_pack_weights(ctx, C.byref(P), _stream()))
        self._packed_key = key
        return ctx


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ------------------------------------------------------------------------------------------------------------------
# Runner / Predictor / Generator  (model.py:328-520)
# ------------------------------------------------------------------------------------------------------------------
class Runner:
    def __init__(self, model):
        super().__init__()
        self.model = model
        self.reset_hidden_states()

    def reset_hidden_states(self):                                       # model.py:335-336
        self.hidden_states = {rnn: None for rnn in self.model.frame_level_rnns}

    def run_rnn(self, rnn, prev_samples, upper_tier_conditioning, cond, spk, writer=None, iterations=None):
        """model.py:338-349: one tier through ``FrameLevelRNN.forward`` with this runner's hidden state, which is replaced
        by the (detached) carry."""
        output, new_hidden = rnn(prev_samples, upper_tier_conditioning, self.hidden_states[rnn], cond, spk, writer, iterations)
        self.hidden_states[rnn] = new_hidden.detach()
        return output


class _PredictFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, input_sequences, cond, spk, mode, *params):
        model = runner.model
        h = model._ensure_packed()
        dev = model._ctx_device
        B, Lseq = input_sequences.shape
        T = Lseq - model.lookback + 1
        # the library reads raw pointers: every shape is checked here (the reference fails in its .view calls, model.py:406-408)
        if T < model.lookback or T % model.lookback:
            raise ValueError("input_sequences must be (B, lookback + T - 1) with T a positive multiple of lookback = %d"
                             % model.lookback)
        if tuple(cond.shape) != (B, T // model.lookback, model.cond_dim):
            raise ValueError("cond must be (B, T / lookback, cond_dim) = (%d, %d, %d), got %s"
                             % (B, T // model.lookback, model.cond_dim, tuple(cond.shape)))
        if spk.numel() != B:
            raise ValueError("spk must hold one speaker id per sequence")
        for name, t, hi in (("input_sequences", input_sequences, model.q_levels), ("spk", spk, model.spk_dim)):
            if not t.is_cuda and t.numel() and (int(t.min()) < 0 or int(t.max()) >= hi):   # host tensors: free to check
                raise ValueError("%s holds values outside [0, %d)" % (name, hi))
        seq = input_sequences.to(device=dev, dtype=torch.int64).contiguous()
        if cond.dtype not in (torch.float32, torch.float64):
            cond = cond.float()
        cond = cond.to(dev).contiguous()
        spk = spk.to(device=dev, dtype=torch.int64).reshape(B).contiguous()
        mask, ptrs = 0, (C.c_void_p * len(model.frame_level_rnns))()
        for i, rnn in enumerate(model.frame_level_rnns):
            hs = runner.hidden_states[rnn]
            if hs is not None and hs.shape[1] != B:                      # the reference's nn.GRU raises on a carried state
                raise ValueError("batch size changed from %d to %d with a carried hidden state: call "
                                 "reset_hidden_states() or pass reset=True" % (hs.shape[1], B))
            if hs is None:                                               # model.py:222-228 start from h0
                hs = torch.empty(model.n_rnn, B, model.dim, device=dev, dtype=torch.float32)
                mask |= 1 << i
            else:
                hs = hs.to(dev).contiguous().clone()
            runner.hidden_states[rnn] = hs                               # overwritten with the carry (model.py:348)
            ptrs[i] = hs.data_ptr()
        out = torch.empty(B, T, model.q_levels, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.check(L.load().srnn_predict_fwd(h, B, T, seq.data_ptr(), cond.data_ptr(),
                                              int(cond.dtype == torch.float64), spk.data_ptr(), ptrs, mask,
                                              out.data_ptr(), mode, _stream()))
        ctx.model, ctx.params, ctx.mode = model, params, mode
        ctx.keep = (seq, cond, spk)                # the library reads them again in the backward pass
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        """srnn_predict_bwd: gradients of every parameter that was passed to forward (trainer/__init__.py:103)."""
        model, params = ctx.model, ctx.params
        (logp,) = ctx.saved_tensors
        dev = model._ctx_device
        dlogp = grad_out.to(device=dev, dtype=torch.float32).contiguous()
        grads = _backward_into(model, params, lambda P, G: L.load().srnn_predict_bwd(
            model._ctx, logp.data_ptr(), dlogp.data_ptr(), C.byref(P), C.byref(G), _stream()))
        return (None, None, None, None, None) + grads


def _backward_into(model, params, call):
    """Run one library backward pass (``call(P, G)`` -> status) and return the per-parameter gradients for autograd.
    Gradient destinations: a fresh tensor per parameter (autograd then accumulates it into p.grad), unless the optimizer has
    published zeroed gradient views for this step (ClampAdam.zero_grad -> model._grad_sink): the library then writes
    straight into them and autograd gets None -- no 46 extra accumulate launches over 229 MB."""
    sink = getattr(model, "_grad_sink", None)
    model._grad_sink = None
    direct = sink is not None and all(id(p) in sink for p in params)
    model._grad_sink_used = direct
    grads = {id(p): (sink[id(p)] if direct else torch.empty_like(p)) for p in params}
    G = model._c_params(ptr=lambda t: grads[id(t)].data_ptr() if id(t) in grads else None)
    P = model._c_params()
    with torch.cuda.device(model._ctx_device):
        L.check(call(P, G))
    return tuple(None if direct else grads[id(p)] for p in params)


class _PredictNllFn(torch.autograd.Function):
    """``sequence_nll_loss_bits(Predictor.forward(...), target)`` as ONE autograd node over the parameters (nn.py:66-70 +
    trainer/__init__.py:102-103): forward = srnn_nll_loss_bits, backward = srnn_predict_bwd_nll, which folds the loss
    gradient into the log-softmax backward.  No dense (B, T, Q) dL/dlogp and no torch kernel in the step."""

    @staticmethod
    def forward(ctx, model, logp, target, *params):
        dev = model._ctx_device
        target = target.to(device=dev, dtype=torch.int64).contiguous()
        if target.numel() * logp.shape[-1] != logp.numel():
            raise ValueError("target must hold one class index per log-prob row")
        loss = torch.empty((), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.check(L.load().srnn_nll_loss_bits(model._ctx, logp.data_ptr(), target.data_ptr(), int(target.numel()),
                                                loss.data_ptr(), _stream()))
        ctx.model, ctx.params, ctx.logp, ctx.target = model, params, logp, target
        return loss

    @staticmethod
    def backward(ctx, g):
        model, params, logp, target = ctx.model, ctx.params, ctx.logp, ctx.target
        g = g.to(device=model._ctx_device, dtype=torch.float32).contiguous()
        grads = _backward_into(model, params, lambda P, G: L.load().srnn_predict_bwd_nll(
            model._ctx, logp.data_ptr(), target.data_ptr(), g.data_ptr(), C.byref(P), C.byref(G), _stream()))
        return (None, None, None) + grads


class Predictor(Runner, tnn.Module):
    """model.py:352-436.  ``mode``: 0 = fp32 parity arithmetic, 1 = bf16 tensor-core arithmetic."""

    def __init__(self, model, mode=L.MODE_FP32):
        super().__init__(model)
        self.mode = mode

    def forward(self, input_sequences, reset, cond, spk, writer=None, iterations=None):
        if reset:
            self.reset_hidden_states()                                   # model.py:358-359
        params = [p for p in self.model.parameters() if p.requires_grad] if torch.is_grad_enabled() else []
        out = _PredictFn.apply(self, input_sequences, cond, spk, self.mode, *params)
        out._srnn_rec = (self.model, tuple(params))    # lets sequence_nll_loss_bits fuse the loss into the backward pass
        return out


class Generator(Runner):
    """model.py:439-520.  The whole autoregressive loop runs on the device; no per-sample host round trip."""

    def __init__(self, model, cuda=False, mode=L.MODE_FP32):
        super().__init__(model)
        self.cuda = cuda
        self.mode = mode

    @torch.no_grad()
    def __call__(self, n_seqs, seq_len, cond, spk, uniforms=None, seed=None, return_samples=False,
                 return_logp=False, device_output=False):
        """Reference form: ``cond`` numpy (n_cond, cond_dim) and python-int ``spk`` shared by all ``n_seqs``
        sequences; ``seq_len`` is ignored exactly as in model.py:455.  Extension: ``cond`` (n_seqs, n_cond, cond_dim)
        and ``spk`` (n_seqs,) per utterance; ``uniforms`` (n_cond*lookback, n_seqs) pre-drawn U[0,1) for the defined
        sampler (drawn on the device from ``seed`` when omitted).  Returns the dequantised audio
        (n_seqs, n_cond*lookback) float32 on the CPU like the reference (on the device with ``device_output``)."""
        model = self.model
        h = model._ensure_packed()
        dev = model._ctx_device
        cond = torch.as_tensor(np.asarray(cond) if not torch.is_tensor(cond) else cond)
        if cond.dim() == 2:
            cond = cond.unsqueeze(0)
        cond_rows, n_cond, cond_dim = cond.shape
        if cond_dim != model.cond_dim:
            raise ValueError("cond has width %d, model expects %d" % (cond_dim, model.cond_dim))
        if cond_rows not in (1, n_seqs):
            raise ValueError("cond must be (n_cond, cond_dim) or (n_seqs, n_cond, cond_dim)")
        cond = cond.to(device=dev, dtype=torch.float32).contiguous()
        spk = torch.as_tensor(np.asarray(spk) if not torch.is_tensor(spk) else spk).reshape(-1)
        if spk.numel() != cond_rows:
            if spk.numel() == 1 and cond_rows == n_seqs:
                spk = spk.expand(n_seqs)
            elif cond_rows == 1 and spk.numel() == n_seqs:
                cond, cond_rows = cond.expand(n_seqs, -1, -1).contiguous(), n_seqs
            else:
                raise ValueError("spk must hold one id per conditioner row")
        if not spk.is_cuda and spk.numel() and (int(spk.min()) < 0 or int(spk.max()) >= model.spk_dim):
            raise ValueError("spk holds ids outside [0, %d)" % model.spk_dim)
        spk = spk.to(device=dev, dtype=torch.int64).contiguous()
        T = n_cond * model.lookback                                      # model.py:455
        if uniforms is None:
            if seed is not None:
                g = torch.Generator(device=dev)
                g.manual_seed(int(seed))
                uniforms = torch.rand(T, n_seqs, device=dev, dtype=torch.float32, generator=g)
            else:       # consume the global CUDA stream like the reference's multinomial (model.py:517): no re-seeding
                uniforms = torch.rand(T, n_seqs, device=dev, dtype=torch.float32)
        else:
            uniforms = torch.as_tensor(uniforms).to(device=dev, dtype=torch.float32).contiguous()
            if tuple(uniforms.shape) != (T, n_seqs):
                raise ValueError("uniforms must be (n_cond*lookback, n_seqs) = (%d, %d)" % (T, n_seqs))
        samples = torch.empty(n_seqs, T, device=dev, dtype=torch.uint8)
        audio = torch.empty(n_seqs, T, device=dev, dtype=torch.float32)
        logp = torch.empty(n_seqs, T, model.q_levels, device=dev, dtype=torch.float32) if return_logp else None
        with torch.cuda.device(dev):
            L.check(L.load().srnn_generate(h, n_seqs, n_cond, cond.data_ptr(), cond_rows, spk.data_ptr(),
                                           uniforms.data_ptr(), samples.data_ptr(), audio.data_ptr(),
                                           logp.data_ptr() if logp is not None else None, self.mode, _stream()))
        out = audio if device_output else audio.cpu()
        if not (return_samples or return_logp):
            return out
        res = [out]
        if return_samples:
            res.append(samples if device_output else samples.cpu())
        if return_logp:
            res.append(logp if device_output else logp.cpu())
        return tuple(res)
