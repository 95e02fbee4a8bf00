"""Host mirror of the reference training-step semantics: ``optim.gradient_clipping(torch.optim.Adam(...))``
(optim.py:4-21, train.py:238-241) as ONE fused CUDA launch (srnn_clamp_adam_step) over all parameter tensors.

    opt = ClampAdam(predictor.parameters(), lr=1e-3)
    opt.zero_grad(); loss = opt.step(closure)          # closure = forward + loss + backward, as trainer/__init__.py:99-112

``zero_grad`` keeps zero tensors (torch-0.4 semantics the reference relies on: a parameter that received no gradient,
e.g. ``h0`` on a non-reset batch, still gets an Adam update from its momentum).
With ``torch.distributed`` initialised, ``step`` averages the gradients over ranks (one all-reduce on a flat bucket) BEFORE
the clamp, so an N-GPU run equals a single-GPU run at N times the batch.
"""
import ctypes as C

import torch

from . import _lib as L


class ClampAdam:
    """All gradients live in ONE flat fp32 bucket (``p.grad`` are views into it): ``zero_grad`` is a single memset, and the
    data-parallel exchange is a single in-place NCCL all-reduce over NVLink with no gather/scatter copies."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, clamp=1.0, process_group=None, model=None):
        self.params = [p for p in params if p.requires_grad]
        # Bucket order = the order in which srnn_predict_bwd finalises gradients (srnn_bwd_wait_stage): sample-level MLP +
        # embedding, then per tier (lowest first) its upsampling and the rest.  Each stage is all-reduced on a side stream as
        # soon as its event fires, i.e. while the backward pass of the following stages still runs.
        self._stages = []                   # [(stage id, first element, one-past-last element)]
        self._stage_params = []             # [(first parameter index, one-past-last)] per stage, same order as _stages
        if model is not None:
            groups = [(0, list(model.sample_level_mlp.parameters()))]
            for i, rnn in enumerate(model.frame_level_rnns):
                up = list(rnn.upsampling.parameters())
                ids = {id(p) for p in up}
                groups.append((1 + 2 * i, up))
                groups.append((2 + 2 * i, [p for p in rnn.parameters() if id(p) not in ids]))
            mine = {id(p) for p in self.params}
            ordered, seen, off = [], set(), 0
            for sid, ps in groups:
                ps = [p for p in ps if id(p) in mine and id(p) not in seen]
                seen.update(id(p) for p in ps)
                n = sum(p.numel() for p in ps)
                if n:
                    self._stages.append((sid, off, off + n))
                    self._stage_params.append((len(ordered), len(ordered) + len(ps)))
                ordered += ps
                off += n
            rest = [p for p in self.params if id(p) not in seen]       # parameters the backward pass does not know about
            if rest:
                self._stages = []
            self.params = ordered + rest
        self._comm_stream = None
        self.lr, self.betas, self.eps, self.clamp = lr, betas, eps, clamp
        self.step_count = 0
        self.exp_avg = self.exp_avg_sq = None          # allocated on the parameters' device at the first step
        self.process_group = process_group
        self.model = model                  # the SampleRNN whose packed weights must be refreshed after an update
        self.param_groups = [{"params": self.params, "lr": lr}]       # enough for torch LR schedulers' read access
        self._flat = None
        self._views = None
        self._grad_scale = 1.0

    def _bucket(self):
        """(Re)build the flat gradient bucket and the Adam moments when the parameters moved (e.g. ``.cuda()`` after
        construction); the moments follow the parameters' device."""
        p0 = self.params[0]
        if self._flat is None or self._flat.device != p0.device:
            self._flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=p0.device)
            self._views, off = [], 0
            for p in self.params:
                n = p.numel()
                self._views.append(self._flat[off:off + n].view_as(p))
                off += n
        if self.exp_avg is None:
            self.exp_avg = [torch.zeros_like(p) for p in self.params]
            self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        elif self.exp_avg[0].device != p0.device:
            self.exp_avg = [t.to(p0.device) for t in self.exp_avg]
            self.exp_avg_sq = [t.to(p0.device) for t in self.exp_avg_sq]
        return self._flat

    def _adopt(self):
        """Make every ``p.grad`` the bucket view (autograd assigns a fresh tensor when ``p.grad`` was None)."""
        self._bucket()
        for p, v in zip(self.params, self._views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            p.grad = v

    def zero_grad(self):
        self._bucket().zero_()
        for p, v in zip(self.params, self._views):
            p.grad = v
        if self.model is not None:      # one backward pass may write its gradients directly into the (zeroed) views
            self.model._grad_sink = {id(p): v for p, v in zip(self.params, self._views)}

    def _allreduce(self, overlap=False, staged_update=False):
        import torch.distributed as dist
        self._grad_scale = 1.0
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.process_group) == 1:
            return
        flat = self._flat
        if overlap and flat.is_cuda and len(self._stages) > 1:
            # one NCCL all-reduce per backward stage over NVLink, enqueued on a side stream behind that stage's event: the
            # MLP bucket is reduced while the tiers' backward passes run, tier i's while tier i+1's runs, ...
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=flat.device)
            comm, lib, works = self._comm_stream, L.load(), []
            for sid, a, b in self._stages:
                L.check(lib.srnn_bwd_wait_stage(self.model._ctx, sid, C.c_void_p(comm.cuda_stream)))
                with torch.cuda.stream(comm):
                    works.append(dist.all_reduce(flat[a:b], group=self.process_group, async_op=True))
            if staged_update:
                # the caller waits for each stage's collective right before that stage's clamp+Adam launch, so the update of
                # the early stages runs while the last stage's gradients are still on the wire
                self._grad_scale = 1.0 / dist.get_world_size(self.process_group)
                return works
            for w in works:
                w.wait()                                               # the compute stream waits for the collectives
        else:
            dist.all_reduce(flat, group=self.process_group)           # sum over ranks, in place, one collective
        # the mean BEFORE the clamp (optim.py:10-13 clamps the full-batch gradient) is folded into k_clamp_adam
        self._grad_scale = 1.0 / dist.get_world_size(self.process_group)

    def step(self, closure=None):
        loss = closure() if closure is not None else None
        # overlap is only valid when the backward pass wrote straight into the bucket (zero_grad published the sink and
        # the pass consumed it): otherwise autograd's accumulation into p.grad happens after the early event
        direct = self.model is not None and getattr(self.model, "_grad_sink_used", False)
        if self.model is not None:
            self.model._grad_sink_used = False
        self._adopt()
        works = self._allreduce(overlap=direct, staged_update=direct)
        self.step_count += 1
        lr = self.param_groups[0]["lr"]
        dev = self.params[0].device
        for t in list(self.params) + [p.grad for p in self.params] + self.exp_avg + self.exp_avg_sq:
            if t.device != dev or not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise L.SrnnError("ClampAdam: every parameter, gradient and moment must be contiguous fp32 on one CUDA device")

        def launch(lo, hi):
            n = hi - lo
            arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts[lo:hi]])
            sizes = (C.c_int64 * n)(*[p.numel() for p in self.params[lo:hi]])
            with torch.cuda.device(dev):
                L.check(L.load().srnn_clamp_adam_step_scaled(
                    n, arr([p.data for p in self.params]), arr([p.grad for p in self.params]), arr(self.exp_avg),
                    arr(self.exp_avg_sq), sizes, lr, self.betas[0], self.betas[1], self.eps, self.step_count, self.clamp,
                    self._grad_scale, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

        if works:                                   # data parallel, staged: one clamp+Adam launch per backward stage
            for w, (lo, hi) in zip(works, self._stage_params):
                w.wait()                            # the compute stream waits for THIS stage's collective only
                launch(lo, hi)
        else:
            launch(0, len(self.params))
        # the library updated the parameters behind torch's back: make SampleRNN re-pack them on the next forward
        if self.model is not None:
            self.model._packed_key = None
        else:
            with torch.no_grad():
                for p in self.params:
                    p.add_(0)                       # bumps the tensor version that SampleRNN._ensure_packed watches
        return loss


def gradient_clipping(optimizer, min=-1, max=1, model=None):
    """optim.py:4-21 ``gradient_clipping(torch.optim.Adam(params))`` (train.py:238-241).  The reference wraps the optimizer's
    ``step`` so that every gradient is clamped element-wise before the update; here clamp and Adam are one kernel, so a
    ``torch.optim.Adam`` is converted into the equivalent ``ClampAdam`` (same parameters, lr, betas, eps; fresh moments) and a
    ``ClampAdam`` is returned with its bound set.  Only the symmetric bound the reference uses is supported."""
    if -min != max:
        raise ValueError("gradient_clipping: symmetric bounds only (the reference clamps to [-1, 1])")
    if isinstance(optimizer, ClampAdam):
        optimizer.clamp = float(max)
        return optimizer
    if not isinstance(optimizer, torch.optim.Adam):
        raise L.SrnnError("gradient_clipping: only Adam has a fused CUDA step (the reference trains with Adam, train.py:238)")
    if len(optimizer.param_groups) != 1:
        raise L.SrnnError("gradient_clipping: one parameter group expected")
    g = optimizer.param_groups[0]
    if g.get("weight_decay", 0) or g.get("amsgrad", False):
        raise L.SrnnError("gradient_clipping: weight_decay / amsgrad are not part of the reference step")
    return ClampAdam(g["params"], lr=g["lr"], betas=tuple(g["betas"]), eps=g["eps"], clamp=float(max), model=model)


def sequence_nll_loss_bits(logp, target, model=None):
    """nn.py:66-70: ``nll_loss(logp.view(-1, Q), target.view(-1)) * log2(e)``, mean over B*T, in bits.

    ``logp`` straight from ``Predictor.forward`` (the training closure, trainer/__init__.py:100-103): the loss is the CUDA
    reduction ``srnn_nll_loss_bits`` and its backward is ``srnn_predict_bwd_nll`` -- the constant loss gradient is folded
    into the log-softmax backward kernel, which writes ``dlogits`` (fp32 + bf16) where the output-layer GEMMs read them, so
    neither a dense dL/dlogp nor any torch kernel runs in the step.  Any other CUDA fp32 ``logp`` needs ``model=`` (the
    SampleRNN whose context owns the reduction scratch); its backward is the plain scatter.  No CPU path."""
    from . import model as M
    rec = getattr(logp, "_srnn_rec", None)
    if rec is not None and logp.grad_fn is not None:
        mdl, params = rec
        return M._PredictNllFn.apply(mdl, logp.detach(), target, *params)
    mdl = rec[0] if rec is not None else model
    if mdl is None or not logp.is_cuda:
        raise L.SrnnError("sequence_nll_loss_bits runs on the CUDA library only: pass Predictor output or model=")
    return _NllBitsFn.apply(mdl, logp, target)


class _NllBitsFn(torch.autograd.Function):
    """Generic form for log-probs that did not come from Predictor.forward: fused forward reduction, scatter backward."""

    @staticmethod
    def forward(ctx, mdl, logp, target):
        lp = logp.detach().to(torch.float32).contiguous()
        tgt = target.to(device=lp.device, dtype=torch.int64).contiguous()
        loss = torch.empty((), device=lp.device, dtype=torch.float32)
        with torch.cuda.device(lp.device):
            L.check(L.load().srnn_nll_loss_bits(mdl._context(), lp.data_ptr(), tgt.data_ptr(), int(tgt.numel()),
                                                loss.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        ctx.shape, ctx.tgt = logp.shape, tgt
        return loss

    @staticmethod
    def backward(ctx, g):
        Q = ctx.shape[-1]
        d = torch.zeros(ctx.tgt.numel(), Q, device=ctx.tgt.device, dtype=torch.float32)
        d.scatter_(1, ctx.tgt.reshape(-1, 1), (-g * (1.4426950408889634 / ctx.tgt.numel())).expand(ctx.tgt.numel(), 1))
        return None, d.reshape(ctx.shape), None
