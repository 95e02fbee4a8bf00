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
        # bucket order: everything below the top tier first (its gradients are final before the top tier's backward pass
        # runs, srnn_bwd_wait_early), the top tier's parameters last
        self._n_early = 0
        if model is not None:
            late = {id(p) for p in model.frame_level_rnns[-1].parameters()}
            early = [p for p in self.params if id(p) not in late]
            self.params = early + [p for p in self.params if id(p) in late]
            self._n_early = sum(p.numel() for p in early)
        self._comm_stream = None
        self.lr, self.betas, self.eps, self.clamp = lr, betas, eps, clamp
        self.step_count = 0
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.process_group = process_group
        self.model = model                  # the SampleRNN whose packed weights must be refreshed after an update
        self.param_groups = [{"params": self.params, "lr": lr}]       # enough for torch LR schedulers' read access
        self._flat = None
        self._views = None

    def _bucket(self):
        """(Re)build the flat gradient bucket when the parameters moved (e.g. ``.cuda()`` after construction)."""
        p0 = self.params[0]
        if self._flat is None or self._flat.device != p0.device:
            self._flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=p0.device)
            self._views, off = [], 0
            for p in self.params:
                n = p.numel()
                self._views.append(self._flat[off:off + n].view_as(p))
                off += n
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

    def _allreduce(self, overlap=False):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.process_group) == 1:
            return
        flat, n0 = self._flat, self._n_early
        if overlap and flat.is_cuda and 0 < n0 < flat.numel():
            # two buckets over NVLink: the early one is reduced on a side stream as soon as the library signals that those
            # gradients are final, i.e. while the top tier's backward pass still runs on the compute stream
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=flat.device)
            comm = self._comm_stream
            L.check(L.load().srnn_bwd_wait_early(self.model._ctx, C.c_void_p(comm.cuda_stream)))
            with torch.cuda.stream(comm):
                work = dist.all_reduce(flat[:n0], group=self.process_group, async_op=True)
            dist.all_reduce(flat[n0:], group=self.process_group)
            work.wait()                                                # the compute stream waits for the early bucket
        else:
            dist.all_reduce(flat, group=self.process_group)           # sum over ranks, in place, one collective
        flat.mul_(1.0 / dist.get_world_size(self.process_group))      # mean BEFORE the clamp (optim.py:10-13 clamps the full-batch gradient)

    def step(self, closure=None):
        loss = closure() if closure is not None else None
        # overlap is only valid when the backward pass wrote straight into the bucket (zero_grad published the sink and
        # the pass consumed it): otherwise autograd's accumulation into p.grad happens after the early event
        direct = self.model is not None and getattr(self.model, "_grad_sink_used", False)
        if self.model is not None:
            self.model._grad_sink_used = False
        self._adopt()
        self._allreduce(overlap=direct)
        self.step_count += 1
        n = len(self.params)
        arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
        sizes = (C.c_int64 * n)(*[p.numel() for p in self.params])
        lr = self.param_groups[0]["lr"]
        dev = self.params[0].device
        with torch.cuda.device(dev):
            L.check(L.load().srnn_clamp_adam_step(n, arr([p.data for p in self.params]), arr([p.grad for p in self.params]),
                                                  arr(self.exp_avg), arr(self.exp_avg_sq), sizes, lr, self.betas[0],
                                                  self.betas[1], self.eps, self.step_count, self.clamp,
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        # the library updated the parameters behind torch's back: make SampleRNN re-pack them on the next forward
        if self.model is not None:
            self.model._packed_key = None
        else:
            with torch.no_grad():
                for p in self.params:
                    p.add_(0)                       # bumps the tensor version that SampleRNN._ensure_packed watches
        return loss


def sequence_nll_loss_bits(logp, target, model=None):
    """nn.py:66-70 through the fused CUDA reduction when no gradient is needed; autograd-friendly torch expression
    otherwise (the gradient of the loss w.r.t. the log-probs is a constant scatter, there is nothing to fuse)."""
    import math
    Q = logp.shape[-1]
    picked = logp.reshape(-1, Q).gather(1, target.reshape(-1, 1).long().to(logp.device))
    return -picked.mean() * math.log(math.e, 2)
