"""Data-parallel training exchange: one all-reduce of the flat gradient bucket, mean over ranks BEFORE the
element-wise clamp, so that an N-rank run equals a single-rank run at N times the batch.

CPU (gloo, world_size 2): the bucket / all-reduce host logic of ClampAdam.  GPU (NCCL, needs >= 2 devices, skipped
otherwise): two ranks with half the batch each reproduce the single-GPU parameters after three steps."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import srnn_b200 as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    ps = [torch.zeros(n).requires_grad_(True) for n in (3, 1000, 17)]
    opt = S.ClampAdam(ps, lr=1e-3)
    opt.zero_grad()
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(ps, opt._views))
    for i, p in enumerate(ps):                      # autograd-style accumulation into the views; one grad replaced outright
        p.grad += float(rank + 1) * (i + 1)
    ps[2].grad = torch.full_like(ps[2], 5.0 * (rank + 1))
    opt._adopt()
    opt._allreduce()
    if rank == 0:
        out.put([float(p.grad.mean()) * opt._grad_scale for p in ps] + [float(opt._flat.numel())])   # the mean is folded into the Adam kernel
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == [1.5, 3.0, 7.5, 1020.0]            # mean over ranks of (1, 2)*k and of (5, 10)


C = dict(frame_sizes=[20, 4], n_rnn=2, dim=64, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=86,
         spk_dim=6)


def _train(model_sd, x, y, cond, spk, steps, dev, mode):
    m = S.SampleRNN(**C)
    p = S.Predictor(m, mode=mode)
    p.load_state_dict(model_sd)
    p.to(dev)
    opt = S.ClampAdam(p.parameters(), lr=1e-3, model=m)
    T = 80
    for i in range(steps):
        xs, ys = x[:, i * T: i * T + 80 + T - 1].contiguous().to(dev), y[:, i * T: (i + 1) * T].contiguous().to(dev)
        cs = cond[:, i: i + 1].contiguous().to(dev)

        def closure():
            out = p(xs, i == 0, cs, spk.to(dev), None, None)
            loss = S.sequence_nll_loss_bits(out, ys)
            loss.backward()
            return loss.detach()

        opt.zero_grad()
        opt.step(closure)
    return {k: v.detach().cpu() for k, v in p.state_dict().items()}


def _inputs():
    g = torch.Generator().manual_seed(5)
    B, steps, T = 4, 3, 80
    data = torch.randint(0, 256, (B, 80 + steps * T), generator=g)
    x, y = data[:, :-1], data[:, 80:]
    cond = torch.rand(B, steps, 86, generator=g)
    spk = torch.randint(0, 6, (B, 1), generator=g)
    torch.manual_seed(11)
    sd = {k: v.clone() for k, v in S.Predictor(S.SampleRNN(**C)).state_dict().items()}
    return sd, x, y, cond, spk, steps


def _nccl_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    sd, x, y, cond, spk, steps = _inputs()
    lo, hi = S.shard_range(x.shape[0], rank, world)
    res = _train(sd, x[lo:hi], y[lo:hi], cond[lo:hi], spk[lo:hi], steps, dev, S.MODE_FP32)
    if rank == 0:
        out.put({k: v.numpy() for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_training_equals_single_gpu_double_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    sd, x, y, cond, spk, steps = _inputs()
    ref = _train(sd, x, y, cond, spk, steps, torch.device("cuda", 0), S.MODE_FP32)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    for k, v in ref.items():
        d = np.abs(got[k] - v.numpy())
        # equal up to the summation order of the batch mean; Adam may flip a few elements with |g| ~ eps by ~lr
        assert d.max() <= 3.1e-3, (k, d.max())
        assert (d > 1e-4).mean() <= 2e-3, (k, (d > 1e-4).mean())
