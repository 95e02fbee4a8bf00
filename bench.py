#!/usr/bin/env python
"""Headline benchmark: autoregressive generation throughput of the 3-tier speaker-conditioned SampleRNN
(BASELINE.json configs[1] = "C2": frame_sizes [20,4], 2 GRU layers, dim 1024, q 256, look-ahead cond 86,
weight-norm, 6 speakers), batch 256 utterances PER GPU, random-init weights, synthetic conditioners.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode fp32|bf16]

One "step" = one pass of the hot path over one batch: generate `--n-cond` conditioner frames (x80 samples) for 256
utterances.  `value` = generated samples/s over all GPUs with inputs resident in HBM (CUDA events around
srnn_generate); `e2e` = the same through the public `Generator.__call__` with HOST (pinned) inputs and the audio
read back to the host inside the timed region.  Utterances shard across ranks with no collective ("weak").
`--impl reference` times the reference algorithm's CPU port (oracle/) on the host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C2 = dict(frame_sizes=[20, 4], n_rnn=2, dim=1024, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True,
          cond_dim=86, spk_dim=6)
F_ALG = 6.423e6      # FLOP per generated sample per utterance, embedding-o-conv folded; what the kernels execute
F_DENSE = 16.889e6   # the reference graph's dense work, for context
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the sample-level kernel at B = 256 from the `ncu --set full`
# captures summarised in profiles/ (cold L2: ncu flushes caches between replays)
NCU_TRAFFIC = {"k_mlp_persist": 34.14e6 + 0.03e6,      # profiles/r1_k_mlp_persist_ncu_full.txt
               "k_mlp_cluster": 36.35e6 + 0.002e6}     # profiles/r2_k_mlp_cluster_ncu_full.txt (second captured launch)
F_MLP = 2.0 * (1024 * 1024 + 1024 * 256) + 20 * 1024   # k_mlp_persist's share of F_ALG per sample per utterance: hidden +
                                                       # output contraction + the folded-table adds (SURVEY 8d components)
SAMPLE_RATE = 16000


def measured_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def synth_inputs(B, n_cond, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    cond = torch.rand(B, n_cond, C2["cond_dim"], generator=g)                     # features are min-max scaled to [0,1]
    spk = torch.randint(0, C2["spk_dim"], (B,), generator=g)
    uni = torch.rand(n_cond * 80, B, generator=g)
    return cond, spk, uni


C1 = dict(frame_sizes=[16], n_rnn=1, dim=1024, learn_h0=True, q_levels=256, ulaw=True, weight_norm=False,
          cond_dim=43, spk_dim=6)      # BASELINE.json configs[0] / BASELINE.md 3: the reference's own CPU-runnable case


def host_threads():
    """All host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to every rank)."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except Exception:
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_port_rate(n_cond, B, cfg=None, shared_cond=False):
    """The oracle (CPU port of the reference algorithm, oracle/srnn_oracle.py) on a bounded sample of the workload."""
    import torch
    from oracle import srnn_oracle as O
    import srnn_b200 as S
    cfg = cfg or C2
    threads = host_threads()
    torch.manual_seed(77977)
    m = S.SampleRNN(**cfg)
    sd = {"model." + k: v.detach() for k, v in m.state_dict().items()}
    w = O.unpack_state_dict(sd, O.Config(**cfg))
    g = torch.Generator().manual_seed(0)
    lookback = m.lookback
    cond = torch.rand(B, n_cond, cfg["cond_dim"], generator=g)
    spk = torch.randint(0, cfg["spk_dim"], (B,), generator=g)
    uni = torch.rand(n_cond * lookback, B, generator=g)
    gen = O.Generator(w)
    t0 = time.perf_counter()
    gen(B, cond.numpy(), spk.numpy(), uni.numpy())
    dt = time.perf_counter() - t0
    return B * n_cond * lookback / dt, dt, threads


def workload_config(args, world):
    """The `config` object both arms print: BASELINE.json configs[1] (C2) at the batch / length of this run."""
    return {"workload": "C2 generation: 3-tier [20,4] SampleRNN, n_rnn 2, dim 1024, q 256, cond %d, weight-norm" % C2["cond_dim"],
            "batch_per_gpu": args.batch, "total_batch": args.batch * world, "samples_per_utterance": args.n_cond * 80}


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path = the oracle port (the reference itself is a
    Python program under /root/reference, which does not exist on the GPU box), all host threads, each step a bounded
    sample of the arm's workload (same model, same batch, fewer conditioner frames) sized so that the whole
    --steps/--warmup run stays within --ref-budget-s seconds."""
    if rank != 0:
        return
    B = args.batch
    # calibrate on one conditioner frame, then size the per-step sample for the time budget
    r1, dt1, th = cpu_port_rate(1, B)
    per_step = args.ref_budget_s / max(1, args.steps + args.warmup)
    n_cond = max(1, min(args.ref_n_cond, int(per_step / max(dt1, 1e-3))))
    rates, times = [], []
    for i in range(args.warmup + args.steps):
        r, dt, th = cpu_port_rate(n_cond, B)
        if i >= args.warmup:
            rates.append(r)
            times.append(dt)
    v = sum(rates) / len(rates)
    sample = ("bounded sample per step: %d of the workload's %d conditioner frames (%d samples) for all %d utterances; "
              "oracle port, %d host threads" % (n_cond, args.n_cond, n_cond * 80, B, th))
    print(json.dumps({
        "impl": "reference", "metric": "generated samples/sec", "value": v, "unit": "samples/s",
        "x_realtime_16k": v / SAMPLE_RATE, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": th, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def train_block(args, rank, world, dev, steps, warmup):
    """BASELINE.json configs[2] ("C3"): teacher-forced training step = forward + NLL(bits) + backward + staged gradient
    all-reduce over NVLink (N>1) + fused mean/clamp/Adam, batch 128 x T 1040 per GPU, bf16, synthetic tokens/conditioners.
    Returns the `train` record of the bench line (process group already initialised by the caller when world > 1)."""
    import torch
    import torch.distributed as dist
    import srnn_b200 as S
    mode = S.MODE_FP32 if args.mode == "fp32" else S.MODE_BF16
    lib = S._lib.load()
    torch.manual_seed(77977)
    model = S.SampleRNN(**C2).to(dev)
    pred = S.Predictor(model, mode=mode)
    opt = S.ClampAdam(pred.parameters(), lr=1e-4, model=model)
    B, T = args.train_batch, args.train_T
    n_it = steps + warmup
    g = torch.Generator().manual_seed(100 + rank)                      # every rank trains on its own rows
    data = torch.randint(0, 256, (B, 80 + T * n_it + T), generator=g).to(dev)
    cond = torch.rand(B, (T // 80) * (n_it + 1) + 1, C2["cond_dim"], generator=g).to(dev)
    spk = torch.randint(0, 6, (B, 1), generator=g).to(dev)
    losses = []

    def step(i):
        s = i * T
        x = data[:, s: s + 80 + T - 1].contiguous()
        y = data[:, s + 80: s + 80 + T].contiguous()
        c = cond[:, i * (T // 80) + 1: (i + 1) * (T // 80) + 1].contiguous()

        def closure():
            out = pred(x, i == 0, c, spk, None, None)
            loss = S.sequence_nll_loss_bits(out, y)                    # fused: srnn_nll_loss_bits / srnn_predict_bwd_nll
            loss.backward()
            return loss.detach()

        opt.zero_grad()
        losses.append(opt.step(closure))

    for i in range(warmup):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = lib.srnn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]      # per-step stamps: the median exposes a transient
    e0.record()
    marks[0].record()
    for k, i in enumerate(range(warmup, n_it)):
        step(i)
        marks[k + 1].record()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    per_step = sorted(marks[k].elapsed_time(marks[k + 1]) for k in range(steps))
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    # data-parallel equivalence inside the bench run: the parameters must stay bit-identical on every rank
    chk = torch.stack([p.detach().double().sum() for p in pred.parameters()]).sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool((lo == hi).item())
    assert same, "parameters diverged across ranks: %r vs %r" % (float(lo), float(hi))
    secs = float(t.item()) / 1e3
    tokens = world * B * T * steps
    peak_tf, _, peak_src = measured_peaks()
    f_step = 3 * F_ALG                                          # fwd + bwd ~ 3x forward, folded form (what the kernels execute)
    ach = tokens / secs * f_step / 1e12 / world
    launches = int(lib.srnn_launch_count() - l0)
    del opt, pred, model, data, cond
    torch.cuda.empty_cache()
    return {
        "metric": "training tokens/sec (teacher-forced step)", "value": tokens / secs, "unit": "tokens/s", "n_gpus": world,
        "tokens_per_s_per_gpu": tokens / secs / world, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * secs / steps,
        "ms_per_step_median_rank0": per_step[len(per_step) // 2], "ms_per_step_max_rank0": per_step[-1],
        "higher_is_better": True, "scaling": "weak", "dtype": "bf16" if mode == S.MODE_BF16 else "f32", "data": "synthetic",
        "config": {"workload": "C3 training step: 3-tier [20,4] SampleRNN dim 1024, batch %d x T %d per GPU, Adam lr 1e-4 "
                               "with element-wise gradient clamp" % (B, T),
                   "allreduce": ("fp32 sum over ranks in backward-stage buckets on a side stream (NCCL, NCCL_MAX_CTAS=%s); mean folded "
                                 "into the clamp+Adam kernel" % os.environ.get("NCCL_MAX_CTAS", "default"))
                                if world > 1 else "none (1 GPU)",
                   "loss": "fused (srnn_nll_loss_bits + srnn_predict_bwd_nll), no torch kernel in the step"},
        "loss_bits_first_last": [float(losses[0]), float(losses[-1])],
        "params_identical_across_ranks": same,
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                     "traffic": None, "peak_source": peak_src + " bf16_tflops_sustained",
                     "flops_per_token": f_step, "kernel": "whole training step"},
    }


def fp32_mode_block(S, model, B, dev, n_cond=2):
    """Throughput of the fp32 parity mode (the mode the 1e-3 / bit-exact-index gates are proven on), same model and batch."""
    import torch
    gen = S.Generator(model, cuda=True, mode=S.MODE_FP32)
    cond, spk, uni = [t.to(dev) for t in synth_inputs(B, n_cond, 7)]
    gen(B, 0, cond, spk, uniforms=uni, device_output=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gen(B, 0, cond, spk, uniforms=uni, device_output=True)
    e1.record()
    torch.cuda.synchronize()
    v = B * n_cond * 80 / (e0.elapsed_time(e1) / 1e3)
    return {"value": v, "unit": "samples/s", "x_realtime_16k": v / SAMPLE_RATE, "samples_per_utterance": n_cond * 80,
            "note": "SRNN_MODE_FP32: every contraction as an fp32 FFMA GEMM; the mode of the 1e-3 logit / bit-exact index gates"}


def x3_mode_block(S, model, B, dev, n_cond=2):
    """The tensor-core parity mode (SRNN_MODE_BF16X3: fp32 control flow, dense contractions as split-bf16 tcgen05 products):
    throughput on the same model and batch, and its agreement with the fp32 mode on the same uniforms -- max relative log-prob
    difference |dlogp| / max(1, |logp|) over the steps both runs shared a history for (up to the first differing sample of
    each utterance) and the fraction of identical sampled indices."""
    import torch
    cond, spk, uni = [t.to(dev) for t in synth_inputs(B, n_cond, 7)]
    out = {}
    for name, mode in (("fp32", S.MODE_FP32), ("x3", S.MODE_BF16X3)):
        gen = S.Generator(model, cuda=True, mode=mode)
        gen(B, 0, cond, spk, uniforms=uni, device_output=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gen(B, 0, cond, spk, uniforms=uni, device_output=True)
        e1.record()
        torch.cuda.synchronize()
        out[name + "_ms"] = e0.elapsed_time(e1)
        _, smp, lp = gen(B, 0, cond, spk, uniforms=uni, device_output=True, return_samples=True, return_logp=True)
        out[name] = (smp.long(), lp)
    (s0, l0), (s1, l1) = out["fp32"], out["x3"]
    same = (s0 == s1)
    shared = torch.cumprod(torch.cat([torch.ones_like(same[:, :1]), same[:, :-1]], 1).long(), 1).bool()   # history identical so far
    rel = ((l0 - l1).abs() / l0.abs().clamp(min=1.0)).amax(-1)
    v = B * n_cond * 80 / (out["x3_ms"] / 1e3)
    return {"value": v, "unit": "samples/s", "x_realtime_16k": v / SAMPLE_RATE, "samples_per_utterance": n_cond * 80,
            "max_rel_dlogp": float(rel[shared].max()), "index_match": float(same.float().mean()),
            "index_match_shared_history": float(same[shared].float().mean()),
            "speedup_vs_fp32_mode": out["fp32_ms"] / out["x3_ms"],
            "note": "SRNN_MODE_BF16X3 against SRNN_MODE_FP32 on the same uniforms: W.x as Wh.xh + Wl.xh + Wh.xl on tcgen05 "
                    "(one GEMM over K' = 3K, fp32 accumulation); gate 1e-3 relative on log-probs"}


def run_sweep(args, rank, world, local):
    """BASELINE.json configs[3] ("C4"): total batch 1 ... 4096 utterances x 10 s (160 000 samples), utterances sharded over
    the ranks (strong scaling over the batch, no collective); one JSON line per batch size: latency of the first period
    (80 samples), latency of the whole utterance, throughput."""
    import torch
    import torch.distributed as dist
    import srnn_b200 as S
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mode = S.MODE_FP32 if args.mode == "fp32" else S.MODE_BF16
    torch.manual_seed(77977)
    model = S.SampleRNN(**C2).to(dev)
    gen = S.Generator(model, cuda=True, mode=mode)
    n_cond = args.sweep_seconds * SAMPLE_RATE // 80
    for total in [int(b) for b in args.sweep_batches.split(",")]:
        lo, hi = S.shard_range(total, rank, world)
        B = hi - lo
        t_first = t_all = 0.0
        if B > 0:
            cond, spk, uni = [t.to(dev) for t in synth_inputs(B, n_cond, 2000 + rank)]
            gen(B, 0, cond[:, :1].contiguous(), spk, uniforms=uni[:80].contiguous(), device_output=True)   # warm-up
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            gen(B, 0, cond[:, :1].contiguous(), spk, uniforms=uni[:80].contiguous(), device_output=True)
            ev[1].record()
            ev[2].record()
            gen(B, 0, cond, spk, uniforms=uni, device_output=True)
            ev[3].record()
            torch.cuda.synchronize()
            t_first, t_all = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])
            del cond, spk, uni
        t = torch.tensor([t_first, t_all], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            secs = float(t[1]) / 1e3
            v = total * n_cond * 80 / secs
            print(json.dumps({"workload": "C4 batch sweep", "total_batch": total, "n_gpus": world, "batch_per_gpu": -(-total // world),
                              "samples_per_utterance": n_cond * 80, "latency_first_period_ms": float(t[0]),
                              "latency_last_sample_ms": float(t[1]), "value": v, "unit": "samples/s",
                              "x_realtime_16k": v / SAMPLE_RATE, "us_per_sample_step": 1e6 * secs / (n_cond * 80),
                              "dtype": "bf16" if mode == S.MODE_BF16 else "f32"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("SRNN_BENCH_MODE", "auto"), choices=["auto", "fp32", "bf16", "bf16_graph"])
    ap.add_argument("--batch", type=int, default=256, help="utterances per GPU")
    ap.add_argument("--n-cond", type=int, default=100, help="conditioner frames (x80 samples) per step")
    ap.add_argument("--ref-n-cond", type=int, default=8, help="reference arm: at most this many cond frames per step")
    ap.add_argument("--ref-budget-s", type=float, default=120.0, help="reference arm: CPU seconds for the whole run")
    ap.add_argument("--cpu-n-cond", type=int, default=24, help="cpu_baseline sample: cond frames (~12 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="generate", choices=["generate", "train", "sweep"],
                    help="generate = the headline metric + a `train` sub-record (default); train = only the C3 training "
                         "step; sweep = C4 batch sweep (one line per batch size)")
    ap.add_argument("--ind-cond-dim", type=int, default=0,
                    help="> 0: BASELINE.json configs[4], the bottle-neck variant (run_sampleneck.sh: 30): k=1 conditioner chain "
                         "43 -> 40 -> 30 -> 20 -> ind_cond_dim in front of the top tier (thesis-derived, parity unpinned); implies --cond-dim 43")
    ap.add_argument("--cond-dim", type=int, default=86,
                    help="conditioner width: 86 = look-ahead (C2, default); 43 = the core of the bottle-neck variant (C5)")
    ap.add_argument("--train-batch", type=int, default=128)
    ap.add_argument("--train-T", type=int, default=1040)
    ap.add_argument("--train-steps", type=int, default=10)
    ap.add_argument("--no-train", action="store_true", help="skip the `train` sub-record of the default run")
    ap.add_argument("--sweep-batches", default="1,4,16,64,256,1024,4096")
    ap.add_argument("--sweep-seconds", type=int, default=10)
    args = ap.parse_args()
    if args.ind_cond_dim > 0:
        args.cond_dim = 43
    C2["cond_dim"] = args.cond_dim

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    # The gradient all-reduce runs beside the backward pass, whose persistent GRU kernels need their 128 CTAs co-resident on
    # the 148 SMs: NCCL is held to 16 CTAs (measured at 8 GPUs, ms per C3 step: 8 -> 8.83, 16 -> 8.66, 32 -> 8.86, default 8.83)
    if world > 1:
        os.environ.setdefault("NCCL_MAX_CTAS", "16")
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.workload == "sweep":
        return run_sweep(args, rank, world, local)

    import torch
    import torch.distributed as dist
    import srnn_b200 as S
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.workload == "train":
        rec = train_block(args, rank, world, dev, args.steps, args.warmup)
        if rank == 0:
            rec["vs_baseline"] = None
            print(json.dumps(rec))
        if world > 1:
            dist.destroy_process_group()
        return
    mode = {"fp32": S.MODE_FP32, "bf16": S.MODE_BF16, "bf16_graph": S.MODE_BF16_GRAPH}.get(args.mode)
    if mode is None:
        mode = S.MODE_BF16 if getattr(S.package, "HAS_BF16", False) else S.MODE_FP32
    lib = S._lib.load()

    torch.manual_seed(77977)                                   # train.py:62; same weights on every rank
    if args.ind_cond_dim > 0:
        full = S.BottleneckSampleRNN(ind_cond_dim=args.ind_cond_dim, **C2).to(dev)
        model, gen = full.core, S.BottleneckGenerator(full, cuda=True, mode=mode)
    else:
        model = S.SampleRNN(**C2).to(dev)
        gen = S.Generator(model, cuda=True, mode=mode)
    B, n_cond = args.batch, args.n_cond
    T = n_cond * 80
    cond_h, spk_h, uni_h = [t.pin_memory() for t in synth_inputs(B, n_cond, 1000 + rank)]   # each rank: its own utterances
    cond_d, spk_d, uni_d = cond_h.to(dev), spk_h.to(dev), uni_h.to(dev)
    audio_h = torch.empty(B, T, dtype=torch.float32).pin_memory()
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return gen(B, 0, cond_d, spk_d, uniforms=uni_d, device_output=True)

    def step_e2e():
        a = gen(B, 0, cond_h.to(dev, non_blocking=True), spk_h.to(dev, non_blocking=True),
                uniforms=uni_h.to(dev, non_blocking=True), device_output=True)
        audio_h.copy_(a, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return audio_h

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        l0 = lib.srnn_launch_count()
        t0 = time.perf_counter()
        for a, b in evs:
            flush.fill_(1)                                     # L2 flush between timed iterations (untimed)
            a.record()
            fn()
            b.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)           # max over ranks
        return float(t.item()) / 1e3, lib.srnn_launch_count() - l0, wall

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    secs, launches, _ = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    secs_e2e, _, _ = timed(step_e2e, args.steps, max(1, args.warmup // 3))

    # Dominant kernel (the persistent sample-level kernel, one launch per tier-0 frame = FS0 samples x B utterances): CUDA
    # events around every launch on the library's generation stream, over one more (untimed-for-`value`) pass with direct launches.
    kern = None
    if mode == S.MODE_BF16:
        import ctypes
        os.environ["SRNN_TIME_KERNELS"] = "1"
        try:
            flush.fill_(1)
            step_device()
            torch.cuda.synchronize()
            ms, n = ctypes.c_double(0), ctypes.c_int64(0)
            lib.srnn_timed_kernel(model._ctx, ctypes.byref(ms), ctypes.byref(n))
            if n.value:
                kern = (ms.value / n.value * 1e-3, n.value, ms.value * 1e-3)
        finally:
            del os.environ["SRNN_TIME_KERNELS"]

    total = world * B * T * args.steps
    value = total / secs
    e2e = total / secs_e2e
    peak_tf, peak_gbs, peak_src = measured_peaks()
    ach_step = value * F_ALG / 1e12 / world                                     # per-GPU TFLOP/s, algorithmic, whole step
    fs0 = C2["frame_sizes"][0]
    kname = lib.srnn_sample_kernel_name().decode()
    if kern:
        flops_launch = F_MLP * B * fs0
        ach = flops_launch / kern[0] / 1e12
        traffic = NCU_TRAFFIC.get(kname) if B == 256 else None
        kinfo = {"kernel": "srnn::" + kname, "launches_timed": kern[1], "avg_launch_us": kern[0] * 1e6, "flops_per_launch": flops_launch,
                 "share_of_step": kern[2] / (secs / args.steps),
                 "how": "CUDA events around each launch on the generation stream (direct launches, SRNN_TIME_KERNELS)"}
    else:
        traffic = None
        ach, kinfo = ach_step, {"kernel": "whole generation step (all launches of srnn_generate)"}
    h2d = cond_h.numel() * 4 + spk_h.numel() * 8 + uni_h.numel() * 4
    d2h = audio_h.numel() * 4
    cfg = workload_config(args, world)
    cfg.update({"l2": "192 MiB flush write between timed iterations", "us_per_sample_step": 1e6 * secs / args.steps / T,
                "schedule": "default (shadow recurrent GEMMs, programmatic dependent launches incl. the sample kernel, input expansion "
                            "folded into the first GRU layer with its known columns summed in the shadow, table sums carried "
                            "across sample launches: fp32 sums re-associated / different bf16 rounding points, inside the bf16 gate)"})
    line = {
        "metric": "generated samples/sec", "value": value, "unit": "samples/s", "x_realtime_16k": value / SAMPLE_RATE,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if mode == S.MODE_FP32 else "bf16", "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e, "unit": "samples/s", "x_realtime_16k": e2e / SAMPLE_RATE, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                     "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write, cold L2)",
                     "peak_source": peak_src + " bf16_tflops_sustained",
                     "flops_per_sample": F_ALG, "whole_step_achieved": ach_step, "whole_step_frac": ach_step / peak_tf,
                     "achieved_dense_equiv": value * F_DENSE / 1e12 / world, **kinfo},
    }
    if args.ind_cond_dim > 0:
        line["config"]["workload"] = ("C5 generation: bottle-neck variant, conditioner chain 43 -> 40 -> 30 -> 20 -> %d -> 1024 in front of "
                                      "the 3-tier [20,4] SampleRNN dim 1024 (thesis-derived chain, parity unpinned)" % args.ind_cond_dim)
    elif mode != S.MODE_FP32:
        line["fp32_parity_mode"] = fp32_mode_block(S, model, B, dev)
        line["parity_mode"] = x3_mode_block(S, model, B, dev)
    del gen, flush, cond_d, uni_d
    torch.cuda.empty_cache()
    if not args.no_train and args.ind_cond_dim == 0:
        line["train"] = train_block(args, rank, world, dev, args.train_steps, 3)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r, dt, th = cpu_port_rate(args.cpu_n_cond, B)
        line["cpu_baseline"] = {"value": r, "unit": "samples/s", "cores": th, "kind": "port",
                                "sample": "C2 model, B=%d x %d cond frames (%.1f s of CPU work)" % (B, args.cpu_n_cond, dt)}
        # the reference's own CPU case -- C1, one utterance, shared conditioner; bounded to 250 of its 1000 frames
        r1, dt1, _ = cpu_port_rate(250, 1, cfg=C1)
        line["cpu_baseline"]["c1"] = {"value": r1, "unit": "samples/s", "x_realtime_16k": r1 / SAMPLE_RATE, "cores": th, "kind": "port",
                                      "sample": "C1: 2-tier [16] SampleRNN dim 1024 cond 43, B=1, 250 of 1000 cond frames "
                                                "(4000 samples, %.1f s of CPU work)" % dt1}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
