#!/usr/bin/env python
"""Benchmark of the PFN training hot path on B200 (contract: see task statement / DESIGN.md section "Measurement").

    python bench.py --gpus 1 --steps 6 --warmup 3                 # this repo's CUDA engine (default)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # data parallel, one rank per GPU
    python bench.py --impl reference --steps 3 --warmup 1          # the reference's CPU path (port) on the host cores

One "step" = one full training step on one batch of synthetic prior data:
    GP prior draw (fused sampler kernel) -> embed -> 6 x {QKV GEMM, masked attention, out-proj, LN, GELU-MLP, LN}
    -> decoder on the query rows -> FullSupportBarDistribution NLL -> backward -> [NCCL grad all-reduce] -> clip -> Adam.
Metric (BASELINE.json): prior-sampled sequences / second.  Workload = configs[1]: priors.fast_gp, seq_len 1000,
1 feature, emsize 512, 6 layers, nhid 1024, 4 heads, 100 bars, single_eval_pos 500, bf16, batch 512 per GPU.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG2 = dict(T=1000, F=1, E=512, H=4, nhid=1024, L=6, n_bars=100, sep=500, batch=512,
            hps={"noise": 1e-4, "outputscale": 1., "lengthscale": .6, "fast_computations": (False, False, False)})


def step_flops(T, B, F, E, nhid, L, n_out, sep):
    """Algorithmic (mask-aware) FLOPs of one training step, SURVEY.md section 8d."""
    dense = T * B * L * (8 * E * E + 4 * E * nhid)
    attn = 4 * E * B * L * (T * sep + (T - sep))
    dec = (T - sep) * B * (2 * E * nhid + 2 * nhid * n_out)
    enc = T * B * 2 * F * E + sep * B * 2 * E
    return 3 * (dense + attn + dec + enc)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16_sustained=p.get("bf16_tflops_sustained"), bf16_burst=p.get("bf16_tflops"), hbm=p.get("hbm_gbs"),
                    source="MEASURED_PEAKS.json (measured)")
    return dict(bf16_sustained=1400.0, bf16_burst=1590.0, hbm=6650.0, source="B200_PROFILING.md fallback")


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path (oracle port, kind 'port'), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_reference_step as C
    cfg = CFG2
    threads = cpu_threads()
    torch.set_num_threads(threads)
    sample_b = args.ref_batch
    borders = torch.linspace(-4.0, 4.0, cfg["n_bars"] + 1)
    step, _ = C.make_step(cfg["T"], cfg["F"], cfg["E"], cfg["H"], cfg["nhid"], cfg["L"], cfg["n_bars"], cfg["sep"], sample_b,
                          cfg["hps"], borders, threads)
    dt = C.time_steps(step, args.steps, args.warmup)
    value = args.steps * sample_b / dt
    sample = (f"{args.steps} timed steps of the cfg2 shape at batch {sample_b} (per-sequence cost is batch-invariant); "
              f"torch {torch.__version__} CPU, {threads} threads")
    line = {"impl": "reference", "metric": "prior-sampled sequences/sec, full training step (cfg2 shape)", "value": value,
            "unit": "seq/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (priors.fast_gp draw restated on CPU)",
            "config": {"workload": workload_name(cfg, sample_b), "bounded_sample_batch": sample_b},
            "cpu_baseline": {"value": value, "unit": "seq/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_threads():
    """Threads for the CPU arm: torch's intra-op pool scales poorly past a few dozen threads on this small per-step
    problem (128 threads measured 40x slower than 8), so use at most 32 and report the number actually used."""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get("PFN_CPU_THREADS", "32"))))


def workload_name(cfg, batch):
    return (f"cfg2: priors.fast_gp T={cfg['T']} F={cfg['F']} hps(noise 1e-4, os 1, ls .6), emsize {cfg['E']}, "
            f"{cfg['L']} layers, nhid {cfg['nhid']}, {cfg['H']} heads, {cfg['n_bars']} bars FullSupport, "
            f"single_eval_pos {cfg['sep']}, batch {batch}/GPU")


def run_engine(args):
    from transformerscandobayesianinference_b200 import _lib as L, bar_distribution, encoders, parallel, priors, transformer
    import torch.distributed as dist

    rank, world, dev = parallel.init_from_env("cuda")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the hot path has no CPU fallback)"
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    cfg = dict(CFG2)
    B = args.batch or cfg["batch"]
    T, F, E, H, nhid, Lyr, n_bars, sep = (cfg[k] for k in ("T", "F", "E", "H", "nhid", "L", "n_bars", "sep"))
    torch.manual_seed(1234 + rank)
    peaks = load_peaks()

    # ---- model, criterion, optimizer (the objects train.train builds)
    torch.manual_seed(1234)
    model = transformer.TransformerModel(encoders.Linear(F, E), n_bars, E, H, nhid, Lyr, 0.0, y_encoder=encoders.Linear(1, E)).to(dev)
    model.precision = args.precision
    parallel.broadcast_parameters(model)
    torch.manual_seed(1234 + rank)
    ys = priors.fast_gp.get_batch(64, T, F, device=str(dev), hyperparameters=cfg["hps"])[1]
    with contextlib.redirect_stdout(sys.stderr):      # the reference-style helper prints; stdout carries the JSON line only
        borders = bar_distribution.get_bucket_limits(n_bars, ys=ys.float().cpu())
    crit = bar_distribution.FullSupportBarDistribution(borders).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    params = [p for p in model.parameters()]
    ls = torch.full((B, F), float(cfg["hps"]["lengthscale"]), device=dev)
    os_ = torch.full((B,), float(cfg["hps"]["outputscale"]), device=dev)
    nz = torch.full((B,), float(cfg["hps"]["noise"]), device=dev)

    def train_step(x_bt=None, z_bt=None):
        """x_bt [B,T,F], z_bt [B,T] on the device (sampled here when None)."""
        if x_bt is None:
            x_bt = torch.rand(B, T, F, device=dev)
            z_bt = torch.randn(B, T, device=dev)
        y_bt = priors.fast_gp.sample_gp(x_bt, z_bt, ls, os_, nz)
        x, y = x_bt.transpose(0, 1), y_bt.transpose(0, 1)
        logits = model((x, y), single_eval_pos=sep)
        loss = crit(logits.reshape(-1, n_bars), y[sep:].flatten()).mean()
        loss.backward()
        parallel.allreduce_gradients(params)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    bar_distribution.BarDistribution.defer_support_check = True
    for _ in range(args.warmup):
        train_step()
    # ---- device-resident throughput (value)
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    L.reset_launch_count()
    L.PROFILE_GEMM = [] if rank == 0 else None
    ms = timed(train_step, args.steps)
    gemm_prof = L.PROFILE_GEMM
    L.PROFILE_GEMM = None
    launches = L.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    value = args.steps * B * world / (ms / 1e3)

    # ---- end to end through the public API with HOST inputs: pinned x ~ U[0,1), z ~ N(0,1) copied H2D every step,
    #      loss read back D2H every step (what a user-side data pipeline + logging would do)
    n_host = min(args.steps, 4)
    hx = [torch.rand(B, T, F).pin_memory() for _ in range(n_host)]
    hz = [torch.randn(B, T).pin_memory() for _ in range(n_host)]
    counter = [0]

    def e2e_step():
        i = counter[0] % n_host
        counter[0] += 1
        x_bt = hx[i].to(dev, non_blocking=True)
        z_bt = hz[i].to(dev, non_blocking=True)
        loss = train_step(x_bt, z_bt)
        return loss.item()
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = args.steps * B * world / (ms_e2e / 1e3)
    h2d = hx[0].numel() * 4 + hz[0].numel() * 4

    if rank != 0:
        return
    flops = step_flops(T, B, F, E, nhid, Lyr, n_bars, sep)
    achieved_step = flops * args.steps / (ms / 1e3) / 1e12
    # dominant kernel: the tcgen05 GEMM (all dense-layer launches of the timed region, CUDA events on the launch stream)
    g_flops = sum(r[0] for r in gemm_prof)
    g_ms = sum(r[1].elapsed_time(r[2]) for r in gemm_prof)
    gemm_tf = g_flops / (g_ms / 1e3) / 1e12 if g_ms > 0 else None
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture of this command (profiles/, not measured live)
    traffic, traffic_src, traffic_alg = None, None, None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_gemm_traffic.json")
    if os.path.exists(tpath) and B == CFG2["batch"] and args.precision == "bf16":
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["mean_dram_bytes_per_launch"], "profiles/r1_gemm_traffic.json"
        traffic_alg = tj.get("mean_algorithmic_bytes_per_launch")
    g_bytes = sum(r[3] for r in gemm_prof)
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 GEMM, all launches of the timed steps)",
                "achieved": gemm_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": (gemm_tf / peaks["bf16_sustained"]) if gemm_tf else None,
                "traffic": traffic, "traffic_unit": "bytes/launch (dram read+write, mean over the captured launches)",
                "traffic_source": traffic_src, "traffic_algorithmic": traffic_alg,   # same 8 captured launches: operands + results once
                "algorithmic_bytes_per_launch": g_bytes / max(len(gemm_prof), 1),   # mean over ALL launches of the timed steps
                "launches": len(gemm_prof), "kernel_ms_per_step": g_ms / args.steps, "peak_source": peaks["source"] + ", sustained bf16",
                "step": {"achieved": achieved_step, "frac": achieved_step / peaks["bf16_sustained"], "flops_per_step": flops}}

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0 only
    cpu_baseline = None
    if not args.no_cpu_baseline:
        from oracle import cpu_reference_step as C
        threads = cpu_threads()
        cb = args.ref_batch
        stepc, _ = C.make_step(T, F, E, H, nhid, Lyr, n_bars, sep, cb, cfg["hps"], borders, threads)
        dtc = C.time_steps(stepc, 3, 1)
        cpu_baseline = {"value": 3 * cb / dtc, "unit": "seq/s", "cores": threads, "kind": "port",
                        "sample": f"3 timed steps (+1 warm-up) of the same workload at batch {cb}, torch {torch.__version__} CPU fp32"}

    line = {"metric": "prior-sampled sequences/sec, full training step (sample+fwd+bwd+allreduce+Adam)", "value": value,
            "unit": "seq/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic (priors.fast_gp draws, random-init weights)",
            "config": {"workload": workload_name(cfg, B), "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2_policy": "inputs larger than L2 (activations ~0.5 GB per tensor)", "precision": args.precision},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "seq/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4},
            "roofline": roofline, "cpu_baseline": cpu_baseline}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's 512)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--ref-batch", type=int, default=4, help="bounded CPU sample: sequences per CPU step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
