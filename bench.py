#!/usr/bin/env python
"""Benchmark of the DGMR GAN training step on the B200-native path.

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torch.distributed.run)
    python bench.py --impl reference ...                    (CPU arm: the oracle port of the reference on host cores)

One "step" = one full GAN step (2 discriminator updates + 1 generator update, hinge + grid-cell losses, Adam) on a
synthetic batch of 4->18-frame 256x256 radar sequences: BASELINE.json configs[2] per GPU (configs[3] when N > 1:
batch 16 per GPU, weak scaling, NCCL all-reduce of the flat G/D gradient buffers).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_G = 521.4e9   # forward FLOPs (2*MAC) of the generator per sample, paper config
F_D = 35.7e9    # forward FLOPs of both discriminators per 22-frame sequence


def flop_step(batch, k):
    """Parity-preserving minimal GAN-step FLOPs: B*[2*(F_G + 6 F_D) + K*(3 F_G + 3 F_D)]."""
    return batch * (2 * (F_G + 6 * F_D) + k * (3 * F_G + 3 * F_D))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d.get("bf16_tflops", 1590.0), bf16_sustained=d.get("bf16_tflops_sustained", 1400.0),
                    hbm=d.get("hbm_gbs", 6650.0), source="measured")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        smax = max([float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()] + [0.0])
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=smax or None, reasons=sorted(reasons), samples=len(self.rows))


# ----------------------------------------------------------------------------------------------------- CPU arm
def build_oracle_state(cfg, seed=0):
    """Seeded construction through the package's parameter containers (CPU, no kernels), as a flat state dict."""
    import skillful_nowcasting_b200 as B

    torch.manual_seed(seed)
    s = cfg["output_shape"]
    gen = B.Generator(B.ContextConditioningStack(input_channels=1, output_channels=cfg["context_channels"]),
                      B.LatentConditioningStack(shape=(8, s // 32, s // 32), output_channels=cfg["latent_channels"]),
                      B.Sampler(forecast_steps=cfg["forecast_steps"], latent_channels=cfg["latent_channels"],
                                context_channels=cfg["context_channels"]))
    disc = B.Discriminator(input_channels=1)
    return gen, disc


def usable_cores() -> int:
    """Host cores this process may really use: affinity mask and cgroup quota, not os.cpu_count() (on a box with a CPU quota,
    one thread per visible core oversubscribes the quota and the oracle crawls)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, min(n, 32))     # the oracle's many small ops do not scale past a few tens of threads


# bounded samples of the workload, largest first: (image side, forecast steps).  The discriminator needs side >= 128.
CPU_SAMPLES = ((128, 18), (128, 6), (128, 2))


def _cpu_sample_child(argv):
    """Child process: time `steps` oracle GAN steps of one bounded sample and print one JSON line."""
    side, t, batch, steps, k, lat, ctx, threads = (int(v) for v in argv)
    from oracle import dgmr_oracle as O

    torch.set_num_threads(threads)
    cfg = dict(output_shape=side, forecast_steps=t, latent_channels=lat, context_channels=ctx)
    gen, disc = build_oracle_state(cfg)
    gs = O.clone_state(gen.state_dict(), requires_grad=True)
    ds = O.clone_state(disc.state_dict(), requires_grad=True)
    g_opt = O.AdamState([gs[n] for n in O._trainable(gs)], lr=5e-5)
    d_opt = O.AdamState([ds[n] for n in O._trainable(ds)], lr=2e-4)
    torch.manual_seed(1234)
    x, y = torch.rand(batch, 4, 1, side, side), torch.rand(batch, t, 1, side, side)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.gan_step(gs, ds, g_opt, d_opt, x, y, t, (8, side // 32, side // 32), generation_steps=k)
        times.append(time.perf_counter() - t0)
        print(json.dumps(dict(times=times)), flush=True)      # partial results survive a timeout


def cpu_reference_steps(cfg, batch, steps, warmup, k, budget_s=150.0):
    """The reference's own CPU implementation of the path = the oracle port (oracle/dgmr_oracle.py: torch fp32 on the host
    cores), timed on a BOUNDED sample and scaled to the metric's unit.  A sample is the same GAN step (same widths, same
    schedule) on smaller frames / fewer lead times; its time is scaled by the pixel-and-frame ratio to the full
    256x256 4->18 step (the step is convolution-dominated, cost ~ pixels x frames).  Runs in a child process under a wall-clock
    budget so that a slow host can never stall the benchmark: on timeout the next smaller sample is tried."""
    import subprocess

    threads = usable_cores()
    full_side, full_t = cfg["output_shape"], cfg["forecast_steps"]
    n_steps = max(1, steps + warmup)
    last_err = "no sample finished"
    for side, t in CPU_SAMPLES:
        side, t = min(side, full_side), min(t, full_t)
        cmd = [sys.executable, os.path.abspath(__file__), "--cpu-sample-child", str(side), str(t), str(batch), str(n_steps), str(k),
               str(cfg["latent_channels"]), str(cfg["context_channels"]), str(threads)]
        t_start = time.perf_counter()
        out = ""
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=budget_s).stdout
        except subprocess.TimeoutExpired as e:
            out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            last_err = f"sample {side}x{side} 4->{t} exceeded {budget_s:.0f} s"
        lines = [ln for ln in out.splitlines() if ln.startswith("{")]
        times = json.loads(lines[-1])["times"] if lines else []
        budget_s = max(30.0, budget_s - (time.perf_counter() - t_start))
        if len(times) > warmup or (times and len(times) == n_steps):
            timed = times[warmup:] if len(times) > warmup else times
            mean = sum(timed) / len(timed)
            scale = (full_side * full_side * full_t) / float(side * side * t)
            full_step_s = mean * scale
            return dict(value=batch * full_t / full_step_s, unit="frames/s", cores=threads, kind="port", s_per_step=full_step_s,
                        sample=f"oracle port, full GAN step (same widths and schedule) on {side}x{side} frames, 4->{t} lead times, batch {batch}: "
                               f"{len(timed)} timed step(s) of {mean:.2f} s, scaled x{scale:.1f} (pixels x frames) to the 256x256 4->18 step")
    return dict(value=None, unit="frames/s", cores=threads, kind="port", s_per_step=float("nan"), sample=f"unavailable: {last_err}")


def gpu_reference_steps(cfg, batch, steps, warmup, k, dev):
    """The reference's GPU path for the '>= 5x cuDNN-backed step' target: the same oracle port, tensors on cuda
    (PyTorch eager, cuDNN convs with torch's default allow_tf32=True), minimal schedule like the B200 arm."""
    from oracle import dgmr_oracle as O

    gen, disc = build_oracle_state(cfg)
    gs = {k_: v.to(dev) for k_, v in O.clone_state(gen.state_dict()).items()}
    ds = {k_: v.to(dev) for k_, v in O.clone_state(disc.state_dict()).items()}
    for st in (gs, ds):
        for k_, v in st.items():
            if v.is_floating_point() and not (k_.endswith("._u") or k_.endswith("._v") or "running_" in k_):
                v.requires_grad_(True)
    g_opt = O.AdamState([gs[n] for n in O._trainable(gs)], lr=5e-5)
    d_opt = O.AdamState([ds[n] for n in O._trainable(ds)], lr=2e-4)
    s = cfg["output_shape"]
    x = torch.rand(batch, 4, 1, s, s, device=dev)
    y = torch.rand(batch, cfg["forecast_steps"], 1, s, s, device=dev)
    try:
        for _ in range(warmup):
            O.gan_step(gs, ds, g_opt, d_opt, x, y, cfg["forecast_steps"], (8, s // 32, s // 32), generation_steps=k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            O.gan_step(gs, ds, g_opt, d_opt, x, y, cfg["forecast_steps"], (8, s // 32, s // 32), generation_steps=k)
        e1.record()
        torch.cuda.synchronize()
    except torch.cuda.OutOfMemoryError:
        return dict(batch=batch, error="out of memory")
    ms = e0.elapsed_time(e1) / steps
    return dict(batch=batch, ms_per_step=ms, value=batch * cfg["forecast_steps"] / (ms * 1e-3), unit="frames/s",
                what="oracle port on cuda (PyTorch eager + cuDNN, allow_tf32 default), minimal schedule")


# ----------------------------------------------------------------------------------------------------- GPU arm
_REAL_STDOUT = None


def emit(line: dict):
    """Write the ONE JSON line of the contract to the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # stdout must carry exactly one JSON line, but libraries print there too (NCCL's "NCCL version ..." banner under torchrun):
    # point fd 1 at stderr for the duration of the run and keep the real stdout for emit()
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch (BASELINE configs[2]/[3]: 16)")
    ap.add_argument("--generation-steps", type=int, default=1)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--forecast-steps", type=int, default=18)
    ap.add_argument("--latent-channels", type=int, default=768)
    ap.add_argument("--context-channels", type=int, default=384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=1)
    ap.add_argument("--ref-gpu", type=int, default=0, metavar="BATCH",
                    help="also time the reference's own GPU path (the oracle port on cuda: PyTorch eager + cuDNN, TF32 convs "
                         "allowed as in torch's default) at this batch; reported as `reference_gpu_eager`")
    args = ap.parse_args()
    cfg = dict(output_shape=args.size, forecast_steps=args.forecast_steps, latent_channels=args.latent_channels,
               context_channels=args.context_channels)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K = args.generation_steps
    workload = (f"DGMR full GAN step (G + spatial+temporal D, hinge + grid-cell, 2 D updates + 1 G update, Adam), 4->{args.forecast_steps} "
                f"frames 1x{args.size}x{args.size}, latent {args.latent_channels} / context {args.context_channels}, batch {args.batch}/GPU, "
                f"generation_steps={K}")

    if args.impl == "reference":
        if rank != 0:
            return
        warm = min(args.warmup, 1)
        r = cpu_reference_steps(cfg, args.cpu_batch, max(1, min(args.steps, 3)), warm, K, budget_s=240.0)
        line = dict(impl="reference", metric="radar frames/sec (G+D step, 256x256, 4->18)", value=r["value"], unit="frames/s",
                    n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=(r["s_per_step"] * 1e3 if r["value"] else None), higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=workload, note="CPU arm runs the oracle port on a bounded sample"),
                    cpu_baseline=dict(value=r["value"], unit="frames/s", cores=r["cores"], kind="port", sample=r["sample"]),
                    e2e=dict(value=r["value"], unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        emit(line)
        return

    import torch.distributed as dist
    from skillful_nowcasting_b200 import _lib, ops
    from skillful_nowcasting_b200.training import Adam, gan_step

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=dev)
    be = _lib.backend()  # raises if libdgmr_b200.so is missing: no fallback
    gen, disc = build_oracle_state(cfg, seed=0)  # identical replicas on every rank
    gen.to(dev).train()
    disc.to(dev).train()
    g_opt = Adam(gen.parameters(), lr=5e-5, betas=(0.0, 0.999))
    d_opt = Adam(disc.parameters(), lr=2e-4, betas=(0.0, 0.999))
    B, T, S = args.batch, args.forecast_steps, args.size
    torch.manual_seed(1234 + rank)
    host_x = torch.rand(B, 4, 1, S, S).pin_memory()
    host_y = torch.rand(B, T, 1, S, S).pin_memory()
    x, y = host_x.to(dev), host_y.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return gan_step(gen, disc, g_opt, d_opt, x, y, generation_steps=K)

    def step_e2e():
        xi = host_x.to(dev, non_blocking=True)
        yi = host_y.to(dev, non_blocking=True)
        out = gan_step(gen, disc, g_opt, d_opt, xi, yi, generation_steps=K)
        return torch.stack([out["d_loss"], out["g_loss"], out["grid_loss"]]).cpu()  # device->host read of the step's result

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = be.launches
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), be.launches - l0

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    # ---- dominant-kernel roofline: one extra instrumented step, CUDA events around every conv launch
    be.profile = []
    step_resident()
    torch.cuda.synchronize()
    prof, be.profile = be.profile, None
    agg = {}
    shapes = {}
    for name, flops, ev0, ev1, tag, info in prof:
        ms_ = ev0.elapsed_time(ev1)
        a = agg.setdefault(tag, [0.0, 0.0, 0])
        a[0] += flops
        a[1] += ms_
        a[2] += 1
        b_ = shapes.setdefault((tag, info), [0.0, 0.0, 0])
        b_[0] += flops
        b_[1] += ms_
        b_[2] += 1
    dump = os.environ.get("DGMR_BENCH_DUMP")
    if dump and rank == 0:
        with open(dump, "w") as f:
            f.write("tag\tshape\tcalls\tms\tTFLOP/s\n")
            for (tag, info), (fl, ms_, n_) in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{tag}\t{info}\t{n_}\t{ms_:.3f}\t{(fl / (ms_ * 1e-3) / 1e12 if ms_ > 0 else 0):.1f}\n")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    ms_per_step = ms / args.steps
    frames = world * B * T
    value = frames / (ms_per_step * 1e-3)
    e2e_value = frames / (ms_e2e / args.steps * 1e-3)
    tf32_peak = pk["bf16_sustained"] / 2.0
    dom = agg.get("conv_umma", [0.0, 0.0, 0])
    achieved = dom[0] / (dom[1] * 1e-3) / 1e12 if dom[1] > 0 else 0.0
    line = dict(
        metric="radar frames/sec (G+D step, 256x256, 4->18)", value=value, unit="frames/s", n_gpus=world, steps=args.steps,
        warmup=max(args.warmup, 3), ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="tf32 (fp32 storage, tcgen05 kind::tf32 operands, fp32 accumulate)", data="synthetic",
        config=dict(workload=workload, global_batch=world * B, parallelism=f"dp{world}",
                    l2="inputs+activations per step (>10 GB) exceed the 126 MB L2; no explicit flush needed",
                    schedule="parity-preserving minimal schedule (SURVEY.md 8d)"),
        e2e=dict(value=e2e_value, unit="frames/s", h2d_bytes_per_step=int(host_x.numel() + host_y.numel()) * 4, d2h_bytes_per_step=12),
        gpu_launches=launches // args.steps,
        clocks=clocks,
        step_tflops=flop_step(world * B, K) / (ms_per_step * 1e-3) / 1e12,
        roofline=dict(bound="tensor", kernel="conv_umma_patch_kernel + conv_umma_fwd[_persist]_kernel (implicit-GEMM conv fwd + dgrad, tcgen05 kind::tf32)",
                      achieved=achieved, peak=tf32_peak, unit="TFLOP/s", frac=achieved / tf32_peak if tf32_peak else None, traffic=None,
                      peak_source=f"{pk['source']} bf16 sustained {pk['bf16_sustained']} TF/s / 2 (TF32 pipe = half the bf16 rate)",
                      launches=dom[2], kernel_ms_per_step=dom[1],
                      note="achieved = executed conv FLOPs (2*M*Cout*Cin*taps) of the tcgen05 launches in one instrumented step / their "
                           "summed CUDA-event durations; traffic is null because the figure aggregates launches of many shapes",
                      traffic_profiled=dict(launch="conv_umma_patch_kernel<32,2,pair> 288x128x128 96->96 (+bias, scale, residual)",
                                            dram_bytes=5.66e9, algorithmic_bytes=5.44e9, tensor_pipe_active_pct=40.5,
                                            source="profiles/conv_umma_patch_pair_96x96_128_r01b_ncu.txt (ncu --set full)")),
        kernel_breakdown_ms={k: round(v[1], 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
    )
    if args.ref_gpu:
        line["reference_gpu_eager"] = gpu_reference_steps(cfg, args.ref_gpu, 2, 1, K, dev)
    if not args.no_cpu_baseline:
        r = cpu_reference_steps(cfg, args.cpu_batch, 1, 0, K, budget_s=90.0)
        line["cpu_baseline"] = dict(value=r["value"], unit="frames/s", cores=r["cores"], kind="port", sample=r["sample"])
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--cpu-sample-child":
        _cpu_sample_child(sys.argv[2:])
    else:
        main()
