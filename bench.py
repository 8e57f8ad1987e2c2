#!/usr/bin/env python
"""bench.py -- denoising steps/sec of the U-Net hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2a|cfg1]
    (N > 1: launched by torch.distributed.run, one rank per GPU)

Workload (default = BASELINE.json configs[2], the configuration the metric's target is quoted on; it fits one GPU):
    cfg3: SR U-Net 64->256, `Unet(**Super.defaults, lowres_cond=True, text_embed_dim=768)`, 256x256, batch 32 per GPU,
          T=1000 schedule, cond_scale=1 (one U-Net forward per step), synthetic inputs, random-init weights.
One "step" = one `Imagen._p_sample`: U-Net forward + x0 prediction + exact dynamic-threshold quantile + posterior sample.

Prints ONE JSON line (rank 0).  `value` = whole-job steps/s with inputs resident in HBM (CUDA-graph replay of the step,
noise drawn on the device); `e2e` = the same step driven through the public API with HOST (pinned) buffers copied in and
the result copied out every step.  `roofline` = algorithmic conv FLOPs of the tcgen05 implicit-GEMM launches divided by
their CUDA-event durations, against the measured bf16/fp16 tensor peak in MEASURED_PEAKS.json.  `cpu_baseline` = the CPU
oracle port (oracle/restatement.py, the reference's algorithm in torch fp32) timed on this box's host cores.
`--impl reference` times that CPU path alone (the reference has no other implementation of this path).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic GFLOP per image per U-Net forward, counted on the reference model
GFLOP_PER_IMG = {"cfg1": 1.30, "cfg2a": 131.70, "cfg3": 800.57}
METRIC = "denoising steps/sec (SR256 U-Net, batch 32 per GPU per step)"


def workload(name):
    from minimagen_b200.Unet import Super, BaseTest
    if name == "cfg3":
        return dict(cfg=dict(Super.defaults, lowres_cond=True, text_embed_dim=768), size=256, batch=32, T=1000,
                    E=768, lowres=True, desc="SR U-Net 64->256 dim=128 (Super.defaults, lowres_cond) b=32 256x256 T=1000")
    if name == "cfg2a":
        return dict(cfg=dict(text_embed_dim=768), size=64, batch=64, T=1000, E=768, lowres=False,
                    desc="base U-Net dim=128 (Unet ctor defaults) b=64 64x64 T=1000")
    if name == "cfg1":
        return dict(cfg=dict(BaseTest.defaults), size=64, batch=2, T=25, E=512, lowres=False,
                    desc="tiny base U-Net dim=8 b=2 64x64 T=25")
    raise SystemExit(f"unknown workload {name}")


def synth_inputs(wl, batch, seed):
    """Synthetic conditioning exactly shaped like the reference's inputs."""
    g = torch.Generator().manual_seed(seed)
    s, E = wl["size"], wl["E"]
    lengths = torch.randint(4, 65, (batch,), generator=g)
    L = int(lengths.max())
    text = torch.randn(batch, L, E, generator=g)
    mask = torch.arange(L)[None, :] < lengths[:, None]
    text = text * mask[..., None]                               # t5.py:82 zeroes padded positions
    d = dict(text_embeds=text, text_mask=mask, x=torch.randn(batch, 3, s, s, generator=g))
    if wl["lowres"]:
        d["lowres_img01"] = torch.rand(batch, 3, s, s, generator=g)     # up-sampled low-res image in [0,1]
        d["lowres_noise"] = torch.randn(batch, 3, s, s, generator=g)
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            load = [c for c, p in zip(sm, pw) if p > 300] or sm
            out.update(sm_mhz=statistics.median(load), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw))
        return out


def cpu_baseline(wl, sd, steps=3, warmup=1):
    """The reference's algorithm (CPU oracle port) on the host cores: U-Net forward + DDPM step at batch 1, scaled
    linearly to the workload batch (the full batch would take ~1 minute per step on 8 cores)."""
    from oracle import restatement as R
    try:
        import psutil
        cores = psutil.cpu_count(logical=False) or os.cpu_count()
    except Exception:
        cores = os.cpu_count()
    # one thread per PHYSICAL core: on the 2 x 32-core (HT) B200 hosts 128 threads is ~200x slower than 64
    torch.set_num_threads(cores)
    inp = synth_inputs(wl, 1, 123)
    tabs = R.ddpm_tables(wl["T"])
    t = torch.full((1,), wl["T"] - 1, dtype=torch.long)
    kw = dict(text_embeds=inp["text_embeds"], text_mask=inp["text_mask"])
    if wl["lowres"]:
        t_aug = torch.full((1,), int(wl["T"] * 0.2), dtype=torch.long)
        lr = R.q_sample(tabs, inp["lowres_img01"], t_aug, inp["lowres_noise"]) * 2 - 1
        kw.update(lowres_cond_img=lr, lowres_noise_times=t_aug)
    sd_cpu = {k: v.detach().float().cpu() for k, v in sd.items()}
    x = inp["x"]
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            eps = R.unet_forward(sd_cpu, wl["cfg"], x, t, **kw)
            x_next = R.p_sample_step(tabs, x, t, eps, torch.randn_like(x))
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    per_b1 = sum(times) / len(times)
    per_step = per_b1 * wl["batch"]
    return dict(value=1.0 / per_step, unit="steps/s", cores=cores, kind="port",
                sample=f"batch 1 of {wl['batch']}: {warmup} warm-up + {steps} timed (U-Net forward + DDPM step) on "
                       f"{cores} host threads, {per_b1:.2f} s each, scaled x{wl['batch']} to the workload batch",
                gflops=GFLOP_PER_IMG.get(wl.get('name', ''), 0) / per_b1 if per_b1 else None), float(x_next.std())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--batch-streams", type=int, default=None, help="concurrent batch slices inside Unet.forward")
    ap.add_argument("--fuse", default=None, choices=["off", "n128", "all"], help="fused GroupNorm+conv kernel usage")
    ap.add_argument("--kernel-table", default=None, help="write a CUPTI per-kernel time table of 3 steps to this path")
    ap.add_argument("--pdl", type=int, default=None, help="programmatic dependent launch on (1) / off (0)")
    ap.add_argument("--gn-f16", action="store_true", help="GroupNorm inputs in fp16 (faster, 1.05e-3 instead of 9e-4 rel-L2)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = workload(args.workload)
    wl["name"] = args.workload
    if args.batch:
        wl["batch"] = args.batch
    B = wl["batch"]
    config = {"workload": f"{args.workload}: {wl['desc']}", "batch_per_gpu": B, "global_batch": B * world,
              "image_size": wl["size"], "cond_scale": 1.0, "parallelism": f"dp{world} (batch-sharded sampling)",
              "l2": "per-step working set (activations + 1.4 GB fp16 weights) >> 126 MB L2, no explicit flush needed",
              "algorithmic_gflop_per_image_forward": GFLOP_PER_IMG.get(args.workload)}

    # ------------------------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        from minimagen_b200.Unet import Unet
        torch.manual_seed(0)
        sd = Unet(**wl["cfg"]).state_dict()
        base, _ = cpu_baseline(wl, sd, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 1)))
        v = base["value"]
        print(json.dumps({
            "impl": "reference", "metric": METRIC if args.workload == "cfg3" else f"denoising steps/sec ({args.workload})",
            "value": v, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": base, "gpu_launches": 0,
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------------------------ our arm (B200)
    assert torch.cuda.is_available(), "bench.py --impl ours needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from minimagen_b200 import _native, layers
    from minimagen_b200.Imagen import Imagen
    from minimagen_b200.Unet import Unet
    _native.load()
    if args.pdl is not None:
        _native.load().mi_set_launch_mode(int(args.pdl))
    if args.gn_f16:
        layers.GN_INPUT_F32 = False
    if args.fuse is not None:
        layers.FUSE_GN_CONV = {"off": False, "n128": True, "all": "all"}[args.fuse]

    torch.manual_seed(0)
    with torch.device(dev):
        unet = Unet(**wl["cfg"]).eval()
        if wl["lowres"]:
            # Imagen treats its first U-Net as the base model (Imagen.py:96-101), so the SR U-Net under test is stage 2
            # behind a tiny stand-in base stage that is never run here
            from minimagen_b200.Unet import BaseTest
            stages, sizes = (Unet(**BaseTest.defaults, text_embed_dim=wl["E"]), unet), (wl["size"] // 4, wl["size"])
        else:
            stages, sizes = (unet,), (wl["size"],)
    imagen = Imagen(unets=stages, text_encoder_name="t5_base" if wl["E"] == 768 else "t5_small",
                    image_sizes=sizes, timesteps=wl["T"], cond_drop_prob=0.1).eval().to(dev)
    assert imagen.unets[-1] is unet, "the U-Net under test was re-instantiated"
    if args.batch_streams is not None:
        unet.batch_streams = args.batch_streams
    sch = imagen.noise_schedulers[-1]
    inp = synth_inputs(wl, B, 1000 + rank)           # each rank owns its own shard of the global batch
    text = inp["text_embeds"].to(dev)
    mask = inp["text_mask"].to(dev)
    shape = (B, 3, wl["size"], wl["size"])
    n_img = 3 * wl["size"] * wl["size"]
    kw = dict(noise_scheduler=sch, text_embeds=text, text_mask=mask, lowres_cond_img=None, lowres_noise_times=None,
              cond_scale=1.0)
    ops = __import__("minimagen_b200.ops", fromlist=["get_ops"]).get_ops()
    if wl["lowres"]:
        t_aug = imagen.lowres_noise_schedule._get_times(B, 0.2, device=dev)
        lr = torch.empty(shape, device=dev)
        ops.q_sample(inp["lowres_img01"].to(dev), inp["lowres_noise"].to(dev), t_aug, sch.sqrt_alphas_cumprod,
                     sch.sqrt_one_minus_alphas_cumprod, B, n_img, 2.0, -1.0, lr)     # noise in [0,1] space, then *2-1
        kw.update(lowres_cond_img=lr, lowres_noise_times=t_aug)

    x_host = inp["x"].pin_memory()
    x = x_host.to(dev)
    T = wl["T"]

    with torch.no_grad():
        # one eager step: packs weights, warms the allocator, counts launches per step
        l0 = _native.launch_count
        t_dev = torch.full((B,), T - 1, dtype=torch.long, device=dev)
        imagen._step(unet, x, t_dev, torch.randn(shape, device=dev), **kw)
        torch.cuda.synchronize()
        l0 = _native.launch_count
        imagen._step(unet, x, t_dev, torch.randn(shape, device=dev), **kw)
        torch.cuda.synchronize()
        launches_per_step = _native.launch_count - l0

        print(f"[bench] launches/step={launches_per_step}", file=sys.stderr, flush=True)
        # per-kernel timing of the dominant kernel (tcgen05 implicit GEMM): CUDA events around every launch of one
        # eager step, on the launching stream
        conv = measure_conv_kernels(imagen, unet, x, t_dev, shape, kw, dev)
        conv_ms, conv_flops, conv_calls = conv["all"][0], conv["all"][1], conv["all"][2]

        # steady-state step function: CUDA graph replay (device-resident inputs)
        use_graph = not args.no_graph
        if use_graph:
            step_fn = imagen._graph_step_fn(unet, shape, **kw)
        else:
            step_fn = lambda xx, tt, nn: imagen._step(unet, xx, tt, nn, **kw)

        def run_steps(k, x0, t_start):
            cur = x0
            for i in range(k):
                t_dev.fill_(max(t_start - i, 0))
                cur = step_fn(cur, t_dev, torch.randn(shape, device=dev))
            return cur

        warm = run_steps(args.warmup, x, T - 1)
        if world > 1:
            # warm the collective too (NCCL builds channels / registers buffers on first use)
            fin = torch.empty_like(warm)
            ops.step_finalize(warm, warm.numel(), 1, fin)
            gathered = torch.empty((world * B, *shape[1:]), device=dev)
            dist.all_gather_into_tensor(gathered, fin)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        cur = run_steps(args.steps, x, T - 1 - args.warmup)
        if world > 1:
            fin = torch.empty_like(cur)
            ops.step_finalize(cur, cur.numel(), 1, fin)
            gathered = torch.empty((world * B, *shape[1:]), device=dev)
            dist.all_gather_into_tensor(gathered, fin)           # the path's single collective (finished images)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        print(f"[bench] device-resident: {ms / args.steps:.2f} ms/step; conv_tc {conv_ms:.2f} ms/step over {conv_calls} "
              f"launches", file=sys.stderr, flush=True)
        clocks = sampler.stop() if sampler else None
        assert torch.isfinite(cur).all(), "non-finite output"

        if args.kernel_table and rank == 0:
            kernel_table(lambda: run_steps(3, x, T - 1), 3, args.kernel_table)

        # end-to-end: public API call per step with host buffers (pinned) in, result out
        out_host = torch.empty(shape, dtype=torch.float32).pin_memory()
        noise_host = torch.randn(shape).pin_memory()
        t_host = torch.full((B,), T - 1, dtype=torch.long).pin_memory()

        def e2e_step(i):
            xd = x_host.to(dev, non_blocking=True)
            nd = noise_host.to(dev, non_blocking=True)
            td = t_host.to(dev, non_blocking=True)
            if use_graph:
                r = step_fn(xd, td, nd)
            else:
                r = imagen._p_sample(unet, xd, td, noise=nd, **kw)
            out_host.copy_(r, non_blocking=True)
            torch.cuda.synchronize()
        for i in range(2):
            e2e_step(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(args.steps):
            e2e_step(i)
        f1.record()
        torch.cuda.synchronize()
        e2e_ms = max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1000.0)

    times = torch.tensor([ms, e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(times[0]), float(times[1])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    burst_tf = peaks.get("bf16_tflops") or 1650.0
    value = world * args.steps / (ms / 1000.0)
    achieved = conv_flops / (conv_ms / 1000.0) / 1e12 if conv_ms > 0 else 0.0
    step_tflops = (value * B * GFLOP_PER_IMG.get(args.workload, 0.0)) / 1000.0 / world    # per GPU
    # dominant kernel = conv3x3_halo_t_kernel: algorithmic FLOPs per launch / CUDA-event time per launch, live; DRAM traffic
    # per launch from the committed ncu capture of the same step (profiles/r01_halo_t_dram_v16.json)
    t_ms, t_fl, t_n, t_bytes = conv["halo_t"]
    dom = t_n > 0
    d_ms, d_fl, d_n, d_bytes = (t_ms, t_fl, t_n, t_bytes) if dom else tuple(conv["all"])
    d_achieved = d_fl / (d_ms / 1000.0) / 1e12 if d_ms > 0 else 0.0
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_halo_t_dram_v16.json")))
        if dom and args.workload == "cfg3" and B == 32 and prof.get("launches") == d_n:
            traffic = prof["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "tensor",
                "kernel": "conv3x3_halo_t_kernel (tcgen05 swapped-operand 3x3 / 15x1 halo conv)" if dom else
                          "tcgen05 implicit-GEMM convolutions (all launches)",
                # launches are timed one by one with CUDA events inside an eager step (idle gaps between launches: not the
                # power-capped regime of the graph-replayed step) -> the BURST cuBLAS figure is the matching denominator
                "achieved": d_achieved, "peak": burst_tf, "unit": "TFLOP/s", "frac": d_achieved / burst_tf,
                "frac_of_sustained_peak": d_achieved / peak_tf,
                "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu, profiles/r01_halo_t_dram_v16.json)",
                "algorithmic_bytes_per_launch": d_bytes / d_n if d_n else None,
                "algorithmic_flops_per_launch": d_fl / d_n if d_n else None,
                "launches_timed": d_n, "kernel_ms_per_launch": d_ms / d_n if d_n else None,
                "kernel_ms_per_step": d_ms, "kernel_share_of_step": d_ms / (ms / args.steps) if ms else None,
                "all_conv_launches": {"achieved": achieved, "frac": achieved / burst_tf, "launches": conv_calls,
                                      "ms_per_step": conv_ms, "share_of_step": conv_ms / (ms / args.steps) if ms else None},
                "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst, kernel timed alone) for `frac`; bf16_tflops_sustained "
                               "(kernel inside a long step) for `whole_step_frac` and `frac_of_sustained_peak`"
                               if peaks else "fallback 1.65 / 1.4 PFLOP/s (B200_PROFILING.md)",
                "whole_step_tflops_per_gpu": step_tflops, "whole_step_frac": step_tflops / peak_tf,
                "whole_step_peak": peak_tf}
    result = {
        "metric": METRIC if args.workload == "cfg3" else f"denoising steps/sec ({args.workload})",
        "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 tensor-core operands, f32 accumulate / residual stream", "data": "synthetic", "config": config,
        "gpu_launches": launches_per_step * args.steps,
        "e2e": {"value": world * args.steps / (e2e_ms / 1000.0), "unit": "steps/s",
                "h2d_bytes_per_step": int(x_host.numel() * 4 + noise_host.numel() * 4 + t_host.numel() * 8),
                "d2h_bytes_per_step": int(out_host.numel() * 4)},
        "roofline": roofline,
        "clocks": clocks, "cuda_graph": use_graph, "launches_per_step": launches_per_step,
        "batch_streams": unet.batch_streams, "gn_input": "f32" if layers.GN_INPUT_F32 else "f16",
    }
    if not args.no_cpu_baseline:
        sd = unet.state_dict()
        base, _ = cpu_baseline(wl, sd)
        result["cpu_baseline"] = base
    print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


def kernel_table(fn, steps, path):
    """Diagnostics only (never a bench value): CUPTI kernel records of `steps` un-serialised steps, summed per kernel."""
    import collections
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fn()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type is not None and str(ev.device_type).endswith("CUDA"):
            a = agg[ev.name[:110]]
            a[0] += 1
            a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    tot = sum(t for _, t in agg.values())
    with open(path, "w") as f:
        f.write(f"# per-step kernel time (CUPTI, {steps} steps averaged), total {tot / steps / 1e3:.3f} ms/step\n")
        for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t / steps / 1e3:9.3f} ms {c / steps:7.1f} launches  {name}\n")


def measure_conv_kernels(imagen, unet, x, t_dev, shape, kw, dev):
    """Run one eager step with CUDA events around every tcgen05 conv launch.  Returns a dict with the totals over all
    conv launches and, separately, over the launches that run on the dominant kernel (conv3x3_halo_t_kernel: 3x3 / 15x1
    stride-1 convs with C_out % 128 == 0 on a 32x8- or 16x16-tileable grid -- the selection rule of csrc/capi.cu)."""
    from minimagen_b200 import ops as ops_mod
    real = ops_mod.get_ops()
    events = []

    class Timed(type(real)):
        def conv_igemm(self, act, B, H, W, lda, c_off, c_in, wp, c_out, kh, kw_, mode, bias, residual, out_f32, out_f16,
                       *a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            super().conv_igemm(act, B, H, W, lda, c_off, c_in, wp, c_out, kh, kw_, mode, bias, residual, out_f32, out_f16,
                               *a, **k)
            e.record()
            halo_t = (mode == 0 and (kh, kw_) in ((3, 3), (15, 1)) and c_out % 128 == 0 and
                      ((W == 16 and H % 16 == 0 and kh == 3) or (H % 32 == 0 and W % 8 == 0 and W != 16)))
            mn = B * H * W * c_out
            nbytes = (B * H * W * c_in * 2 * (4 if mode == 6 else 1) + c_out * kh * kw_ * c_in * 2 +
                      (4 * mn if residual is not None else 0) + (4 * mn if out_f32 is not None else 0) +
                      (2 * mn if out_f16 is not None else 0))
            events.append((s, e, 2.0 * mn * kh * kw_ * c_in, nbytes, halo_t))

    ops_mod.set_ops(Timed())
    streams = unet.batch_streams
    unet.batch_streams = 1            # one stream: every launch is timed alone, not while sharing SMs with the other half
    try:
        imagen._step(unet, x, t_dev, torch.randn(shape, device=dev), **kw)
        torch.cuda.synchronize()
    finally:
        ops_mod.set_ops(real)
        unet.batch_streams = streams
    res = {"all": [0.0, 0.0, 0, 0.0], "halo_t": [0.0, 0.0, 0, 0.0]}       # ms, flops, launches, algorithmic bytes
    for s, e, fl, nb, ht in events:
        ms = s.elapsed_time(e)
        for key in (("all", "halo_t") if ht else ("all",)):
            r = res[key]
            r[0] += ms; r[1] += fl; r[2] += 1; r[3] += nb
    return res


if __name__ == "__main__":
    main()
