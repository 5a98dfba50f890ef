#!/usr/bin/env python
"""bench.py — headline benchmark of the TEMPO-VAE hot path on B200.

  python bench.py --gpus N --steps K --warmup W            our arm (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K ...  the reference's algorithm on the host CPU cores

Metric (BASELINE.json): train samples/sec (fwd + bwd + clip + AdamW) of the default TEMPO-VAE
(configs/training shape [1028,64,64], chs [512,256,128], z 32) at batch 256 per GPU, bf16 tensor-core operands,
synthetic radiance-like patches, random-init weights. One "step" = Trainer.train_step on one batch.

  value  : device-resident input (NCHW fp32 already in HBM), CUDA-event timed, max over ranks
  e2e    : same step through the public API (Trainer.train_step) fed from PINNED HOST memory through the
           DevicePrefetcher (H2D copy of every batch inside the timed region, overlapped with the previous step)
           and a device->host read of the metrics every step
  roofline: the dominant kernel (conv_gemm_kernel on the 512->512 3x3 @64x64 layers, fwd + dgrad launches),
           bracketed by CUDA events inside the timed region; peak = MEASURED_PEAKS.json bf16_tflops_sustained
  cpu_baseline: the oracle port of the reference path on the host cores, bounded sample (N = 1 only)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_MODEL = dict(
    architecture_type="vae",
    architecture_params=dict(enc_dec_params=dict(
        shape=[1028, 64, 64], embed_dim=32, chs=[512, 256, 128], attn_sizes=[], mid_attn=True, num_res_blocks=1,
        dropout_prob=0.0, z_channels=32, double_z=True, n_attention_heads=4, norm_groups=8, norm_eps=1e-6,
        norm_affine=True, act="gelu", conv_kernel_size=3, conv_padding_mode="zeros", kl_weight=1e-6,
        nll_loss_type="l1")),
    optimizer_type="AdamW",
    optimizer_params=dict(lr=1e-4, betas=[0.9, 0.95], weight_decay=0.05),
)
FWD_GF, BWD_GF = 165.776, 292.746          # algorithmic conv GFLOP / sample
# dram bytes of one 512->512 3x3 @64x64 launch at B=256 from the committed ncu capture (profiles/ncu_gemm_r1.md,
# launch 1: 1.084 GB read + 2.104 GB written; the launch reads a 1.07 GB bf16 activation + 4.7 MB of weights and
# writes a 2.15 GB fp32 tensor)
NCU_CONV_TRAFFIC_BYTES = 3.188e9


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tflops=float(p["bf16_tflops_sustained"]), burst=float(p["bf16_tflops"]), hbm=float(p["hbm_gbs"]),
                    source="measured (MEASURED_PEAKS.json)")
    except Exception:  # noqa: BLE001
        return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:  # noqa: BLE001
                continue
        if not sm:
            return None
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def synthetic_batch(torch, B, shape, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn((B, *shape), device=device, generator=g, dtype=torch.float32)
    return x.clamp_(-10, 10)


# ================================================================================================= reference arm
def oracle_step_fn(torch, device="cpu", autocast=False):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tempo_vae_oracle as orc
    import tempo_vae_b200.model as m
    torch.manual_seed(42)
    vae = m.AutoencoderKL({k: v for k, v in m.DEFAULT_ENC_DEC.items()}, embed_dim=32, kl_weight=1e-6, nll_loss_type="l1")
    params = {k: v.detach().clone().to(device) for k, v in m.SpectralVAE(vae).state_dict().items()}
    state = {}
    cfg = orc.DEFAULT_CFG
    g = torch.Generator(device=device).manual_seed(0)
    step_no = [0]

    def step(B, x=None):
        if x is None:
            x = torch.randn((B, 1028, 64, 64), generator=g, device=device).clamp_(-10, 10)
        eps = torch.randn((B, 32, 16, 16), generator=g, device=device)

        def loss_fn(leaves):
            if autocast:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return orc.vae_loss(leaves, x, eps, cfg)
            return orc.vae_loss(leaves, x, eps, cfg)
        grads, out = orc.grads_of(loss_fn, params)
        step_no[0] += 1
        orc.clip_and_adamw(params, grads, state, step=step_no[0])
        return out["loss"]
    return step


def run_reference_gpu(args):
    """Same-box comparator: the oracle's restatement of the reference train step executed by
    stock PyTorch eager kernels (cuDNN/cuBLAS/ATen) on the B200 -- what a user of the reference gets on this GPU.
    `--ref-device cuda --ref-precision {fp32,tf32,bf16}`; inputs resident in HBM, CUDA events."""
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    tf32 = args.ref_precision != "fp32"
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    step = oracle_step_fn(torch, dev, autocast=(args.ref_precision == "bf16"))
    B = args.batch
    xs = [synthetic_batch(torch, B, (1028, 64, 64), dev, seed=i) for i in range(2)]
    for i in range(max(args.warmup, 1)):
        step(B, xs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(B, xs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    v = B / (ms / 1e3)
    print(json.dumps({
        "impl": "reference", "comparator": f"oracle restatement on PyTorch eager CUDA kernels, {args.ref_precision}",
        "metric": "train samples/sec (fwd+bwd+AdamW)", "value": v, "unit": "samples/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.ref_precision, "data": "synthetic",
        "config": {"workload": "default TEMPO-VAE train step, synthetic patches [1028,64,64]", "batch_per_gpu": B},
        "peak_hbm_gb": torch.cuda.max_memory_allocated(dev) / 1e9, "final_loss": float(loss),
    }), flush=True)


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = oracle_step_fn(torch)
    t0 = time.perf_counter(); step(1); t1 = time.perf_counter() - t0            # also the first warm-up
    budget = 150.0
    B = int(max(1, min(8, budget / max(1e-3, (args.steps + max(args.warmup - 1, 0)) * t1))))
    for _ in range(max(args.warmup - 1, 0)):
        step(B)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(B)
    dt = time.perf_counter() - t0
    v = B * args.steps / dt
    sample = f"{args.steps} timed oracle train steps (fwd+bwd+clip+AdamW, fp32) of the default model at batch {B}"
    print(json.dumps({
        "impl": "reference", "metric": "train samples/sec (fwd+bwd+AdamW)", "value": v, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "default TEMPO-VAE train step, synthetic patches [1028,64,64]", "batch_per_step": B},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ================================================================================================= our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import tempo_vae_b200 as t
    from tempo_vae_b200 import ops
    from tempo_vae_b200.parallel import DataParallel
    from tempo_vae_b200.tempo_data import DevicePrefetcher

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        from tempo_vae_b200.parallel import bind_to_gpu_numa
        numa_bound = bind_to_gpu_numa(local)          # pinned host batches on the GPU's own NUMA node
        dist.init_process_group("nccl", device_id=dev)
    else:
        numa_bound = None
    B = args.batch
    shape = (1028, 64, 64)

    t.seed_all(42)
    model = t.get_model(DEFAULT_MODEL, dev)
    trainer = t.Trainer(model, model.optimizer, dev, tempfile.mkdtemp(prefix="tvae_bench_"))
    dp = DataParallel(model, model.optimizer) if world > 1 else None

    def step_device(x):
        if dp is not None:
            m = dp.train_step_device(x)
            m["pixel_mse"] = model.vae.last_pixel_mse()
            return m
        return trainer.train_step_device(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        tns = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    # ---------------------------------------------------------------- kernel-only number (inputs resident in HBM)
    xs = [synthetic_batch(torch, B, shape, dev, seed=1000 * rank + i) for i in range(2)]
    for i in range(args.warmup):
        step_device(xs[i % 2])
        trainer.step = 1                  # (the reference prints batch statistics while step == 0)
    M = B * 64 * 64
    ops.PROFILE["conv"] = {"match": lambda px, co, ci, kind, R: px == M and co == 512 and ci == 512 and kind == 0 and R == 3,
                           "events": []}
    ops.PROFILE["wgrad"] = {"match": lambda px, cm, cn, kind, R: px == M and cm == 512 and cn == 512 and kind == 0 and R == 3,
                            "events": []}
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = ops.KERNEL_LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    th0 = time.perf_counter()
    for i in range(args.steps):
        last = step_device(xs[i % 2])
    host_enqueue_ms = (time.perf_counter() - th0) / args.steps * 1e3     # CPU time to enqueue one step (no sync)
    e1.record()
    barrier()
    launches = ops.KERNEL_LAUNCHES[0] - n0
    clocks = sampler.stop() if sampler is not None else None
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    conv_ms = [a.elapsed_time(b) for a, b in ops.PROFILE["conv"]["events"]]
    wg_ms = [a.elapsed_time(b) for a, b in ops.PROFILE["wgrad"]["events"]]
    ops.PROFILE.clear()
    final = {k: float(v.detach()) for k, v in last.items()}
    value = world * B / (ms / 1e3)

    # ---------------------------------------------------------------- end to end (host buffers, public API)
    if args.skip_e2e:
        if rank == 0:
            print(json.dumps({"profiling_run": True, "value": value, "ms_per_step": ms, "gpu_launches": launches,
                              "host_enqueue_ms_per_step": host_enqueue_ms}))
        if world > 1:
            dist.destroy_process_group()
        return
    host = [torch.empty((B, *shape), dtype=torch.float32).pin_memory() for _ in range(2)]
    for i, h in enumerate(host):
        h.copy_(xs[i])
    # the same batches in the engine's own loader format (channels-last bf16 rows, pitch 1032: what DeviceTileCache /
    # a bf16 tile store holds; cast once at load time, bit-identical results because the engine rounds its input to
    # bf16 first thing) -- used for the extra e2e_loader_format leg below
    pitch = (shape[0] + 7) // 8 * 8
    host_cl = [torch.zeros((B, shape[1], shape[2], pitch), dtype=torch.bfloat16).pin_memory() for _ in range(2)]
    for i, h in enumerate(host_cl):
        h[..., :shape[0]].copy_(xs[i].permute(0, 2, 3, 1))
    del xs
    torch.cuda.empty_cache()

    def host_stream(n):
        for i in range(n):
            yield host[i % 2]

    if dp is None:
        e2e_step = trainer.train_step                       # floats: one device->host read per step
    else:
        def e2e_step(x):
            m = step_device(x)
            return t.train_utils._to_floats(m)
    pf = DevicePrefetcher(host_stream(args.warmup + args.steps), dev)
    it = iter(pf)
    for _ in range(args.warmup):
        e2e_step(next(it))
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        m = e2e_step(next(it))
    e1.record()
    barrier()
    wall = (time.perf_counter() - t0) / args.steps * 1e3
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1) / args.steps, wall))
    e2e_value = world * B / (e2e_ms / 1e3)
    h2d = B * shape[0] * shape[1] * shape[2] * 4
    d2h = 4 * len(m)

    # extra leg: host batches in the loader format (half the H2D bytes), same API call, same timing rules
    def cl_stream(n):
        for i in range(n):
            yield host_cl[i % 2]
    it = iter(DevicePrefetcher(cl_stream(args.warmup + args.steps), dev, dtype=torch.bfloat16))

    def as_nchw_view(d):
        return d[..., :shape[0]].permute(0, 3, 1, 2)        # [B, C, H, W]-shaped, channels-last strides: used in place
    for _ in range(args.warmup):
        e2e_step(as_nchw_view(next(it)))
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        e2e_step(as_nchw_view(next(it)))
    e1.record()
    barrier()
    wall = (time.perf_counter() - t0) / args.steps * 1e3
    cl_ms = max_over_ranks(max(e0.elapsed_time(e1) / args.steps, wall))
    cl_h2d = host_cl[0].numel() * 2

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    conv_flops = 2.0 * M * 512 * (9 * 512)
    conv_avg = statistics.mean(conv_ms) if conv_ms else float("nan")
    achieved = conv_flops / (conv_avg * 1e-3) / 1e12
    wg_avg = statistics.mean(wg_ms) if wg_ms else float("nan")
    out = {
        "metric": "train samples/sec (fwd+bwd+AdamW)", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "default TEMPO-VAE train step (configs/training/train_vae_default.yaml model), "
                               "synthetic patches [1028,64,64] clamp(N(0,1),-10,10), random-init weights",
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "l2_flush": "inputs (4.3 GB/step) and activations are far larger than the 126 MB L2",
                   "useful_gflop_per_sample": FWD_GF + BWD_GF},
        "step_tflops": value * (FWD_GF + BWD_GF) / 1e3,
        "step_frac_of_peak": value * (FWD_GF + BWD_GF) / 1e3 / (pk["tflops"] * world),
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_kernel 512->512 3x3 @64x64 (fwd and dgrad launches)",
                     "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                     "traffic": NCU_CONV_TRAFFIC_BYTES if B == 256 else None,
                     "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch "
                                       "(profiles/ncu_gemm_r1.md); algorithmic bytes per launch 3.22e9",
                     "peak_source": pk["source"] + ", bf16_tflops_sustained (cuBLAS back to back for 4 s: the figure "
                                    "for a kernel timed inside a long step)",
                     "frac_of_burst_peak": achieved / pk["burst"], "burst_peak": pk["burst"],
                     "launches_timed": len(conv_ms), "avg_ms": conv_avg,
                     "wgrad_kernel": {"avg_ms": wg_avg, "achieved": conv_flops / (wg_avg * 1e-3) / 1e12,
                                      "launches_timed": len(wg_ms)}},
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "api": "Trainer.train_step over DevicePrefetcher (pinned host batches)"},
        "e2e_loader_format": {"value": world * B / (cl_ms / 1e3), "unit": "samples/s", "ms_per_step": cl_ms,
                              "h2d_bytes_per_step": cl_h2d, "d2h_bytes_per_step": d2h,
                              "note": "same call, pinned host batches held as channels-last bf16 (the engine's tile-store "
                                      "format, results bit-identical); not the headline e2e, which copies the "
                                      "reference loader's fp32 NCHW batches"},
        "gpu_launches": launches,
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "numa_bound": numa_bound, "host_cpus": len(os.sched_getaffinity(0)),
        "peak_hbm_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
        "clocks": clocks,
        "final_metrics": final,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        step = oracle_step_fn(torch)
        step(1)
        cb = 4
        t0 = time.perf_counter(); step(cb); dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": cb / dt, "unit": "samples/s", "cores": cores, "kind": "port",
                               "sample": f"1 oracle train step (fp32 fwd+bwd+clip+AdamW) of the default model at "
                                         f"batch {cb} after a batch-1 warm-up, torch CPU threads = {cores}"}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_extra(args):
    """Secondary workloads (parity-test configs of BASELINE.json measured for reference, not the headline)."""
    import torch
    import torch.distributed as dist
    import tempo_vae_b200 as t

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    t.seed_all(42)
    model = t.get_model(DEFAULT_MODEL, dev)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.workload == "train_l2":
        l2 = t.VAEWithL2Supervision(model.vae, latent_channels=32, mlp_hidden=[512, 512]).to(dev)
        opt = t.FusedAdamW(l2.parameters(), lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05)
        trainer = t.L2SupervisedTrainer(l2, opt, dev, tempfile.mkdtemp(prefix="tvae_bench_"), kl_weight=1e-6,
                                        l2_weights={"NO2": 0.1, "O3TOT": 0.1, "HCHO": 0.1, "CLDO4": 0.1})
        trainer.step = 1
        g = torch.Generator(device=dev).manual_seed(rank)
        batch = {"spectral": synthetic_batch(torch, B, (1028, 64, 64), dev, seed=rank)}
        for p in ("NO2", "O3TOT", "HCHO", "CLDO4"):
            tg = torch.randn((B, 64, 64), device=dev, generator=g)
            blob = torch.nn.functional.interpolate(torch.rand((B, 1, 8, 8), device=dev, generator=g), size=(64, 64))[:, 0]
            tg[blob < 0.15] = float("nan")            # ~15 % invalid pixels in contiguous blobs
            batch[p] = tg
        for _ in range(args.warmup):
            trainer.train_step_device(batch)
        sync(); e0.record()
        for _ in range(args.steps):
            m = trainer.train_step_device(batch)
        e1.record(); sync()
        ms = e0.elapsed_time(e1) / args.steps
        out = {"workload": "VAEWithL2Supervision train step (config 3), default model + 282,628-parameter L2 head",
               "metric": "train samples/sec (fwd+bwd+AdamW)", "value": world * B / ms * 1e3, "unit": "samples/s",
               "ms_per_step": ms, "n_gpus": world, "batch_per_gpu": B,
               "final_metrics": {k: float(v.detach()) for k, v in m.items()}}
    elif args.workload == "train_cached":
        # the train step fed from a device-resident channels-last bf16 tile cache
        trainer = t.Trainer(model, model.optimizer, dev, tempfile.mkdtemp(prefix="tvae_bench_"))
        trainer.step = 1
        n_tiles = 4 * B
        cache = t.DeviceTileCache(dev, 64, 64, 1028, n_tiles)
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        for _ in range(n_tiles // 64):
            cache.add(torch.randn((64, 64, 64, 1028), device=dev, generator=g).clamp_(-10, 10))
        it = cache.batches(B, seed=rank)
        for _ in range(args.warmup):
            trainer.train_step_device(next(it))
        sync(); e0.record()
        for _ in range(args.steps):
            m = trainer.train_step_device(next(it))
        e1.record(); sync()
        ms = e0.elapsed_time(e1) / args.steps
        out = {"workload": "train step fed from DeviceTileCache (tiles resident in HBM as channels-last bf16, batches "
                           "gathered on the device and consumed without a layout pass)",
               "metric": "train samples/sec (fwd+bwd+AdamW)", "value": world * B / ms * 1e3, "unit": "samples/s",
               "ms_per_step": ms, "n_gpus": world, "batch_per_gpu": B, "cached_tiles_per_gpu": n_tiles,
               "final_metrics": {k: float(v.detach()) for k, v in m.items()}}
    else:
        # synthetic granules [131, 2048, 1028] -> normalise -> crop [128, 2048] -> 64 patches of [1028, 64, 64] each
        n_gran = 4
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        mean_s = torch.full((1028,), 3.0, device=dev)
        std_s = torch.full((1028,), 0.5, device=dev)
        patches = []
        for _ in range(n_gran):
            rad = torch.exp(torch.randn((131, 2048, 1028), device=dev, generator=g) * 0.5 + 3.0)
            patches.append(t.granule_to_patches(t.normalize_radiance(rad, mean_s, std_s)))
            del rad
        patches = torch.cat(patches)                    # [256, 1028, 64, 64] per rank
        for _ in range(args.warmup):
            t.encode_patches(model, patches, batch_size=B)
        sync(); e0.record()
        for _ in range(args.steps):
            lat = t.encode_patches(model, patches, batch_size=B)
        e1.record(); sync()
        ms = e0.elapsed_time(e1) / args.steps
        n = patches.shape[0]
        out = {"workload": "encode-only patch sweep (config 5): posterior means of 64x64 patches of synthetic granules",
               "metric": "encoded patches/sec", "value": world * n / ms * 1e3, "unit": "patches/s", "ms_per_sweep": ms,
               "patches_per_gpu": n, "n_gpus": world, "latent_shape": list(lat.shape[1:]),
               "encoder_tflops": world * n / ms * 1e3 * 84.248 / 1e3}
    if world > 1:
        tns = torch.tensor([out["value"]], device=dev, dtype=torch.float64)
        dist.all_reduce(tns, op=dist.ReduceOp.MIN)
        out["value"] = float(tns.item())
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="samples per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: cpu = the contract's reference arm; cuda = same-box PyTorch-eager comparator")
    ap.add_argument("--ref-precision", default="tf32", choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: skip the host-fed leg")
    ap.add_argument("--workload", default="train", choices=["train", "train_l2", "train_cached", "encode"],
                    help="train = headline (BASELINE config 2/4); train_l2 = L2-supervised variant (config 3); "
                         "encode = inference-only patch sweep of synthetic granules (config 5); train_cached = train step "
                         "fed from the device-resident tile cache (SURVEY 8f). The extra workloads "
                         "print their own JSON line and are not the headline metric.")
    args = ap.parse_args()
    if args.impl == "reference" and args.ref_device == "cuda":
        run_reference_gpu(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_ours(args)
    else:
        run_extra(args)


if __name__ == "__main__":
    main()
