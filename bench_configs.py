"""The other BASELINE.json configs behind `bench.py --config ...` — same JSON contract as the headline line
(metric / value / unit / n_gpus / steps / warmup / ms_per_step / higher_is_better / scaling / vs_baseline / dtype / data /
config.workload / e2e / gpu_launches / clocks / roofline / cpu_baseline), one line on stdout.

    unet_train   configs[0]  U-Net train step, 32^3, batch 8 (4 input channels: unet.py:240 default)
    vae64        configs[3]  VAE+DFC train step, 64^3, batch 16
    unet64       configs[3]  U-Net train step, 64^3, batch 16
    inference    configs[4]  generate.py:202-225 — decode 100k latent samples + lattice parameters + U-Net segmentation
                             + argmax / 0.8 threshold, device resident (icsg3d_b200/pipeline.py), batch 100
    voxeliser    configs[4]  on-device Gaussian voxeliser throughput (utils.density_matrix, csrc/voxelize.cu)

`value` is device-timed with inputs resident in HBM (CUDA-graph replay where the step is captured); `e2e` goes through
the public API with pinned host buffers, copies inside the timed region.  N > 1 (torchrun): independent replicas /
batch shards per rank as noted per config; value = all ranks' units / max-over-ranks time.  `--impl reference` times
the oracle port of the same workload on the host cores (rank 0 only).
"""
from __future__ import annotations

import json
import os
import time

from bench import ClockSampler, load_peaks

UNET_GF = 377.94   # U-Net train step GFLOP per sample @32^3 (4 input channels); x8 @64^3
VAE_GF = 36.05     # VAE+DFC train step GFLOP per sample @32^3; x8 @64^3
INFER_GF = 127.76  # decoder 1.70 + full U-Net 126.06 GFLOP per generated sample
VOX_BYTES = 557056  # 32^3 x (4 ch fp32 + uint8 species) written per voxelised sample


def _dist():
    import torch
    import torch.distributed as dist
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the icsg3d hot path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local, dev


def _timed_loop(fn, steps, warmup, world, dev, sampler=None):
    """W untimed + K timed calls of fn bracketed by barrier + synchronize, CUDA events, max over ranks -> ms per step."""
    import torch
    import torch.distributed as dist
    for _ in range(max(warmup, 3)):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


def _wall_loop(fn, steps, world, dev):
    import torch
    import torch.distributed as dist
    for _ in range(3):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


def _conv_roofline(run_eager, peaks, clocks, reps=2):
    """Per-launch CUDA-event timing of every tcgen05 conv launch of one eager step (one stream)."""
    import torch
    from icsg3d_b200 import ops
    run_eager()
    torch.cuda.synchronize()
    ops.TIMING = []
    for _ in range(reps):
        torch.cuda._sleep(40_000_000)
        run_eager()
    torch.cuda.synchronize()
    rec, ops.TIMING = ops.TIMING, None
    agg = {}
    for (kind, tag), fl, a, b, ex in rec:
        if kind == "bn":
            continue
        d = agg.setdefault(kind, [0.0, 0.0, 0, 0.0])
        d[0] += fl
        d[1] += a.elapsed_time(b)
        d[2] += 1
        d[3] += ex
    if not agg:
        return None
    capped = bool(clocks and "sw_power_cap" in (clocks.get("reasons") or []))
    peak = peaks["bf16_tflops_sustained"] if capped else peaks["bf16_tflops"]
    # FLOPs credited to a kernel = min(nominal, executed): padding executed beyond the reference's math earns nothing, and
    # where the algorithm executes FEWER multiply-adds than the reference formulation (Upsample-into-Conv fold: 8 taps for
    # 27) the tensor pipe is credited with what it really did; the nominal rate is reported beside it.
    cred = {k: min(v[0], v[3]) if v[3] > 0 else v[0] for k, v in agg.items()}
    by = {k: {"tflops": cred[k] / (v[1] * 1e-3) / 1e12, "frac_of_peak": cred[k] / (v[1] * 1e-3) / 1e12 / peak,
              "nominal_tflops": v[0] / (v[1] * 1e-3) / 1e12,
              "launches_per_step": v[2] // reps, "ms_per_step": v[1] / reps, "gflop_per_step": v[0] / reps / 1e9,
              "executed_gflop_per_step": v[3] / reps / 1e9} for k, v in agg.items()}
    dom = max(by, key=lambda k: by[k]["ms_per_step"])
    tf, tms, tn = sum(cred.values()), sum(v[1] for v in agg.values()), sum(v[2] for v in agg.values())
    tnom = sum(v[0] for v in agg.values())
    kn = {"stream": "conv3d_k3_stream_kernel", "halo": "conv3d_k3_halo_kernel", "pertap": "conv3d_k3_igemm_kernel",
          "igemm": "conv3d_k1_igemm (1x1x1 heads)", "wgrad": "conv3d_k3_wgrad(_stream)_kernel",
          "upfold": "conv3d_k3_upfold_kernel"}
    return {"bound": "tensor", "kernel": "dominant by time: " + kn.get(dom, dom), "achieved": by[dom]["tflops"], "peak": peak,
            "unit": "TFLOP/s", "frac": by[dom]["tflops"] / peak, "traffic": None,
            "peak_source": peaks["src"] + (" sustained (sw_power_cap seen)" if capped else " burst (no power cap during the run)"),
            "algorithmic_gflop_per_launch": cred[dom] / reps / 1e9 / by[dom]["launches_per_step"],
            "avg_launch_ms": by[dom]["ms_per_step"] / by[dom]["launches_per_step"],
            "aggregate_all_conv": {"achieved": tf / (tms * 1e-3) / 1e12, "frac": tf / (tms * 1e-3) / 1e12 / peak,
                                   "nominal_tflops": tnom / (tms * 1e-3) / 1e12,
                                   "launches_per_step": tn // reps, "kernel_ms_per_step": tms / reps},
            "by_kernel": by}


def _emit(args, world, metric, value, ms_step, dtype, workload, e2e, launches, clocks, roof, cpu, extra=None, scaling="weak"):
    line = {"metric": metric, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": dtype,
            "data": "synthetic", "config": workload, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roof, "cpu_baseline": cpu}
    if extra:
        line.update(extra)
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
# CPU arms (oracle port; bounded samples)
# ----------------------------------------------------------------------------------------------------------
def _cpu_unet_train(d, batch, steps):
    import torch
    from oracle import nets
    from tests.util import synthetic_batch
    torch.set_num_threads(os.cpu_count() or 1)
    M, _, S = synthetic_batch(batch, d=d, seed=0)
    p = nets.init_unet_params(2)
    opt = nets.KerasAdam(3e-6)
    names = nets.trainable_names(p)

    def step():
        leaves = {k: p[k].clone().requires_grad_(True) for k in names}
        q = dict(p)
        q.update(leaves)
        out, _, _ = nets.unet_loss(q, M, S.long(), training=True, weight=95.0)
        grads = torch.autograd.grad(out[0], [leaves[k] for k in names])
        opt.step(p, dict(zip(names, grads)))

    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def _cpu_vae_train(d, batch, steps):
    import torch
    from oracle import nets
    from tests.util import synthetic_batch
    torch.set_num_threads(os.cpu_count() or 1)
    M, cond, _ = synthetic_batch(batch, d=d, seed=0)
    pv, pu = nets.init_vae_params(1, d=d), nets.init_unet_params(2)
    opt = nets.KerasAdam(5e-4)
    gen = torch.Generator().manual_seed(0)
    nets.vae_train_step(pv, pu, opt, M, cond, torch.randn(batch, 256, generator=gen))
    t0 = time.perf_counter()
    for _ in range(steps):
        nets.vae_train_step(pv, pu, opt, M, cond, torch.randn(batch, 256, generator=gen))
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def _cpu_inference(batch, steps):
    import numpy as np
    import torch
    from oracle import nets
    torch.set_num_threads(os.cpu_count() or 1)
    pv, pu = nets.init_vae_params(1), nets.init_unet_params(2)
    g = torch.Generator().manual_seed(3)
    cond = torch.eye(10)[torch.randint(0, 10, (batch,), generator=g)]

    def step():
        with torch.no_grad():
            z = torch.randn(batch, 256, generator=g) * 0.5
            xh = nets.vae_decoder(pv, z, cond, training=False)
            p = xh[..., 1:].numpy()
            mx, mn = p.max(axis=(1, 2, 3)), p.min(axis=(1, 2, 3))
            ap = (mx - mn) / 1.5 / (1 - 1 / 32)
            _ = ap - ap / 32
            soft, sig = nets.unet_forward(pu, xh, training=False)
            _ = soft.argmax(-1), torch.sigmoid(sig) >= 0.8

    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def _cpu_voxeliser(ncells):
    import numpy as np
    from oracle import voxelizer as vox
    rng = np.random.default_rng(0)
    cells = [vox.synthetic_cell(rng) for _ in range(ncells)]
    t0 = time.perf_counter()
    for N, z, l, sigma in cells:
        vox.density_matrix(N, z, l, dims=(32, 32, 32), sigma=sigma)
        vox.coordinate_grid(l, dim=32)
    dt = time.perf_counter() - t0
    return ncells / dt, dt / ncells * 1e3, 1


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    c = args.config
    steps = max(1, args.steps)
    if c in ("unet_train", "unet64"):
        d = 32 if c == "unet_train" else 64
        b = 2 if d == 32 else 1
        k = min(steps, 4 if d == 32 else 1)
        sps, ms, cores = _cpu_unet_train(d, b, k)
        sample = f"{k} oracle U-Net train steps at batch {b} @{d}^3 (per-sample cost is batch independent)"
    elif c == "vae64":
        k = min(steps, 3)
        sps, ms, cores = _cpu_vae_train(64, 2, k)
        sample = f"{k} oracle VAE+DFC train steps at batch 2 @64^3"
    elif c == "inference":
        k = min(steps, 4)
        sps, ms, cores = _cpu_inference(4, k)
        sample = f"{k} batches of 4 samples: oracle decoder + lattice params + U-Net forward + argmax/threshold"
    else:
        n = min(max(steps, 1) * 8, 64)
        sps, ms, cores = _cpu_voxeliser(n)
        sample = f"{n} cells through the numpy restatement of utils.density_matrix + coordinate_grid (1 thread)"
    line = {"impl": "reference", "metric": METRICS[c], "value": sps, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if c == "voxeliser" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[c], "note": "reference CPU arm: oracle port (TF/Keras not installable)"},
            "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


METRICS = {"unet_train": "unet_train_samples_per_sec_32cubed", "vae64": "vae_dfc_train_samples_per_sec_64cubed",
           "unet64": "unet_train_samples_per_sec_64cubed", "inference": "generate_samples_per_sec_32cubed",
           "voxeliser": "voxeliser_samples_per_sec_32cubed"}
WORKLOADS = {"unet_train": "U-Net train step @32^3, batch 8/GPU, 4 input channels (configs[0])",
             "vae64": "VAE+DFC train step @64^3, batch 16/GPU (configs[3])",
             "unet64": "U-Net train step @64^3, batch 16/GPU (configs[3])",
             "inference": "generate.py loop: decode + lattice params + U-Net + argmax/0.8 threshold, 100 samples per batch "
                          "(configs[4]; default --steps 1000 = 100k samples)",
             "voxeliser": "Gaussian density voxeliser + coordinate grid + species grid @32^3, 4096 cells per launch (configs[4])"}


# ----------------------------------------------------------------------------------------------------------
# CUDA arms
# ----------------------------------------------------------------------------------------------------------
def _run_unet_train(args, d, B):
    import torch
    from icsg3d_b200 import ops, utils
    from icsg3d_b200.engine import Dist
    from icsg3d_b200.unet.unet import AtomUnet
    world, rank, local, dev = _dist()
    peaks = load_peaks()
    # data parallel U-Net (N > 1): batch shards, NCCL all-reduce of BatchNorm sums + flat gradient, eager launches
    unet = AtomUnet(input_shape=(d, d, d, 4), lr=3e-6, device=dev, dist=Dist() if world > 1 else None, seed=2)
    eng = unet.engine(B)
    M, _, S = utils.synthetic_batch(B, d=d, seed=1000 + rank, device=dev)
    eng.set_inputs(M, S)
    l0 = ops.launch_count()
    eng._train_body()
    torch.cuda.synchronize()
    launches = ops.launch_count() - l0
    if not args.no_graph:
        eng.capture_train_graph()
    sampler = ClockSampler(local)
    ms = _timed_loop(eng.train_step, args.steps, args.warmup, world, dev, sampler if rank == 0 else None)
    value = B * world / (ms * 1e-3)
    Mh, Sh = M.cpu().pin_memory(), S.cpu().pin_memory()
    # the batch loop of fit_generator (AtomUnet.fit_epoch): H2D of every batch from pinned memory, metrics D2H every step
    n_e2e = max(3, args.steps // 4)
    unet.fit_epoch([(Mh, Sh)] * 3)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    unet.fit_epoch([(Mh, Sh)] * n_e2e)
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item()) / n_e2e
    clocks = sampler.stop() if rank == 0 else None
    eng.overlap_wgrad = False  # one stream: every launch of the roofline pass is timed alone
    roof = _conv_roofline(eng._train_body, peaks, clocks) if rank == 0 or world > 1 else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b, k = (2, 2) if d == 32 else (1, 1)
        sps, _, cores = _cpu_unet_train(d, b, k)
        cpu = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{k} oracle U-Net train steps at batch {b} @{d}^3 of the batch-{B} workload"}
    if rank == 0:
        gf = UNET_GF * (d // 32) ** 3
        _emit(args, world, METRICS[args.config], value, ms, "bf16",
              {"workload": WORKLOADS[args.config], "grid": d, "batch_per_gpu": B, "global_batch": B * world,
               "parallelism": f"dp{world}" if world > 1 else "single", "gflop_per_sample": gf,
               "l2": "activations of one step exceed the 126 MB L2; no flush needed", "cuda_graph": eng.use_graph},
              {"value": B * world / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": Mh.numel() * 4 + Sh.numel(),
               "d2h_bytes_per_step": 5 * 4}, launches * args.steps, clocks, roof, cpu,
              {"conv_tflops_whole_step": gf * value / 1e3, "loss": eng.metrics_host()})


def _run_vae64(args):
    import torch
    from icsg3d_b200 import ops, utils
    from icsg3d_b200.engine import Dist
    from icsg3d_b200.vae.lattice_vae import LatticeDFCVAE
    world, rank, local, dev = _dist()
    peaks = load_peaks()
    d, B = 64, args.batch if args.batch != 32 else 16
    vae = LatticeDFCVAE(input_shape=(d, d, d, 4), perceptual_model=None, device=dev, dist=Dist() if world > 1 else None, seed=1,
                        use_cuda_graph=not args.no_graph)
    vae._set_model(batch_size=B)
    eng = vae.engine(B)
    M, cond, _ = utils.synthetic_batch(B, d=d, seed=1000 + rank, device=dev)
    eng.set_inputs(M, cond, torch.randn(B, 256, device=dev))
    l0 = ops.launch_count()
    eng._train_body()
    torch.cuda.synchronize()
    launches = ops.launch_count() - l0
    if not args.no_graph:
        eng.capture_train_graph(snapshot=False)
    sampler = ClockSampler(local)
    ms = _timed_loop(eng.train_step, args.steps, args.warmup, world, dev, sampler if rank == 0 else None)
    value = B * world / (ms * 1e-3)
    Mh, ch = M.cpu().pin_memory(), cond.cpu().pin_memory()
    # the batch loop of train() (LatticeDFCVAE.fit_epoch): H2D of every batch from pinned memory, metrics D2H every step
    n_e2e = max(3, args.steps // 4)
    vae.fit_epoch([(Mh, ch)] * 3)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vae.fit_epoch([(Mh, ch)] * n_e2e)
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item()) / n_e2e
    clocks = sampler.stop() if rank == 0 else None
    eng.overlap_pm = eng.overlap_wgrad = False
    roof = _conv_roofline(eng._train_body, peaks, clocks)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sps, _, cores = _cpu_vae_train(64, 2, 2)
        cpu = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"2 oracle VAE+DFC train steps at batch 2 @64^3 of the batch-{B} workload"}
    if rank == 0:
        gf = VAE_GF * 8
        _emit(args, world, METRICS["vae64"], value, ms, "bf16",
              {"workload": WORKLOADS["vae64"], "grid": d, "batch_per_gpu": B, "global_batch": B * world,
               "parallelism": f"dp{world}" if world > 1 else "single", "gflop_per_sample": gf,
               "l2": "activations of one step exceed the 126 MB L2; no flush needed", "cuda_graph": not args.no_graph},
              {"value": B * world / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": Mh.numel() * 4 + ch.numel() * 4,
               "d2h_bytes_per_step": 16}, launches * args.steps, clocks, roof, cpu,
              {"conv_tflops_whole_step": gf * value / 1e3, "loss": eng.metrics_host()})


def _run_inference(args):
    import numpy as np
    import torch
    from icsg3d_b200 import ops
    from icsg3d_b200.pipeline import GeneratePipeline
    from icsg3d_b200.unet.unet import AtomUnet
    from icsg3d_b200.vae.lattice_vae import LatticeDFCVAE
    world, rank, local, dev = _dist()
    peaks = load_peaks()
    B = args.batch if args.batch != 32 else 100   # generate.py's default batch_size
    vae = LatticeDFCVAE(perceptual_model=None, device=dev, seed=1)
    vae._set_model(batch_size=B)
    unet = AtomUnet(device=dev, seed=2)
    pipe = GeneratePipeline(vae, unet, B, use_cuda_graph=not args.no_graph)
    rng = np.random.default_rng(3 + rank)
    zs = torch.from_numpy(rng.normal(0.0, 0.5, (B, 256)).astype(np.float32)).pin_memory()
    cond = torch.from_numpy(np.eye(10, dtype=np.float32)[rng.integers(0, 10, B)]).pin_memory()
    zd, cd = zs.to(dev), cond.to(dev)
    l0 = ops.launch_count()
    pipe.use_graph = False
    pipe.run(zd, cd)
    torch.cuda.synchronize()
    launches = ops.launch_count() - l0
    pipe.use_graph = not args.no_graph
    sampler = ClockSampler(local)
    ms = _timed_loop(lambda: pipe.run(zd, cd), args.steps, args.warmup, world, dev, sampler if rank == 0 else None)
    value = B * world / (ms * 1e-3)
    # e2e: z from pinned host memory, results the CPU tail needs (labels, mask, density channel, lattice, voxel) back D2H
    out = {k: None for k in ("species", "mask", "density", "lattice", "voxel")}
    host = {}

    def e2e_step():
        r = pipe.run(zs, cond)
        for k in out:
            t = r[k]
            if k not in host:
                host[k] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            host[k].copy_(t, non_blocking=True)
        torch.cuda.synchronize()

    e2e_s = _wall_loop(e2e_step, max(10, args.steps // 10), world, dev)
    d2h = sum(h.numel() * h.element_size() for h in host.values())
    clocks = sampler.stop() if rank == 0 else None
    pipe.use_graph = False
    roof = _conv_roofline(lambda: pipe.run(zd, cd), peaks, clocks)
    pipe.use_graph = not args.no_graph
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sps, _, cores = _cpu_inference(4, 2)
        cpu = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": "2 batches of 4 samples: oracle decoder + lattice params + U-Net forward + argmax/threshold"}
    if rank == 0:
        _emit(args, world, METRICS["inference"], value, ms, "bf16",
              {"workload": WORKLOADS["inference"], "grid": 32, "batch_per_gpu": B, "samples_total": B * world * args.steps,
               "parallelism": f"{world} independent replicas" if world > 1 else "single", "gflop_per_sample": INFER_GF,
               "l2": "a batch-100 pass moves > 126 MB per layer at 32^3; no flush needed", "cuda_graph": not args.no_graph},
              {"value": B * world / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": zs.numel() * 4 + cond.numel() * 4,
               "d2h_bytes_per_step": d2h}, launches * args.steps, clocks, roof, cpu,
              {"conv_tflops_whole_step": INFER_GF * value / 1e3}, scaling="weak")


def _run_voxeliser(args):
    import torch
    from icsg3d_b200 import ops, utils
    world, rank, local, dev = _dist()
    peaks = load_peaks()
    n = args.batch if args.batch != 32 else 4096
    sites, nsites, lat = utils.synthetic_cells(n, seed=5 + rank, device=dev)
    l0 = ops.launch_count()
    utils.voxelize_cells(sites, nsites, lat, d=32)
    torch.cuda.synchronize()
    launches = ops.launch_count() - l0
    sampler = ClockSampler(local)
    ms = _timed_loop(lambda: utils.voxelize_cells(sites, nsites, lat, d=32), args.steps, args.warmup, world, dev,
                     sampler if rank == 0 else None)
    value = n * world / (ms * 1e-3)
    hs, hn, hl = sites.cpu().pin_memory(), nsites.cpu().pin_memory(), lat.cpu().pin_memory()
    hm = torch.empty(n, 32, 32, 32, 4, dtype=torch.float32).pin_memory()
    hsp = torch.empty(n, 32, 32, 32, dtype=torch.uint8).pin_memory()

    def e2e_step():
        m32, _, s8, _ = utils.voxelize_cells(hs.to(dev, non_blocking=True), hn.to(dev, non_blocking=True),
                                             hl.to(dev, non_blocking=True), d=32)
        hm.copy_(m32, non_blocking=True)
        hsp.copy_(s8, non_blocking=True)
        torch.cuda.synchronize()

    e2e_s = _wall_loop(e2e_step, max(3, args.steps // 10), world, dev)
    clocks = sampler.stop() if rank == 0 else None
    gbs = VOX_BYTES * n / (ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "voxelize_fast_kernel (csrc/voxelize.cu; network outputs: fp32 input tensor + uint8 species)", "achieved": gbs, "peak": peaks["hbm_gbs"],
            "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": None,
            "algorithmic_bytes_per_launch": VOX_BYTES * n, "avg_launch_ms": ms,
            "note": "instruction-issue bound (ncu: issue slots 89 %, FP64 pipe 31 %, DRAM 13 %): the bit-exact species predicate needs ~100 fp64 instructions per (voxel, site); "
                    "profiles/r02_ncu_voxelize_fast_summary.txt"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sps, _, cores = _cpu_voxeliser(24)
        cpu = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": "24 cells through the numpy restatement of utils.density_matrix + coordinate_grid"}
    if rank == 0:
        _emit(args, world, METRICS["voxeliser"], value, ms, "f64",
              {"workload": WORKLOADS["voxeliser"], "grid": 32, "cells_per_launch": n, "sites_per_cell": 5,
               "parallelism": f"{world} independent replicas" if world > 1 else "single",
               "l2": "2.3 GB written per launch; no flush needed"},
              {"value": n * world / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": hs.numel() * 8 + hn.numel() * 4 + hl.numel() * 8,
               "d2h_bytes_per_step": hm.numel() * 4 + hsp.numel()}, launches * args.steps, clocks, roof, cpu)


def run(args):
    if args.impl == "reference":
        return run_reference(args)
    c = args.config
    if c == "inference" and args.steps == 200:
        args.steps = 1000   # the 100k-sample run of configs[4]
    if c in ("vae64", "unet64") and args.steps == 200:
        args.steps = 30
    if c == "unet_train":
        _run_unet_train(args, 32, args.batch if args.batch != 32 else 8)
    elif c == "unet64":
        _run_unet_train(args, 64, args.batch if args.batch != 32 else 16)
    elif c == "vae64":
        _run_vae64(args)
    elif c == "inference":
        _run_inference(args)
    else:
        _run_voxeliser(args)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
