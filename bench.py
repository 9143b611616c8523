#!/usr/bin/env python
"""bench.py — headline benchmark of the d3fk hot path (contract in the task statement, tier ④).

Workload (BASELINE.json configs[1]): the d3f denoiser U-Net (smp resnet34 U-Net restated) at 64x64,
bf16 tensor-core compute, batch 256 per GPU: one full training step = q_sample + U-Net forward +
MSE/SSIM loss + U-Net backward + Adam, data-parallel over N GPUs with the NCCL gradient allreduce
overlapped with backward.  metric = train img/s (whole job).  The same JSON line carries the sampling
metric (img-steps/s, configs[2] shape: 128x128, batch 64 per GPU, CUDA-graph replayed steps) under
"sample", the roofline of the dominant kernel family (tcgen05 implicit-GEMM convolutions) under
"roofline", and the oracle's CPU step under "cpu_baseline".

`--impl reference` times the reference path's CPU restatement (oracle/) on the host cores."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_GFLOP_PER_IMG_64 = 0.97910          # SURVEY §8d (2*MAC, convs only)
STEM_DGRAD_GFLOP_64 = 0.01927
METRIC = "train_img_per_s"
UNIT = "img/s"


def synthetic_faces(B, H, W, seed, device):
    """Low-pass Gaussian field in [-1,1] (SURVEY §8d 'Synthetic inputs')."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    x = 0.5 * torch.randn(B, 3, H, W, generator=g, device=device)
    x = torch.nn.functional.avg_pool2d(x, 5, stride=1, padding=2) * 2.5
    return x.clamp(-1, 1).contiguous()


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def cpu_train_step_rate(batch, H, W, steps, warmup, lam=5.0):
    """The reference path restated (oracle/): q_sample + U-Net fwd/bwd + MSE/SSIM + Adam, fp32 eager on all
    host threads.  Returns (img/s, threads)."""
    import torch
    import oracle
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    model = oracle.Unet().train()
    crit = oracle.MseStructuralSimilarityLoss(-1.0, 1.0)
    opt = torch.optim.Adam(model.parameters(), lr=0.02)
    x = synthetic_faces(batch, H, W, 1234, "cpu")
    gen = torch.Generator().manual_seed(1)

    def step():
        noisy, _, _ = oracle.blend_random_amount_of_noise_with_each_sample(x, lam, gen)
        loss = crit(model(noisy), x)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, torch.get_num_threads(), dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_batch = args.batch       # the full per-GPU batch of configs[1]: one CPU step of 256 images takes < 1 s on 16 cores
    rate, threads, spt = cpu_train_step_rate(sample_batch, args.size, args.size, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": spt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "d3f denoiser train step, resnet34 U-Net 64x64, batch 256/GPU (configs[1])",
                   "global_batch": sample_batch, "parallelism": "cpu",
                   "sample": f"the whole step: batch {sample_batch} @{args.size}x{args.size} on the host CPU"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of batch {sample_batch} @{args.size}x{args.size} fp32 (oracle restatement; the "
                                   f"reference's own modules need smp/piqa/lightning, not installed)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cudnn_baseline(dev, B, H, sb, ss, steps, warmup):
    """"the existing Blackwell kernel to beat" = what `self.model(image_noisy)` executes today on this
    GPU — torch eager dispatching to cuDNN / ATen — for the SAME two workloads: the training step (noising + U-Net fwd/bwd +
    MSE/SSIM + Adam(fused=True)) at B x H x H and the eval forward at sb x ss x ss.  Two settings: bf16 autocast +
    channels_last (the speed comparison) and plain fp32 with TF32 off (the parity-mode comparison).  The network is the
    oracle restatement (test infrastructure; this leg is a reported baseline, never the product path)."""
    import torch
    import oracle
    torch.backends.cudnn.benchmark = True
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n, w):
        for _ in range(w):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for tag, amp in (("bf16_autocast_channels_last", True), ("fp32_tf32_off", False)):
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.manual_seed(0)
        model = oracle.Unet().to(dev).train()
        if amp:
            model = model.to(memory_format=torch.channels_last)
        crit = oracle.MseStructuralSimilarityLoss(-1.0, 1.0)
        opt = torch.optim.Adam(model.parameters(), lr=0.02, fused=True)
        gen = torch.Generator(device=dev).manual_seed(1)
        x = synthetic_faces(B, H, H, 1234, dev)
        if amp:
            x = x.contiguous(memory_format=torch.channels_last)

        def train_step():
            noisy, _, _ = oracle.blend_random_amount_of_noise_with_each_sample(x, 5.0, gen)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                pred = model(noisy)
            loss = crit(pred.float(), x)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()

        ms_train = timed(train_step, max(3, min(steps, 10)), max(3, warmup))
        model.eval()
        xe = synthetic_faces(sb, ss, ss, 99, dev)
        if amp:
            xe = xe.contiguous(memory_format=torch.channels_last)

        def eval_fwd():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                model(xe)

        ms_eval = timed(eval_fwd, 10, 5)
        out[tag] = {"train_ms_per_step": ms_train, "train_img_per_s": B / (ms_train / 1e3),
                    "eval_fwd_ms": ms_eval, "sample_img_steps_per_s": sb / (ms_eval / 1e3)}
        del model, opt
    torch.cuda.empty_cache()
    out["what"] = (f"oracle U-Net through torch {torch.__version__} eager (cuDNN {torch.backends.cudnn.version()}): train step "
                   f"B={B} @{H}x{H} incl. noising, MSE/SSIM loss and Adam(fused=True); eval forward B={sb} @{ss}x{ss}")
    return out


def swap_step_rate(dev, size, batch, steps, warmup, precision, seed):
    """configs[3]: the face-swap training batch of train_deep_fake (mode "swap": per batch 2 EMA updates, 2 no-grad EMA
    forwards, 2 forward/backward passes, 2 Adam steps — d3f/train_deep_fake/lit_module.py:142-156, :183-206) on the fast path.
    images/s counts both identities' images (2 * batch per step)."""
    import torch
    from denoising_diffusion_deep_fake_b200 import _lib
    from denoising_diffusion_deep_fake_b200.train import DeepFakeModule
    torch.manual_seed(seed)
    mod = DeepFakeModule(encoder_name="resnet34", learning_rate=0.02, noise_exponential_sampling_lambda=8, max_epochs=1,
                         cosine_scheduler_max_epoch=50, mode="swap", adam_b1=0.5, adam_b2=0.999, batch_size=batch,
                         ema_beta=0.9999, ema_update_every=10, precision=precision, seed=seed).to(dev).train()
    mod.configure_optimizers()
    xa, xb = synthetic_faces(batch, size, size, 500 + seed, dev), synthetic_faces(batch, size, size, 600 + seed, dev)
    for _ in range(max(3, warmup) + 10):
        mod.training_step(xa, xb)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = mod.training_step(xa, xb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res = {"batch_per_identity": batch, "size": size, "ms_per_step": ms, "img_per_s": 2 * batch / (ms / 1e3),
           "launches_per_step": (_lib.launch_count() - l0) / steps, "loss_a": float(out["a"]), "loss_b": float(out["b"])}
    del mod
    torch.cuda.empty_cache()
    return res


def sweep256(dev, model, batches, n_steps, seed):
    """configs[4]: sampling throughput at 256x256 over batch 1..512 (the reference U-Net has no attention — SURVEY §0 —
    so this is the same resnet34 U-Net at 256x256).  Small batches cannot fill 148 SMs: they are latency-bound and are
    flagged "report only"."""
    import torch
    from denoising_diffusion_deep_fake_b200.sampler import Sampler
    rows = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for b in batches:
        smp = Sampler(model, b, 256, 256, n_steps, r_start=1.0, eta=1.0, seed=seed, use_graph=True)
        smp.run()
        torch.cuda.synchronize()
        e0.record()
        smp.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_steps
        rows.append({"batch": b, "ms_per_step": ms, "img_steps_per_s": b / (ms / 1e3),
                     "tflops": b * FWD_GFLOP_PER_IMG_64 * 16 / ms,
                     "note": "latency-bound: report only" if b < 16 else ""})
        smp.plan.pending_backward = False
        del smp
        model._plans.clear()
        torch.cuda.empty_cache()
    return rows


def conv_only_oplist(plan):
    """Every convolution / weight-gradient launch of one training step as a stand-alone op list (the fused conv+BN ops
    contribute their convolution only: d3fk_convbn_params starts with the d3fk_conv_params)."""
    from denoising_diffusion_deep_fake_b200 import _lib
    import ctypes
    kinds = (_lib.OP_CONV, _lib.OP_WGRAD, _lib.OP_CONV_BN, _lib.OP_WGRAD_GROUP, _lib.OP_JOIN)   # JOIN: branch-lane bookkeeping
    ops = [op for op in plan.fwd_ops if op.kind in kinds]
    for seg in plan.bwd_segments or []:
        ops += [op for op in seg if op.kind in kinds]
    copies = []
    for op in ops:                       # detach from the plan's arrays
        c = _lib.Op()
        ctypes.memmove(ctypes.byref(c), ctypes.byref(op), ctypes.sizeof(_lib.Op))
        if c.kind == _lib.OP_CONV_BN:
            c.kind = _lib.OP_CONV
        copies.append(c)
    return _lib.OpList(copies), sum(1 for c in copies if c.kind != _lib.OP_JOIN)


def run_d3fk(args):
    import torch
    import torch.distributed as dist
    import denoising_diffusion_deep_fake_b200 as d3
    from denoising_diffusion_deep_fake_b200 import _lib
    from denoising_diffusion_deep_fake_b200.train import DenoiserModule
    from denoising_diffusion_deep_fake_b200.sampler import Sampler

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if args.main_priority:
        # the training step's main chain on a high-priority stream: its small latency-bound kernels get the next free CTA
        # slot ahead of the weight-gradient / optimiser kernels running on libd3fk's (lowest-priority) side streams
        torch.cuda.set_stream(torch.cuda.Stream(dev, priority=-args.main_priority))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, H, W = args.batch, args.size, args.size
    torch.manual_seed(0)
    mod = DenoiserModule(encoder_name="resnet34", learning_rate=0.02, noise_exponential_sampling_lambda=5,
                         cosine_scheduler_max_epoch=100, precision=args.precision, seed=1234 + rank).to(dev)
    mod.train()
    mod.configure_optimizers(fused=True)
    if world > 1:
        with torch.no_grad():                      # identical replicas
            dist.broadcast(mod.optimizer.flat_p, src=0)
        mod.enable_data_parallel()
    x = synthetic_faces(B, H, W, 1234 + rank, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ("value")
    # W untimed warm-up steps (at least 3) — and, because W steps of a 4 ms step are over before the GPU has settled at its
    # boost clocks, 64 more untimed steps (≈0.3 s; a fixed count, identical on every rank: each step is a collective).
    # Reported as "warmup_actual"; the timed region is untouched.
    n_warm = max(args.warmup, 3) + int(os.environ.get("D3FK_BENCH_EXTRA_WARMUP", "64"))   # (0 for ncu launch lists)
    for _ in range(n_warm):
        loss = mod.training_step(x)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    def kernels_so_far():      # libd3fk launches: issued directly + executed as nodes of the replayed step graph
        g = getattr(mod, "_graphed", None)
        return _lib.launch_count() + (g.kernels_replayed if g else 0)

    launches0 = kernels_so_far()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = mod.training_step(x)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = kernels_so_far() - launches0
    graphed = bool(getattr(mod, "_graphed", None)) and mod._graphed.replays > 0
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = world * B * args.steps / (ms / 1e3)
    final_loss = float(loss)

    # ---------------- end to end through the public API with host buffers ("e2e")
    # Every step copies its input batch from pinned host memory and reads its loss back to the host.  As a training
    # loop does, the copy of step i+1 is issued on a copy stream while step i computes (two device buffers), and the
    # loss of step i is read through a pinned buffer one step later (no host stall in the middle of the pipeline);
    # all K copies and all K reads happen inside the timed region.
    host = torch.empty((B, 3, H, W), dtype=torch.float32).pin_memory()
    host.copy_(x)
    xin = [torch.empty_like(x), torch.empty_like(x)]
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(dev)
    main_stream = torch.cuda.current_stream(dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]

    def issue_copy(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i & 1])          # the step that last read this buffer is done with it
            xin[i & 1].copy_(host, non_blocking=True)
            copied[i & 1].record(copy_stream)

    def run_e2e(n):
        losses = []
        for ev in consumed:
            ev.record(main_stream)
        issue_copy(0)
        for i in range(n):
            if i + 1 < n:
                issue_copy(i + 1)
            main_stream.wait_event(copied[i & 1])
            l = mod.training_step(xin[i & 1])
            consumed[i & 1].record(main_stream)
            loss_host[i & 1:(i & 1) + 1].copy_(l.detach().reshape(1), non_blocking=True)
            loss_ready[i & 1].record(main_stream)
            if i >= 1:
                loss_ready[(i - 1) & 1].synchronize()
                losses.append(float(loss_host[(i - 1) & 1]))          # device->host read of step i-1's result
        loss_ready[(n - 1) & 1].synchronize()
        losses.append(float(loss_host[(n - 1) & 1]))
        return losses

    run_e2e(6)           # warm-up: both input buffers get their step graph here, not inside the timed region
    barrier()
    e0.record()
    e2e_losses = run_e2e(args.steps)
    e1.record()
    barrier()
    assert len(e2e_losses) == args.steps
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item()
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    clocks.stop_flag = True
    clocks.join(timeout=2)
    if _lib.load().d3fk_device_error_flag():
        raise SystemExit("a d3fk kernel tripped its barrier watchdog inside the timed region: the numbers are void")

    # ---------------- roofline of the dominant kernel family: tcgen05 convolutions (fwd + dgrad + wgrad)
    plan = next(p for plans in mod.model._plans.values() for p in plans if p.training)
    conv_ops, n_conv = conv_only_oplist(plan)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        conv_ops.run(stream)
    torch.cuda.synchronize()
    reps = max(3, min(10, args.steps))
    e0.record()
    for _ in range(reps):
        conv_ops.run(stream)
    e1.record()
    torch.cuda.synchronize()
    conv_ms = e0.elapsed_time(e1) / reps
    scale = (H * W) / (64.0 * 64.0)
    conv_gflop = B * (3 * FWD_GFLOP_PER_IMG_64 - STEM_DGRAD_GFLOP_64) * scale
    achieved = conv_gflop / conv_ms            # GFLOP/ms == TFLOP/s
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    # DRAM bytes of the family per step from the committed ncu launch list (profiles/conv_traffic_r02.json, written by
    # tools/summarise_launches.py --json from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over the same
    # command; cold-cache per launch).  Only quoted for the workload it was captured on.
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "conv_traffic_r02.json")))
        if (B, H, W, args.precision) == (256, 64, 64, "bf16") and abs(tj["conv_family"]["launches_per_step"] - n_conv) < 0.5:
            traffic = tj["conv_family"]["dram_read_bytes_per_step"] + tj["conv_family"]["dram_write_bytes_per_step"]
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_unit": "bytes of DRAM traffic per step over the family's launches (ncu, profiles/conv_traffic_r02.json)",
                "kernel": "conv_tc_kernel + conv_slab_kernel + wgrad_tc_kernel + wgrad_slab_kernel (tcgen05 implicit GEMM)",
                "launches_per_step": n_conv, "ms_per_step": conv_ms,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                "share_of_step": conv_ms / (ms / args.steps)}

    # ---------------- sampling (configs[2] shape), batch-sharded, CUDA-graph replayed
    sample = None
    if not args.no_sample:
        sb, ss, n_steps = args.sample_batch, args.sample_size, args.sample_steps
        mod.model.eval()
        smp = Sampler(mod.model, sb, ss, ss, n_steps, r_start=1.0, eta=1.0, seed=7 + rank, use_graph=True,
                      chains=args.sample_chains, steps_per_graph=args.sample_steps_per_graph)
        smp.run()
        barrier()
        e0.record()
        smp.run()
        e1.record()
        barrier()
        sms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([sms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sms = t.item()
        sample = {"metric": "sample_img_steps_per_s", "value": world * sb * n_steps / (sms / 1e3), "unit": "img-steps/s",
                  "config": {"workload": f"{n_steps}-step DDPM sampling @{ss}x{ss}, batch {sb}/GPU, CUDA-graph replay (BASELINE configs[2])",
                             "chains": smp.chains, "steps_per_graph": smp.steps_per_graph},
                  "ms_per_step": sms / n_steps, "kernels_per_step": smp.kernels_per_step}
        mod.model.train()

    # ---------------- configs[3]: the face-swap training batch at 128x128 (reference batch 14, and 64)
    swap = None
    if not args.no_swap:
        swap = [swap_step_rate(dev, 128, b, max(5, min(args.steps, 20)), args.warmup, args.precision, 1234 + rank)
                for b in (14, 64)]
        if world > 1:
            for row in swap:
                t = torch.tensor([row["ms_per_step"]], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                row["ms_per_step"] = t.item()
                row["img_per_s"] = world * 2 * row["batch_per_identity"] / (t.item() / 1e3)
                row["note"] = f"{world} independent replicas (one identity pair per GPU), no communication"

    # ---------------- configs[4]: 256x256 sampling sweep (opt-in: --sweep256, or --workload sweep256)
    sweep = None
    if args.sweep256 or args.workload == "sweep256":
        mod.model.eval()
        sweep = sweep256(dev, mod.model, [1, 2, 4, 8, 16, 32, 64, 128, 256, 512], 10, 7 + rank)
        if world > 1:
            for row in sweep:
                t = torch.tensor([row["ms_per_step"]], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                row["ms_per_step"] = t.item()
                row["img_steps_per_s"] = world * row["batch"] / (t.item() / 1e3)
        mod.model.train()

    # ---------------- the library bar on the same GPU (rank 0, N = 1 only)
    cudnn = None
    if rank == 0 and world == 1 and not args.no_cudnn:
        cudnn = cudnn_baseline(dev, B, H, args.sample_batch, args.sample_size, args.steps, args.warmup)
        ref = cudnn["bf16_autocast_channels_last"]
        cudnn["d3fk_over_cudnn_bf16"] = {"train": value / ref["train_img_per_s"],
                                         "sample": (sample["value"] / ref["sample_img_steps_per_s"]) if sample else None}
    if _lib.load().d3fk_device_error_flag():
        raise SystemExit("a d3fk kernel tripped its barrier watchdog: the numbers are void")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, threads, spt = cpu_train_step_rate(8, 64, 64, steps=10, warmup=3)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "10 steps of batch 8 @64x64 fp32 (BASELINE configs[0]) through oracle/ on the host cores"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "warmup_actual": n_warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"d3f denoiser train step (q_sample+U-Net fwd/bwd+MSE/SSIM+Adam), resnet34 U-Net "
                                   f"{H}x{W}, batch {B}/GPU (BASELINE configs[1])",
                       "global_batch": world * B, "parallelism": f"dp{world}",
                       "l2": "per-step working set (activations+gradients > 1 GB) exceeds the 126 MB L2",
                       "step_launch": "CUDA graph replay" if graphed else "eager"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": host.numel() * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clocks.summary(), "roofline": roofline, "cpu_baseline": cpu,
            "sample": sample, "swap": swap, "sweep256": sweep, "cudnn_baseline": cudnn, "final_loss": final_loss,
        }
        if args.workload == "swap" and swap:
            row = swap[0]
            line.update(metric="swap_img_per_s", value=row["img_per_s"], ms_per_step=row["ms_per_step"],
                        config={"workload": "d3f train_deep_fake swap batch (2 EMA updates, 2 no-grad EMA forwards, 2 fwd/bwd, 2 Adam) "
                                            f"@128x128, batch {row['batch_per_identity']} per identity per GPU (BASELINE configs[3])",
                                "global_batch": world * 2 * row["batch_per_identity"], "parallelism": f"replicas{world}"},
                        train={"value": value, "ms_per_step": ms / args.steps}, e2e=None, roofline=roofline)
        if args.workload == "sweep256" and sweep:
            row = max(sweep, key=lambda r: r["img_steps_per_s"])
            line.update(metric="sample_img_steps_per_s", unit="img-steps/s", value=row["img_steps_per_s"], ms_per_step=row["ms_per_step"],
                        config={"workload": f"DDPM sampling @256x256, batch sweep 1-512 (best: {row['batch']}/GPU), CUDA-graph replay "
                                            "(BASELINE configs[4]; plain resnet34 U-Net — the reference has no attention)",
                                "global_batch": world * row["batch"], "parallelism": f"shard{world}"},
                        train={"value": value, "ms_per_step": ms / args.steps}, e2e=None)
        print(json.dumps(line), flush=True)
    if world > 1:
        # Teardown: the replayed step graphs hold NCCL work; release them before the communicator goes away, and never let a
        # stuck teardown keep the launcher waiting — the measurement is finished and printed at this point.
        import gc
        threading.Timer(45.0, lambda: os._exit(0)).start()
        g = getattr(mod, "_graphed", None)
        if g:
            g.entries.clear()
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="d3fk", choices=["d3fk", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--sample-batch", type=int, default=64)
    ap.add_argument("--sample-size", type=int, default=128)
    ap.add_argument("--sample-steps", type=int, default=1000, help="BASELINE configs[2]: 1000-step DDPM sampling")
    ap.add_argument("--sample-chains", type=int, default=None, help="sub-batches sampled as parallel graph branches")
    ap.add_argument("--sample-steps-per-graph", type=int, default=int(os.environ.get("D3FK_STEPS_PER_GRAPH", "1")))
    ap.add_argument("--no-sample", action="store_true")
    ap.add_argument("--no-swap", action="store_true")
    ap.add_argument("--no-cudnn", action="store_true")
    ap.add_argument("--sweep256", action="store_true")
    ap.add_argument("--workload", default="train", choices=["train", "swap", "sweep256"],
                    help="which measurement becomes the headline metric/value of the JSON line (train = BASELINE configs[1])")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--main-priority", type=int, default=int(os.environ.get("D3FK_MAIN_PRIORITY", "0")))
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_d3fk(args)


if __name__ == "__main__":
    main()
