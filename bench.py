"""MST-DINOv2 forward throughput (BASELINE.json metric: volumes/s and slices/s, % of bf16 tensor peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 64] [--saliency]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one forward of a batch of synthetic volumes (config 2: 64 x 32 x 224 x 224, bf16, random-init
ViT-S/14) per GPU.  Volumes shard whole across GPUs (weak scaling, no data-path collective; logits are
all-gathered once per step).  Prints ONE JSON line on rank 0.
"""
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

# algorithmic matmul FLOPs, ViT-S/14 @224, 32 slices
FLOP_PER_SLICE_S = 12_247_123_968
FLOP_SLICE_TRANSFORMER_S = 60_066_816


def flops_per_volume(D=32, model="s", img=224):
    """Algorithmic matmul FLOPs of one volume (2*M*N*K, matmuls only)."""
    if model == "s" and img == 224:
        return FLOP_PER_SLICE_S * D + (FLOP_SLICE_TRANSFORMER_S if D == 32 else slice_transformer_flops(384, D))
    E = {"s": 384, "b": 768}[model]
    P = (img // 14) ** 2
    N = P + 1
    per_block = 2 * N * E * (3 * E + E + 4 * E + 4 * E) + 4 * (E // 64) * N * N * 64
    return (2 * P * 588 * E + 12 * per_block) * D + slice_transformer_flops(E, D)


def slice_transformer_flops(E, D):
    L = D + 1   # in-proj, QK^T, PV, out-proj, two FFN matrices (dim_feedforward = E), head
    return 2 * L * E * 3 * E + 4 * L * L * E + 2 * L * E * E + 4 * L * E * E + 2 * E * 2


def gemm_flops_executed(BD, N=257, E=384, depth=12, KP=256):
    """Tensor-core FLOPs the tcgen05 GEMM launches of one forward really execute (last block runs proj/MLP on the
    CLS rows only; the patch GEMM runs with K padded to 256 on channel-summed weights)."""
    M = BD * N
    per_layer_full = 2 * M * E * (3 * E + E + 4 * E + 4 * E)
    last = 2 * M * E * 3 * E + 2 * BD * E * (E + 4 * E + 4 * E)
    return 2 * BD * (N - 1) * KP * E + (depth - 1) * per_layer_full + last


def kernel_models(BD, N=257, E=384, heads=6):
    """Algorithmic FLOPs and HBM bytes PER LAUNCH of the kernels of one full encoder block:
    bf16 activations read/written once, the in-place residual update counted as read + write, weights negligible."""
    M = BD * N
    return {
        "gemm_qkv": (2 * M * 3 * E * E, M * E * 2 + M * 3 * E * 2 + M * 4),
        "gemm_proj": (2 * M * E * E, M * E * 2 + 2 * M * E * 2),
        "gemm_fc1": (2 * M * 4 * E * E, M * E * 2 + M * 4 * E * 2 + M * 4),
        "gemm_fc2": (2 * M * 4 * E * E, M * 4 * E * 2 + 2 * M * E * 2),
        "attention": (4 * BD * heads * N * N * 64, M * 3 * E * 2 + M * E * 2),
        "layernorm": (0, M * E * 2 + M * 4),   # bf16 path: row statistics only (the normalisation rides in the next GEMM)
    }


def map_kernel_models(B, D, heads, N, H, W, sheads=12):
    """Algorithmic HBM bytes per launch of the saliency kernels: the combiner reads the CLS rows
    [BD, heads, N] + [B, 12, D+1] fp32 and writes the coarse map; the upsampler reads the coarse map and writes
    6 422 528 B per 32x224x224 volume ([B,1,D,H,W] fp32)."""
    P = N - 1
    return {
        "saliency_combine": (0, B * D * heads * N * 4 + B * sheads * (D + 1) * 4 + B * D * P * 4),
        "saliency_upsample": (0, B * D * P * 4 + B * D * H * W * 4),
    }


class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run this rank on the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned input buffer is
    allocated (first touch puts the pages there), so the per-step host-to-device copies of 8 ranks do not all cross the
    socket interconnect.  Best effort: returns the node, or None when sysfs does not say."""
    try:
        import torch
        bdf = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        if bdf is None:
            o = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                               capture_output=True, text=True, timeout=10).stdout.strip()
            bdf = o
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:      # nvidia-smi prints an 8-digit domain, sysfs uses 4
            bdf = bdf[4:]
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


CPU_SAMPLE_VOLUMES = 4   # volumes per CPU step: batching helps ATen's threading a little (about +15 % over one volume here)


def cpu_baseline(steps, warmup, threads=None):
    """Time the oracle (CPU port of the reference path) on a bounded sample of the workload: CPU_SAMPLE_VOLUMES synthetic
    volumes per step, all host threads."""
    import torch
    from new_vit_b200 import synth
    from oracle import mst_oracle as O
    torch.set_num_threads(threads or os.cpu_count())
    sd = synth.make_state_dict("s", 2, seed=0)
    nv = CPU_SAMPLE_VOLUMES
    x = synth.make_volume(nv, 32, 224, 224, seed=0)
    for _ in range(warmup):
        O.forward(sd, x)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.forward(sd, x)
        ts.append(time.perf_counter() - t0)
    return {"value": nv / statistics.median(ts), "unit": "volumes/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} x {nv} volumes 32x224x224 fp32 (oracle/mst_oracle.py, torch CPU), median; min {min(ts):.3f}s per step"}, ts


def load_peaks():
    peaks = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "_src": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks.update(json.load(f)); peaks["_src"] = "measured"
    except Exception:
        pass
    return peaks


def run_extras(model_s, dev, rank, world, timed, peaks):
    """Short runs of the other configurations on the same box, every rank (values are whole-job, max time over ranks):
    config 3 (save_attn + full-resolution saliency volume, 256 volumes sharded over the ranks, at least 32 per GPU), config 4
    (ViT-B/14, 64 slices x 252 x 252, 16 volumes per GPU), the reference's own predict loop (batch 1, with and without the 8-flip
    TTA, main_predict.py:208,288) as a latency, and the input pipeline (mst_prepare_volume)."""
    import torch
    from new_vit_b200 import DinoV2ClassifierSlice, synth
    from new_vit_b200.model import run_pred
    from new_vit_b200.transforms import duke_transform
    out = {}
    hbm = (peaks or {}).get("hbm_gbs", 1.0)
    # ---- config 3 ----
    B3, D, H = max(32, 256 // world), 32, 224
    x3 = synth.make_volume(B3, D, H, H, seed=100 + rank).to(dev)

    def sal():
        with torch.no_grad():
            model_s(x3, save_attn=True)
            model_s.saliency_volume()
    for _ in range(3):
        sal()
    ms = timed(sal, 5)
    model_s.profile_begin()
    for _ in range(5):
        sal()
    prof = model_s.profile_end()
    mk = map_kernel_models(B3, D, 6, 257, H, H)
    c3 = {"workload": f"config 3: save_attn forward + saliency volume [B,1,32,224,224] fp32, {B3} volumes per GPU x {world} GPU",
          "value": B3 * world / (ms * 1e-3), "unit": "volumes/s", "ms_per_step": ms}
    for k in ("saliency_combine", "saliency_upsample"):
        t, n = prof[k]
        if n:
            c3[k] = {"ms_per_launch": t / n, "algorithmic_bytes_per_launch": mk[k][1], "gbs": mk[k][1] / (t / n) / 1e6,
                     "frac_of_hbm_peak": mk[k][1] / (t / n) / 1e6 / hbm}
    out["config3_saliency"] = c3
    del x3
    # ---- the reference's predict loop: one volume per call (main_predict.py:208), device time per call ----
    x1 = synth.make_volume(1, D, H, H, seed=7).to(dev)
    with torch.no_grad():
        for _ in range(5):
            model_s(x1)
        lat = timed(lambda: model_s(x1), 50)
        batch = {"source": x1}
        for _ in range(3):
            run_pred(model_s, batch, save_attn=True, use_tta=True)
        lat_tta = timed(lambda: run_pred(model_s, batch, save_attn=True, use_tta=True), 20)
        lat_sal = timed(lambda: run_pred(model_s, batch, save_attn=True, use_tta=False), 20)
    out["latency_batch1"] = {"workload": "1 volume 32x224x224 per call, device-resident (scripts/main_predict.py:208,288)",
                             "cuda_graph_replays": model_s.graph_replays(),
                             "forward_ms": lat, "run_pred_saliency_ms": lat_sal, "run_pred_saliency_tta8_ms": lat_tta,
                             "forward_model_tflops": flops_per_volume(32) / lat / 1e9}
    # ---- input pipeline: 64 x (256x256x40 -> 224x224x32), mst_prepare_volume ----
    raw = torch.randn(64, 256, 256, 40, device=dev) * 300 + 500
    for _ in range(3):
        duke_transform(raw, (224, 224, 32), check=False)
    ms = timed(lambda: duke_transform(raw, (224, 224, 32), check=False), 10)
    by = 64 * (224 * 224 * 32) * 4 * 2    # one read of the kept region + one write of the result
    out["prepare_volume"] = {"workload": "64 x (256x256x40 fp32 -> crop 224x224x32, percentile-clamped z-norm, axis swap)",
                             "ms_per_step": ms, "volumes_per_sec": 64 / (ms * 1e-3), "algorithmic_bytes": by,
                             "gbs": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / hbm}
    del raw
    # ---- config 5: training step = forward + CrossEntropy + backward + gradient all-reduce (NCCL, when world > 1) + AdamW,
    #      8 volumes per GPU (base_model.py:148-170,103-110; main_train.py:110-126).  Two constructions: every parameter trainable
    #      (the default, what main_train.py trains) and freeze=True (dino.py:69-71: slice transformer + head only) ----
    for key, frozen in (("config5_train", False), ("config5_train_frozen_encoder", True)):
        mt = DinoV2ClassifierSlice(1, 2, pretrained=False, precision="bf16", freeze=frozen).to(dev)
        mt.load_state_dict(synth.make_state_dict("s", 2, seed=0))
        mt.train()
        opt = mt.configure_optimizers()[0]
        xt = synth.make_volume(8, D, H, H, seed=300 + rank).to(dev)
        batch = {"source": xt, "target": torch.randint(0, 2, (8,), device=dev), "uid": ["x"] * 8}

        def train_step():
            opt.zero_grad()
            loss = mt.training_step(batch, 0)
            loss.backward()
            opt.step()
            return loss
        for _ in range(3):
            train_step()
        ms = timed(train_step, 10)
        mt.profile_begin()
        for _ in range(3):
            train_step()
        prof = mt.profile_end()
        trainable = sum(p.numel() for p in mt.parameters() if p.requires_grad)
        # forward = 1x the algorithmic matmul FLOPs, backward = 2x (dgrad + wgrad) for what trains
        fl = flops_per_volume(32) * (1 if frozen else 3) * 8
        out[key] = {
            "workload": f"config 5{' (frozen encoder)' if frozen else ''}: training step fwd + CrossEntropy + bwd + "
                        f"{'NCCL gradient all-reduce + ' if world > 1 else ''}AdamW, bf16, 8 volumes x 32 x 224x224 per GPU x {world} GPU; "
                        f"{trainable} trainable parameters" + ("; the encoder runs forward-only" if frozen else " (all of them)"),
            "value": 8 * world / (ms * 1e-3), "unit": "volumes/s", "ms_per_step": ms, "trainable_parameters": trainable,
            "grad_allreduce_bytes": trainable * 4 if world > 1 else 0, "loss": float(train_step().detach()),
            "model_tflops": fl * world / (ms * 1e-3) / 1e12,
            "ms_by_category": {k: round(v[0] / 3, 3) for k, v in prof.items() if v[1]}}
        if world > 1:
            # every rank saw different volumes and the same all-reduced gradient: the parameters must still be bit-identical
            chk = torch.stack([p.detach().double().sum() for p in mt.parameters() if p.requires_grad]).sum().reshape(1)
            allc = [torch.zeros_like(chk) for _ in range(world)]
            torch.distributed.all_gather(allc, chk)
            out[key]["ranks_in_sync"] = bool(all(float(c) == float(allc[0]) for c in allc))
        del mt, opt, xt, batch
        torch.cuda.empty_cache()
    # ---- config 4 ----
    torch.cuda.empty_cache()
    mb = DinoV2ClassifierSlice(1, 2, pretrained=False, precision="bf16", model_size="b", img_size=252).to(dev).eval()
    mb.load_state_dict(synth.make_state_dict("b", 2, seed=0, img_size=252))
    x4 = synth.make_volume(16, 64, 252, 252, seed=200 + rank).to(dev)
    with torch.no_grad():
        for _ in range(3):
            mb(x4)
        ms = timed(lambda: mb(x4), 5)
    v4 = 16 * world / (ms * 1e-3)
    tf = v4 * flops_per_volume(64, "b", 252) / 1e12
    out["config4_vitb"] = {"workload": f"config 4: ViT-B/14, 16 volumes x 64 slices x 252x252 per GPU x {world} GPU (256x256 is rejected by "
                                       "the reference, patch_embed.py:72-73)", "value": v4, "unit": "volumes/s", "ms_per_step": ms,
                           "model_tflops": tf, "model_frac_of_bf16_burst_peak": tf / ((peaks or {}).get("bf16_tflops", 1.0) * world) if peaks else None}
    del mb, x4
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="volumes per GPU per step")
    ap.add_argument("--slices", type=int, default=32)
    ap.add_argument("--model", default="s", choices=["s", "b"], help="encoder size (config 4: b)")
    ap.add_argument("--img", type=int, default=224, help="slice height = width (config 4: 252; 256 is rejected by the reference)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--saliency", action="store_true", help="config 3: save_attn + full-resolution saliency volume")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short config-3 / config-4 / batch-1 / input-pipeline runs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    # The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner on the first
    # communicator), so file descriptor 1 is pointed at stderr for the run and the result goes to the saved descriptor.
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(result_fd, (json.dumps(obj) + "\n").encode())
    cfg = {"workload": f"MST-DINOv2 (DinoV2ClassifierSlice, random-init ViT-{args.model.upper()}/14) {args.precision} inference, "
                       f"{args.batch} volumes x {args.slices} slices x {args.img}x{args.img} per GPU"
                       + (" + --get_attention saliency maps" if args.saliency else ""),
           "volumes_per_gpu": args.batch, "slices": args.slices, "img": args.img, "parallelism": f"volume-sharded dp{world}",
           "l2_policy": "inputs larger than L2 (411 MB fp32 per step; 0.4-1.6 GB activations per kernel)"}

    if args.impl == "reference":
        # The reference is a Python/PyTorch package that cannot travel to the GPU box; its path is timed through
        # the oracle port on the host cores (bounded sample: CPU_SAMPLE_VOLUMES volumes per step).
        if rank != 0:
            return
        base, ts = cpu_baseline(max(1, args.steps), max(1, args.warmup))
        v = base["value"]
        # `config` stays the GPU arm's (the driver matches the two arms on it); what THIS arm runs per step is a bounded sample of
        # that workload, stated in `reference_arm_sample` and in cpu_baseline.sample
        sample_note = (f"{CPU_SAMPLE_VOLUMES} volumes x 32 x 224 x 224 fp32 per step on the host cores (CPU port of the reference "
                       "path, kind 'port'); the GPU-over-CPU ratio is a GPU-versus-CPU-port figure, not a same-device comparison")
        emit(({"impl": "reference", "metric": "volumes_per_sec", "value": v, "unit": "volumes/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * statistics.median(ts),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "slices_per_sec": v * 32, "config": cfg, "reference_arm_sample": sample_note,
                          "cpu_baseline": base,
                          "e2e": {"value": v, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from new_vit_b200 import DinoV2ClassifierSlice, synth
    from new_vit_b200.dist import gather_volumes
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, D = args.batch, args.slices
    torch.manual_seed(0)
    model = DinoV2ClassifierSlice(1, 2, pretrained=False, precision=args.precision, model_size=args.model, img_size=args.img).to(dev).eval()
    model.load_state_dict(synth.make_state_dict(args.model, 2, seed=0, img_size=args.img))
    x_host32 = synth.make_volume(B, D, args.img, args.img, seed=rank).pin_memory()
    # what travels host -> device end to end: the bf16 path rounds every voxel to bf16 before the patch GEMM anyway, so the
    # loader hands over bf16 volumes (bit-identical results, half the bytes); fp32 mode keeps fp32
    x_host = x_host32.to(torch.bfloat16).pin_memory() if args.precision == "bf16" else x_host32
    x_dev = x_host32.to(dev)

    def step(src):
        with torch.no_grad():
            y = model(src, save_attn=args.saliency)
            if args.saliency:
                model.saliency_volume()
            if world > 1:
                y = gather_volumes(y, B * world)  # only the [B,2] logits cross NVLink
        return y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """device time of `steps` calls of fn (CUDA events on the launching stream), max over ranks, in ms per call"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / steps

    for _ in range(max(3, args.warmup)):
        step(x_dev)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- device-resident timing ----
    l0 = model.launch_count()
    ms_step = timed(lambda: step(x_dev), args.steps)
    launches = model.launch_count() - l0
    # ---- end-to-end timing: pinned host input -> H2D -> forward -> logits D2H, every step ----
    y_host = step(x_host).cpu()

    def e2e_step(src):
        nonlocal y_host
        y_host = step(src).cpu()
    # the same W warm-up steps as the device-resident leg: the first DMA passes over a freshly pinned buffer (and the first uses of the
    # staging buffers / copy stream) are slower on some hosts; one warm-up call left the bf16 leg 20 % slow on two of seven boxes
    for _ in range(max(args.warmup, 3)):
        e2e_step(x_host)
    ms_step_e2e = timed(lambda: e2e_step(x_host), args.steps)
    ms_step_e2e32 = None
    if x_host is not x_host32:
        for _ in range(max(args.warmup, 3)):
            e2e_step(x_host32)
        ms_step_e2e32 = timed(lambda: e2e_step(x_host32), args.steps)
    sampler.stop_flag = True
    vols = B * world
    value = vols / (ms_step * 1e-3)
    e2e = vols / (ms_step_e2e * 1e-3)

    # ---- per-kernel device time (CUDA events on the launch stream, separate profiled pass of the same steps) ----
    prof = None
    if not args.no_profile and rank == 0:
        model.profile_begin()
        for _ in range(args.steps):
            with torch.no_grad():
                model(x_dev, save_attn=args.saliency)
                if args.saliency:
                    model.saliency_volume()
        prof = model.profile_end()
    out = None
    peaks = load_peaks()
    if rank == 0:
        roof = None
        kernels = {}
        traffic = {}
        try:  # DRAM bytes per launch from the committed ncu --set full capture of this workload (profiles/)
            with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as f:
                traffic = json.load(f)
        except Exception:
            pass
        if prof:
            gemm_cats = ["gemm_patch", "gemm_qkv", "gemm_proj", "gemm_fc1", "gemm_fc2", "gemm_cls_rows"]
            gemm_ms = sum(prof[c][0] for c in gemm_cats) / args.steps
            gemm_n = sum(prof[c][1] for c in gemm_cats) // args.steps
            fl = gemm_flops_executed(B * D, N=1 + (args.img // 14) ** 2, E={'s': 384, 'b': 768}[args.model])
            ach = fl / (gemm_ms * 1e-3) / 1e12
            peak = peaks["bf16_tflops_sustained"]
            roof = {"kernel": "gemm_tc_kernel (tcgen05/TMA/TMEM bf16 GEMM family, %d launches/step)" % gemm_n, "bound": "tensor",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "peak_kind": f"{peaks['_src']} sustained (kernel timed inside a long step); burst {peaks['bf16_tflops']}",
                    "frac_of_burst": ach / peaks["bf16_tflops"], "flops_per_step": fl, "ms_per_step": gemm_ms, "traffic": None}
            tot = sum(v[0] for v in prof.values()) / args.steps
            models = kernel_models(B * D, N=1 + (args.img // 14) ** 2, E={'s': 384, 'b': 768}[args.model], heads={'s': 6, 'b': 12}[args.model])
            models.update(map_kernel_models(B, D, {'s': 6, 'b': 12}[args.model], 1 + (args.img // 14) ** 2, args.img, args.img))
            for k, (m_, n_) in prof.items():
                kernels[k] = {"ms_per_step": m_ / args.steps, "launches_per_step": n_ / args.steps, "share": (m_ / args.steps) / tot if tot else 0}
                if k in models and n_ and args.precision == "bf16":
                    fl, by = models[k]
                    # qkv of ViT-S is two kernels per block (features 0..1023 on gemm_wt + the 128-feature tail): one logical launch
                    logical = n_ / 2 if (k == "gemm_qkv" and n_ / args.steps > 1.5 * 12) else n_
                    ms_launch = m_ / logical   # (the last block's shorter launches are included in the average: < 1 % effect)
                    t_tensor, t_hbm = fl / (peaks["bf16_tflops_sustained"] * 1e12), by / (peaks["hbm_gbs"] * 1e9)
                    kernels[k].update({"ms_per_launch": ms_launch, "tflops": fl / ms_launch / 1e9, "gbs": by / ms_launch / 1e6,
                                       "bound": "tensor" if t_tensor >= t_hbm else "hbm",
                                       "frac_of_roofline": max(t_tensor, t_hbm) * 1e3 / ms_launch})
            kernels["_sum_ms_per_step"] = tot
            # the dominant single kernel (largest share of the step) against the roofline that bounds it
            top = max((k for k in kernels if k in models and "bound" in kernels[k]), key=lambda k: kernels[k]["share"], default=None)
            if top:
                kt = kernels[top]
                fl, by = models[top]
                tensor = kt["bound"] == "tensor"
                roof_top = {"kernel": top, "bound": kt["bound"], "achieved": kt["tflops"] if tensor else kt["gbs"],
                            "peak": peaks["bf16_tflops_sustained"] if tensor else peaks["hbm_gbs"],
                            "unit": "TFLOP/s" if tensor else "GB/s", "ms_per_launch": kt["ms_per_launch"],
                            "algorithmic_flops_per_launch": fl, "algorithmic_bytes_per_launch": by,
                            "traffic": traffic.get(top), "traffic_src": "static: profiles/r02_ncu_traffic.json (one ncu --set full capture of this "
                            "workload; refreshed by profiles/ncu_traffic.py, not measured in this run)", "share_of_step": kt["share"],
                            "peak_kind": f"{peaks['_src']} sustained (kernel timed inside a long step)"}
                roof_top["frac"] = roof_top["achieved"] / roof_top["peak"]
                roof = dict(roof_top, family=roof)
        base = None
        if not args.no_cpu_baseline and world == 1:
            base, _ = cpu_baseline(3, 1)
    # ---- the other BASELINE.json configs and the reference's own predict loop, short runs on the same box (extra keys) ----
    extras = None
    if not args.no_extras and not args.saliency and args.model == "s" and args.precision == "bf16":
        extras = run_extras(model, dev, rank, world, timed, peaks if rank == 0 else None)
    if rank == 0:
        model_tflops = value * flops_per_volume(D, args.model, args.img) / 1e12
        out = {"metric": "volumes_per_sec", "value": value, "unit": "volumes/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": cfg,
               "slices_per_sec": value * D,
               "model_tflops": model_tflops, "model_frac_of_bf16_burst_peak": model_tflops / (peaks["bf16_tflops"] * world),
               "e2e": {"value": e2e, "unit": "volumes/s", "ms_per_step": ms_step_e2e, "source_dtype": str(x_host.dtype),
                       "h2d_bytes_per_step": x_host.numel() * x_host.element_size(), "d2h_bytes_per_step": y_host.numel() * 4,
                       "fp32_source": None if ms_step_e2e32 is None else
                       {"value": vols / (ms_step_e2e32 * 1e-3), "ms_per_step": ms_step_e2e32, "h2d_bytes_per_step": x_host32.numel() * 4}},
               "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": base, "kernels": kernels,
               "numa_node_rank0": numa, "extras": extras}
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
