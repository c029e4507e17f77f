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
        emit(({"impl": "reference", "metric": "volumes_per_sec", "value": v, "unit": "volumes/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * statistics.median(ts),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "slices_per_sec": v * 32, "config": cfg,
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
    x_host = synth.make_volume(B, D, args.img, args.img, seed=rank).pin_memory()
    x_dev = x_host.to(dev)

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

    for _ in range(max(3, args.warmup)):
        step(x_dev)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- device-resident timing ----
    l0 = model.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(x_dev)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    launches = (model.launch_count() - l0) + (2 * args.steps if args.saliency else 0)
    # ---- end-to-end timing: pinned host input -> H2D -> forward -> logits D2H, every step ----
    step(x_host).cpu()
    barrier()
    e0.record()
    for _ in range(args.steps):
        y = step(x_host)
        y_host = y.cpu()
    e1.record()
    barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    sampler.stop_flag = True
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    ms_step = ms.item() / args.steps
    ms_step_e2e = ms_e2e.item() / args.steps
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
        prof = model.profile_end()
    out = None
    if rank == 0:
        peaks = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "_src": "fallback"}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks.update(json.load(f)); peaks["_src"] = "measured"
        except Exception:
            pass
        roof = None
        kernels = {}
        traffic = {}
        try:  # DRAM bytes per launch from the committed ncu --set full capture of this workload (profiles/)
            with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as f:
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
            for k, (m_, n_) in prof.items():
                kernels[k] = {"ms_per_step": m_ / args.steps, "launches_per_step": n_ / args.steps, "share": (m_ / args.steps) / tot if tot else 0}
                if k in models and n_ and args.precision == "bf16":
                    fl, by = models[k]
                    ms_launch = m_ / n_   # (the last block's shorter launches are included in the average: < 1 % effect)
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
                            "traffic": traffic.get(top), "share_of_step": kt["share"],
                            "peak_kind": f"{peaks['_src']} sustained (kernel timed inside a long step)"}
                roof_top["frac"] = roof_top["achieved"] / roof_top["peak"]
                roof = dict(roof_top, family=roof)
        base = None
        if not args.no_cpu_baseline and world == 1:
            base, _ = cpu_baseline(3, 1)
        model_tflops = value * flops_per_volume(D, args.model, args.img) / 1e12
        out = {"metric": "volumes_per_sec", "value": value, "unit": "volumes/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": cfg,
               "slices_per_sec": value * D,
               "model_tflops": model_tflops, "model_frac_of_bf16_burst_peak": model_tflops / (peaks["bf16_tflops"] * world),
               "e2e": {"value": e2e, "unit": "volumes/s", "ms_per_step": ms_step_e2e,
                       "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": y_host.numel() * 4},
               "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": base, "kernels": kernels,
               "numa_node_rank0": numa}
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
