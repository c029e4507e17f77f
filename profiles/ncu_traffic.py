"""Turn an `ncu --set full` report into the per-kernel table under profiles/ and the DRAM-traffic JSON bench.py quotes as
`roofline.traffic` (static: refreshed by running this script on a new capture, not measured inside bench.py).

    ncu -i gpurun_out/<name>.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/ncu_traffic.py /tmp/raw.csv profiles/r02_top_kernels.md [profiles/r02_ncu_traffic.json]
"""
import csv
import json
import sys

COLS = [("gpu__time_duration.sum", "duration us", 1e-3 if False else 1.0),
        ("dram__bytes_read.sum", "DRAM read", 1.0), ("dram__bytes_write.sum", "DRAM write", 1.0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM thr %", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 1.0),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
        ("launch__registers_per_thread", "regs", 1.0), ("launch__grid_size", "grid", 1.0), ("launch__block_size", "block", 1.0)]

CATEGORY = [("gemm_wt_kernel<5>", "gemm_qkv_main_1024"), ("gemm_tc_kernel<128, 0, 5, 3, 8, 0>", "gemm_qkv_tail_128"), ("gemm_wt_kernel<6>", "gemm_fc1"),
            ("gemm_tc_kernel<192, 0, 7, 1, 8, 2>", "gemm_fc2"), ("gemm_tc_kernel<192, 6, 4, 1, 8, 1>", "gemm_proj"), ("attention_tc257", "attention"),
            ("row_stats", "layernorm"), ("saliency_combine", "saliency_combine"), ("saliency_upsample", "saliency_upsample")]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    head, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(head)}
    name_i = idx["Kernel Name"]
    out = ["| kernel | " + " | ".join(c[1] for c in COLS) + " |", "|---|" + "---|" * len(COLS)]
    traffic = {}
    for r in data:
        if len(r) <= name_i:
            continue
        cells = []
        rd = wr = 0.0
        for key, label, _ in COLS:
            if key not in idx:
                cells.append("-")
                continue
            v, u = r[idx[key]], units[idx[key]]
            if "bytes" in key:
                b = to_bytes(v, u)
                cells.append(f"{b / 1e6:.1f} MB")
                if key.endswith("read.sum"):
                    rd = b
                else:
                    wr = b
            elif key.startswith("gpu__time"):
                t = float(v.replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "second": 1e6}.get(u, 1.0)
                cells.append(f"{t:.1f}")
            else:
                cells.append(v)
        short = r[name_i].replace("void mst::", "").split("(")[0][:70]
        out.append(f"| {short} | " + " | ".join(cells) + " |")
        for pat, cat in CATEGORY:
            if pat in r[name_i] and cat not in traffic:
                traffic[cat] = int(rd + wr)
    open(sys.argv[2], "w").write("\n".join(out) + "\n")
    if len(sys.argv) > 3:
        traffic["_source"] = "ncu --set full --clock-control none (profiles/prof_forward.py), dram__bytes_read.sum + dram__bytes_write.sum of the first launch of each kernel; written by profiles/ncu_traffic.py"
        if "gemm_qkv_main_1024" in traffic:
            traffic["gemm_qkv"] = traffic["gemm_qkv_main_1024"] + traffic.get("gemm_qkv_tail_128", 0)
        json.dump(traffic, open(sys.argv[3], "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main()
