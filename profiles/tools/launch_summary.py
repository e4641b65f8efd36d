"""Summarises an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`)
per kernel and per kernel family, and writes the per-family DRAM-traffic JSON that bench.py reports as
`roofline.traffic`.   python profiles/tools/launch_summary.py LAUNCHES.csv [--json OUT.json] [--md]
"""
import argparse
import csv
import json
import re
from collections import defaultdict

# kernel-name pattern -> family name used by change3d_b200/ops.py's per-launch profiler
FAMILIES = [
    (r"pw_gemm_tc_kernel|pw_gemm_kernel", "pw_gemm"), (r"pw_wgrad", "pw_wgrad"),
    (r"dw_fwd", "dw_conv_fwd"), (r"dw_bwd|dw_dy_kernel", "dw_conv_bwd"), (r"relu_bwd_stats", "relu_bwd_stats"),
    (r"bn_add_relu", "bn_add_relu"), (r"stem_bwd", "stem_bwd"), (r"stem_fwd", "stem_fwd"),
    (r"convt_col2im", "convt_col2im"), (r"convt_im2col", "convt_im2col"), (r"dec_head_fwd", "dec_head_fwd"),
    (r"dec_head_bwd", "dec_head_bwd"), (r"bce_dice_fwd|ce2d_fwd|sim_kernel<\d+, false|sim_kernel<\(int\)\d+, \(bool\)0", "loss_fwd"),
    (r"bce_dice_bwd|ce2d_bwd|sim_kernel", "loss_bwd"), (r"finalize|se_bn_bwd_fast|se_bn_bwd_cluster", "finalizers"), (r"adam_kernel", "adam"),
]


def family(name: str) -> str:
    for pat, fam in FAMILIES:
        if re.search(pat, name):
            return fam
    return "torch / other"


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    return name if len(name) < 90 else name[:87] + "..."


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--json")
    ap.add_argument("--md", action="store_true")
    ap.add_argument("--by-shape", metavar="REGEX",
                    help="also list the launches whose kernel name matches REGEX grouped by (kernel, grid, DRAM MB)")
    ap.add_argument("--one-step", action="store_true",
                    help="the launch sequence of a training loop is periodic: keep exactly one period (one step)")
    a = ap.parse_args()
    rows = defaultdict(dict)
    names = {}
    grids = {}
    with open(a.csv) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        rows[r["ID"]][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * \
            ({"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"], 1.0) if "time" in r["Metric Name"] else
             {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0))
        names[r["ID"]] = r["Kernel Name"]
        grids[r["ID"]] = r["Grid Size"]
    if a.one_step:
        ids = sorted(rows, key=int)
        seq = [names[i] for i in ids]
        period = next((L for L in range(50, len(seq)) if all(seq[j] == seq[j + L] for j in range(len(seq) - L))), None)
        if period is None:
            raise SystemExit("no period found: the window does not contain more than one step")
        print(f"period = {period} launches per step (window of {len(seq)})")
        rows = {i: rows[i] for i in ids[:period]}
    per_kernel, per_family = defaultdict(lambda: [0, 0.0, 0.0]), defaultdict(lambda: [0, 0.0, 0.0])
    for i, m in rows.items():
        t = m.get("gpu__time_duration.sum", 0.0)
        b = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        for table, key in ((per_kernel, short(names[i])), (per_family, family(names[i]))):
            table[key][0] += 1
            table[key][1] += t
            table[key][2] += b
    total = sum(v[1] for v in per_kernel.values())
    print(f"{len(rows)} launches, {total:.2f} ms of serialised kernel time")
    import signal
    signal.signal(signal.SIGPIPE, signal.SIG_DFL)
    for table in (per_family, per_kernel):
        print("| share | ms | launches | DRAM GB | DRAM GB/s | name |\n|---|---|---|---|---|---|" if a.md else "")
        for k, (n, t, b) in sorted(table.items(), key=lambda kv: -kv[1][1])[:24]:
            gbs = b / (t * 1e-3) / 1e9 if t > 0 else 0.0
            if a.md:
                print(f"| {100 * t / total:.1f} % | {t:.2f} | {n} | {b / 1e9:.1f} | {gbs:.0f} | `{k}` |")
            else:
                print(f"{100 * t / total:5.1f}%  {t:8.2f} ms  {n:4d}  {b / 1e9:7.2f} GB  {gbs:6.0f} GB/s  {k}")
    if a.by_shape:
        shapes = defaultdict(lambda: [0, 0.0, 0.0])
        for i, m in rows.items():
            if not re.search(a.by_shape, names[i]):
                continue
            b = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
            step_mb = 1 if b < 50e6 else 5
            key = (short(names[i]).replace("void ", ""), grids[i], round(b / 1e6 / step_mb) * step_mb)
            shapes[key][0] += 1
            shapes[key][1] += m.get("gpu__time_duration.sum", 0.0)
            shapes[key][2] += b
        print("| kernel | grid | DRAM MB / launch | launches | us / launch | DRAM GB/s | ms / step |\n|---|---|---|---|---|---|---|")
        for (k, grid, _), (n, t, b) in sorted(shapes.items(), key=lambda kv: -kv[1][1])[:30]:
            print(f"| `{k}` | {grid} | {b / n / 1e6:.0f} | {n} | {1e3 * t / n:.0f} | {b / (t * 1e-3) / 1e9:.0f} | {t:.2f} |")
    if a.json:
        fams = {k: {"launches": n, "time_ms": round(t, 3), "dram_bytes": int(b), "dram_bytes_per_launch": int(b / n)}
                for k, (n, t, b) in per_family.items() if k != "torch / other"}
        with open(a.json, "w") as f:
            json.dump({"source": a.csv, "families": fams}, f, indent=1)


if __name__ == "__main__":
    main()
