"""Turns the ncu reports that came back in gpurun_out/ into small committed summaries under profiles/.

    python tools/summarize_profiles.py r01 k3_full k2_full      # full-set captures -> profiles/r01_<name>.md
    python tools/summarize_profiles.py r01 --launches launches  # launch list -> profiles/r01_launches_summary.md
"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes.sum.per_second", "launch__registers_per_thread ", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct", "sm__inst_executed.sum ", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
        "launch__block_size", "smsp__average_warp", "smsp__warp_issue_stalled", "sm__cycles_elapsed.avg ", "launch__shared_mem_per_block",
        "smsp__inst_executed.sum ", "lts__t_sectors_op_atom", "lts__t_sectors_op_red", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ",
        "sm__pipe_fmaheavy_cycles_active.avg.pct", "sm__pipe_fmalite_cycles_active.avg.pct", "smsp__issue_active.avg.pct", "smsp__average_warps_issue_stalled")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def full(tag, names):
    for name in names:
        rep = os.path.join(ROOT, "gpurun_out", name + ".ncu-rep")
        hdr, units, data = raw(rep)
        lines = [f"# ncu --set full summary: {name} ({tag})", ""]
        for row in data:
            kname = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            lines += [f"## {kname}", "| metric | unit | value |", "|---|---|---|"]
            for h, u, v in zip(hdr, units, row):
                if any((h.startswith(k.strip()) if k.endswith(" ") else k in h) for k in KEEP) and v != "":
                    lines.append(f"| {h} | {u} | {v} |")
            lines.append("")
        with open(os.path.join(ROOT, "profiles", f"{tag}_{name}.md"), "w") as f:
            f.write("\n".join(lines))
        print("wrote", f"profiles/{tag}_{name}.md")


def launches(tag, name):
    rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", name + ".csv"))))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = [f"# ncu launch list summary ({tag}): gpu__time_duration.sum per kernel, cold-cache & serialised — compare SHARES",
             "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"| `{k[:110]}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.2f}% |")
    with open(os.path.join(ROOT, "profiles", f"{tag}_{name}_summary.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if "--launches" in sys.argv:
        launches(sys.argv[1], sys.argv[sys.argv.index("--launches") + 1])
    else:
        full(sys.argv[1], sys.argv[2:])
