"""Summarise an ncu report: key raw metrics + the hottest SASS instructions with their stall reasons.
    python tools/ncu_hot.py gpurun_out/x.ncu-rep [n_top]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, d = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__issue_active.avg.pct',
        'sm__warps_active.avg.pct', 'launch__occupancy_limit', 'launch__registers_per_thread ', 'sm__cycles_elapsed.avg ', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic', 'lts__throughput.avg.pct', 'gpu__dram_throughput.avg.pct']
for a, b, c in zip(h, u, d):
    if any((a.startswith(k.strip()) if k.endswith(' ') else k in a) for k in keys):
        print(f"{a} [{b}] = {c}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ia, isrc, isamp, iex = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stall = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
data = [r for r in rows[2:] if len(r) > isamp and r[isamp].isdigit()]
tot = sum(int(r[isamp]) for r in data)
c = Counter()
for r in data:
    for i in stall:
        if r[i].isdigit():
            c[h[i]] += int(r[i])
print("total samples", tot, [(k, v) for k, v in c.most_common(8)])
for r in sorted(data, key=lambda r: -int(r[isamp]))[:ntop]:
    st = sorted(((h[i], int(r[i])) for i in stall if r[i].isdigit() and int(r[i]) > 0), key=lambda kv: -kv[1])[:2]
    print(r[ia][-5:], r[isamp].rjust(6), r[iex].rjust(9), r[isrc][:64].ljust(64), st)
