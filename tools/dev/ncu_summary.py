"""Summarise an .ncu-rep (first kernel): key raw metrics + top stall lines. usage: ncu_summary.py rep [n_units]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run("ncu -i %s --page raw --csv" % rep, shell=True, capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h, u, v = r[0], r[1], r[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'launch__waves_per_multiprocessor', 'lts__t_sector_hit_rate.pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_bytes.sum', 'sm__inst_executed_pipe_tensor.sum']
for i, k in enumerate(h):
    if k in keys or ('smsp__average_warps_issue_stalled' in k and k.endswith('.ratio') and float(v[i]) > 0.25):
        print(k.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', ''), v[i], u[i])
out = subprocess.run("ncu -i %s --page source --csv --print-source sass" % rep, shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; data = rows[2:]
ia = h.index("Instructions Executed"); isamp = h.index("# Samples"); isrc = h.index("Source")
ts = sum(int(x[isamp]) for x in data)
tot = sum(int(x[ia]) for x in data)
print("samples", ts, "warp-instr", tot)
idx = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in sorted(idx):
    x = data[i]
    print("%5d %5.2f%% n=%8s  %s" % (i, 100 * int(x[isamp]) / ts, x[ia], x[isrc][:100]))
