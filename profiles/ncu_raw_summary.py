#!/usr/bin/env python
"""Key metrics per kernel from `ncu -i X.ncu-rep --page raw --csv` (argv[1]); optional argv[2] = JSON output."""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
H = rows[0]; U = rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active']
out = {}
ki = H.index('Kernel Name')
for r in rows[2:]:
    name = r[ki].split('(')[0]
    if name.startswith('void '):
        name = name[5:]
    d = {}
    for w in want:
        if w in H:
            i = H.index(w); d[w] = {"value": r[i], "unit": U[i]}
    st = [(float(r[i].replace(',', '')), H[i]) for i in range(len(H))
          if H[i].startswith('smsp__average_warps_issue_stalled_') and H[i].endswith('_per_issue_active.ratio') and r[i]]
    d["top_stalls_per_issue"] = [[round(v, 2), n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')]
                                 for v, n in sorted(st, reverse=True)[:7]]
    out[name] = d
    print('----', name)
    for k, v in d.items():
        print('  ', k, v if k == "top_stalls_per_issue" else (v["value"], v["unit"]))
if len(sys.argv) > 2:
    src = sys.argv[3] if len(sys.argv) > 3 else "ncu --set full --clock-control none, bench.py default workload (config 2, 100 M entries, 1 GPU), one launch per kernel"
    json.dump({"source": src, "kernels": out}, open(sys.argv[2], 'w'), indent=1)
