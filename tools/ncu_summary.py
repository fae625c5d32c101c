#!/usr/bin/env python3
"""Print the key metrics of an .ncu-rep (first profiled kernel): tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    print("kernel:", vals[hdr.index("Kernel Name")][:100])
    want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
            'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
            'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active']
    for k in want:
        if k in hdr:
            i = hdr.index(k); print("  %-72s %s %s" % (k, vals[i], units[i]))
    st = []
    for i, k in enumerate(hdr):
        if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k:
            try: v = float(vals[i])
            except ValueError: continue
            if v > 0: st.append((v, k.replace('smsp__pcsamp_warps_issue_stalled_', '')))
    tot = sum(v for v, _ in st) or 1
    print("  stall samples: " + ", ".join("%s %.0f (%.0f%%)" % (k, v, 100 * v / tot) for v, k in sorted(st, reverse=True)[:9]))
