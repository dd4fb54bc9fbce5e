"""Writes a text summary of an .ncu-rep (key metrics per kernel) -- the files kept under profiles/.
Usage: ncu_summary.py rep.ncu-rep > profiles/xxx.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    print("=" * 100)
    print(r[idx['Kernel Name']], " grid", r[idx.get('Grid Size', 0)] if 'Grid Size' in idx else "")
    for k in keys:
        if k in idx:
            print("  %-72s %s %s" % (k, r[idx[k]], units[idx[k]]))
    st = [(h, float(r[i].replace(',', ''))) for h, i in idx.items()
          if 'smsp__average_warps_issue_stalled' in h and '_not_issued' not in h and r[i] not in ('', 'n/a')]
    print("  warp stall breakdown (warps per issue-active cycle):")
    for h, v in sorted(st, key=lambda x: -x[1])[:10]:
        print("     %-40s %.3f" % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
