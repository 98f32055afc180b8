"""Summarise an .ncu-rep (read here, no GPU): key metrics and warp-stall breakdown per kernel."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
key = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
       'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
       'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
       'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
       'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum',
       'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_op_shared_atom.sum']
stall = [h for h in H if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
if not stall:
    stall = [h for h in H if 'warp_issue_stalled' in h and h.endswith('.pct')]
ki = H.index('Kernel Name')
for r in rows[2:]:
    print('===', r[ki][:90])
    for k in key:
        if k in H:
            print(f"   {k:66s} {r[H.index(k)]:>16s} {U[H.index(k)]}")
    st = []
    for h in stall:
        try:
            st.append((float(r[H.index(h)].replace(',', '')), h))
        except ValueError:
            pass
    for v, h in sorted(st, reverse=True)[:7]:
        print(f"   stall {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','').replace('smsp__average_','')[:40]:42s} {v:8.2f}")
