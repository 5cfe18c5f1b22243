"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_global_red.sum']
for r in rows[2:]:
    print('----')
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f'{w:66s} {r[i]:>22s} {units[i]}')
    st = [(hdr[i], float(r[i].replace(',', '') or 0)) for i in range(len(hdr))
          if 'smsp__average_warps_issue_stalled' in hdr[i] and 'per_issue_active' in hdr[i]]
    st.sort(key=lambda x: -x[1])
    print('stall reasons (warps per issue-active cycle):')
    for n, v in st[:8]:
        print(f'  {n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""):28s} {v:.2f}')
