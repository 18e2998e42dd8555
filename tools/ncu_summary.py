"""Summarise an .ncu-rep (read on the CPU box): one block of key metrics per distinct kernel."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_issued.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__issue_active.avg.per_cycle_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_lg.sum', 'sm__cycles_active.avg']
import re
want += [h for h in hdr if re.match(r'(sm__inst_executed_pipe_\w+\.avg\.pct_of_peak_sustained_active|sm__pipe_\w+_cycles_active\.avg\.pct_of_peak_sustained_active|'
                                    r'l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|'
                                    r'l1tex__data_pipe_lsu_wavefronts\.sum|l1tex__lsu_writeback_active\.avg\.pct_of_peak_sustained_elapsed|'
                                    r'l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed|l1tex__t_.*pct_of_peak_sustained_elapsed|'
                                    r'smsp__inst_executed_op_shared_ld\.sum|l1tex__f_.*pct.*|l1tex__m_.*pct.*)$', h)]
ki = hdr.index('Kernel Name')
seen = set()
for r in rows[2:]:
    name = r[ki].split('(')[0].replace('pbf::<unnamed>::', '')
    if name in seen:
        continue
    seen.add(name)
    print("=====", name)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:82s} {r[i]:>16s} {units[i]}")
