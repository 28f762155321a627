"""Selected lines of `ncu --page raw` for one report:  python tools/ncu_summary.py report.ncu-rep"""
import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit','sm__throughput.avg.pct','l1tex__throughput.avg.pct','lts__throughput.avg.pct','smsp__inst_executed.sum','smsp__issue_active.avg.pct','l1tex__data_pipe_lsu_wavefronts.avg','l1tex__data_pipe_lsu_wavefronts_mem','smsp__average_warps_issue_stalled','launch__grid_size','launch__block_size','l1tex__t_sector_hit_rate','lts__t_sector_hit_rate','smsp__thread_inst_executed_per_inst_executed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__cycles_active.avg','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','smsp__inst_executed_op_shared','smsp__inst_executed_op_global','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum','local']
for h,u,v in zip(hdr,units,vals):
    if any(w in h for w in want) and 'pcsamp' not in h:
        print(f"{h:88s} {v:>20s} {u}")
