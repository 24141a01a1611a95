import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('==',d['Kernel Name'][:80])
    keys=['gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active',
     'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
     'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum','lts__t_sector_hit_rate.pct','lts__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed']
    for k in keys: print(' ',k,d.get(k))
    for h,v in d.items():
        if 'average_warps_issue_stalled' in h and 'per_issue_active.ratio' in h and float(v or 0)>0.3: print('  stall',h.split('stalled_')[1].split('_per')[0],v)
