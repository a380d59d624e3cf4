import csv,sys,subprocess
f=sys.argv[1]
out=subprocess.run(["ncu","-i",f,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
keys=['Kernel Name','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct','launch__occupancy_limit','launch__waves_per_multiprocessor','dram__bytes_read.sum','dram__bytes_write.sum','smsp__average_warp','smsp__warp_issue_stalled','l1tex__data_bank_conflicts','smsp__inst_executed.sum','sm__cycles_elapsed.max','launch__grid_size','lts__t_bytes.sum','sm__ctas_launched','smsp__cycles_active.avg','smsp__inst_executed_op_local','local']
for r in rows[2:]:
    print('=========')
    for i,h in enumerate(hdr):
        if any(h.startswith(k) for k in keys):
            try:
                v=float(r[i].replace(',',''))
                if 'stalled' in h and v<0.3: continue
            except: pass
            print(f'  {h} [{rows[1][i]}] = {r[i]}')
