import csv, subprocess, sys
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
want=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__grid_size","launch__block_size","launch__waves_per_multiprocessor","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","smsp__inst_executed.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__throughput.avg.pct_of_peak_sustained_active","lts__throughput.avg.pct_of_peak_sustained_elapsed","smsp__warps_eligible.avg.per_cycle_active","smsp__warps_active.avg.per_cycle_active","smsp__cycles_active.avg","sm__cycles_elapsed.max"]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:70], "grid", r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
    for w in want:
        if w in hdr: print("  %-70s %s %s"%(w, r[hdr.index(w)], units[hdr.index(w)]))
    for i,h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in h:
            try:
                if float(r[i])>0: print("  ",h.replace("smsp__pcsamp_warps_issue_stalled_","stall:"), r[i])
            except: pass
