import csv,sys,subprocess,json
rep=sys.argv[1]; out=sys.argv[2]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
want=["Kernel Name","launch__grid_size","launch__block_size","launch__registers_per_thread","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active","smsp__average_warp_latency_per_inst_issued.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio","smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio","l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct","l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct","sass__inst_executed_local_loads","sass__inst_executed_local_stores","lts__t_sector_hit_rate.pct","launch__occupancy_limit_registers","sm__maximum_warps_per_active_cycle_pct","launch__waves_per_multiprocessor"]
with open(out,"w") as f:
    f.write("# ncu --set full --clock-control none, summary of %s (ncu -i ... --page raw --csv, selected metrics)\n" % rep.split("/")[-1])
    for r in rows[2:]:
        f.write("\n")
        for w in want:
            if w in hdr:
                i=hdr.index(w); f.write("%s [%s] = %s\n" % (w, units[i], r[i]))
print(open(out).read()[:3000])
