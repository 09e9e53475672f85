"""Print the key metrics of an .ncu-rep (first kernel): python scripts/ncu_keys.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, u, v = r[0], r[1], r[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warps_issue_stalled", "smsp__warps_issue_stalled", "l1tex__throughput.avg.pct"]
for i, k in enumerate(h):
    if any(k.startswith(x) for x in keys):
        print("%-95s %-14s %s" % (k, u[i], v[i]))
