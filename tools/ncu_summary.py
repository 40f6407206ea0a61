"""Print the handful of ncu metrics we judge kernels by from a .ncu-rep (run where ncu exists)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__shared_mem_per_block_dynamic",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("== kernel:", d.get("Kernel Name", "?")[:100])
    for k in KEYS:
        if k in d:
            print(f"  {k:75s} {d[k]:>16s} {units[hdr.index(k)]}")
    stalls = [(float(v.replace(',', '')), h) for h, v in d.items()
              if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and v]
    for v, h in sorted(stalls, reverse=True)[:8]:
        print(f"  stall {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio',''):40s} {v:8.2f}")
