"""Print a compact set of raw ncu metrics of a report (first profiled launch): python tools/ncu_keys.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active_realtime.avg.pct",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "launch__registers_per_thread", "sass__inst_executed_local_stores",
        "smsp__average_warps_issue_stalled", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__inst_executed_pipe_uniform", "smsp__inst_executed_pipe_lsu", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "smsp__warps_issue_stalled"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for h, u, v in zip(rows[0], rows[1], rows[2 + idx]):
    if any(k in h for k in KEYS) and "per_second" not in h and "peak_sustained" not in h.replace("pct_of_peak_sustained", ""):
        print(f"{h:95s} {u:8s} {v}")
