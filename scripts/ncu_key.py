"""Prints the metrics that decide what bounds a kernel from an .ncu-rep (ncu --set full): usage ncu_key.py rep [row]"""
import csv
import re
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = [r"^Kernel Name$", r"^Grid Size$", r"gpu__time_duration.sum$", r"sm__cycles_elapsed.avg$", r"sm__cycles_active.avg$",
        r"smsp__inst_executed.sum$", r"thread_inst_executed_per_inst_executed.ratio$", r"launch__registers_per_thread$",
        r"sm__warps_active.avg.pct_of_peak_sustained_active$", r"sm__inst_issued.avg.pct_of_peak_sustained_active$",
        r"sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active$", r"sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active$",
        r"sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active$", r"sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active$",
        r"^l1tex__throughput.avg.pct_of_peak_sustained_elapsed$", r"^lts__throughput.avg.pct_of_peak_sustained_elapsed$",
        r"^l1tex__t_sector_hit_rate.pct$", r"^lts__t_sector_hit_rate.pct$", r"^lts__t_sectors_op_read.sum$", r"^lts__t_sectors_op_write.sum$",
        r"^lts__t_bytes.sum$", r"^l1tex__t_bytes.sum$", r"^l1tex__m_xbar2l1tex_read_bytes.sum$", r"^l1tex__m_l1tex2xbar_write_bytes.sum$",
        r"^dram__bytes_read.sum$", r"^dram__bytes_write.sum$", r"^gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed$",
        r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$", r"l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed$",
        r"l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed$", r"l1tex__f_wavefronts.avg.pct_of_peak_sustained_elapsed$",
        r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio$", r"smsp__warps_eligible.avg.per_cycle_active$",
        r"smsp__issue_active.avg.per_cycle_active$", r"sm__throughput.avg.pct_of_peak_sustained_elapsed$",
        r"^lts__t_sectors_srcunit_tex_op_read.sum$", r"^lts__t_sectors_srcunit_tex_op_write.sum$",
        r"^lts__d_sectors_fill_sysmem.sum$", r"^l1tex__m_xbar2l1tex_read_sectors.sum$", r"^l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum$",
        r"^l1tex__t_sectors_pipe_lsu_mem_global_op_ldgsts.sum$", r"^l1tex__t_sectors_pipe_lsu_mem_global_op_ldgsts_lookup_hit.sum$",
        r"^l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum$", r"^l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum$",
        r"^launch__occupancy_limit_.*", r"^launch__waves_per_multiprocessor$", r"^sm__maximum_warps_per_active_cycle_pct$",
        r"achieved_occupancy", r"^launch__shared_mem_per_block", r"^smsp__cycles_active.avg$"]
sel = [int(a) for a in sys.argv[2:]] or list(range(len(rows) - 2))
for k in sel:
    r = rows[2 + k]
    print(f"--- launch {k}")
    for h, u, v in zip(hdr, units, r):
        if any(re.search(w, h) for w in want):
            if h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) < 0.05:
                continue
            print(f"{h:90s} {v} {u}")
