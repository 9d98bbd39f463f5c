"""Summarise an .ncu-rep (ncu --set full) into a small CSV for profiles/: one row per captured launch, a fixed metric set.
usage: ncu_summary.py <in.ncu-rep> <out.csv>"""
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "launch__grid_size", "launch__block_size",
           "launch__registers_per_thread", "launch__cluster_size" , "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum",
           "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
           "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__lsu_writeback_active_mem_lgds.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum",
           "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
cols = [hdr.index("Kernel Name")] + [hdr.index(m) for m in METRICS if m in hdr]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[c] for c in cols])
    w.writerow([units[c] for c in cols])
    for r in rows[2:]:
        w.writerow([r[c] for c in cols])
print("wrote", sys.argv[2], len(rows) - 2, "launches")
