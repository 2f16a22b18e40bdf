"""Key per-launch metrics of every kernel in an ncu report:  python profiles/ncu_summary.py REPORT.ncu-rep"""
import csv
import subprocess
import sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("smsp__inst_executed.sum", "warp_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dsmem"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%")]
w = csv.writer(sys.stdout)          # kernel names contain commas: quoted
w.writerow(["kernel"] + [f"{n}[{units[idx[m]]}]" for m, n in cols if m in idx])
for r in data:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    w.writerow([name] + [r[idx[m]] for m, n in cols if m in idx])
