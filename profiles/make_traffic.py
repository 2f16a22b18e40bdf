"""profiles/traffic.json from the committed ncu captures (what bench.py reports as `roofline.traffic`).

    python profiles/make_traffic.py profiles/r2_ncu_shapes_traffic.csv profiles/r2_ncu_kernels_summary.csv

* SHAPES csv: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size
  --clock-control none -k regex:env_kernel --csv python profiles/capture_kernels.py shapes` -- per shape one reset
  launch followed by two step launches (MODE 1); the step launches are averaged;
* SUMMARY csv: `profiles/ncu_summary.py` of the `--set full` capture at the headline shape (uf100-430 x 65,536).
Launches are cold-cache and serialised under ncu: for batches whose per-step working set fits the 126 MB L2 the
DRAM write count is far below the algorithmic bytes (the lines are still in L2 when the kernel ends)."""
import csv
import json
import re
import sys
from pathlib import Path

shapes_csv, summary_csv = sys.argv[1], sys.argv[2]
ORDER = [("uf100-430", 32768), ("uf100-430", 16384), ("uf100-430", 8192), ("uf50-218", 4096), ("uf250-1065", 16384),
         ("uf250-1065", 2048), ("mixed-k3-7", 32768), ("mixed-k3-7", 4096), ("uf20-91", 65536)]
rows = [r for r in csv.reader(l for l in open(shapes_csv) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows[1:]:
    launches.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]]})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]])
out = {}
steps, shape_i, seen_reset = [], -1, False
for _id in sorted(launches):
    L = launches[_id]
    mode = int(re.search(r"env_kernel<\d+, (\d)", L["kernel"]).group(1))
    if mode == 0:                                   # a reset launch opens the next shape
        shape_i += 1
        continue
    if mode == 1 and 0 <= shape_i < len(ORDER):
        out.setdefault(ORDER[shape_i], []).append(L)
res = {}
for (name, bs), ls in out.items():
    rd = sum(l["dram__bytes_read.sum"] for l in ls) / len(ls)
    wr = sum(l["dram__bytes_write.sum"] for l in ls) / len(ls)
    res[f"{name}:{bs}"] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                           "ncu_time_us": sum(l["gpu__time_duration.sum"] for l in ls) / len(ls) / 1e3,
                           "source": f"{shapes_csv} (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, mean of "
                                     f"{len(ls)} step launches, cold cache, serialised)"}
srows = list(csv.reader(open(summary_csv)))
ncol = len(srows[0])
# older summaries wrote the kernel name unquoted (it contains commas): re-join the surplus leading fields
srows = [srows[0]] + [[",".join(r[:len(r) - ncol + 1])] + r[len(r) - ncol + 1:] for r in srows[1:]]
h = {c.split("[")[0]: i for i, c in enumerate(srows[0])}
units = {c.split("[")[0]: c.split("[")[1].rstrip("]") for c in srows[0] if "[" in c}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
head = [r for r in srows[1:] if r[0].startswith("env_kernel<256, 1, 1, 0, 0>") and r[h["grid"]] == "65536"]
if head:
    rd = sum(float(r[h["dram_rd"]]) for r in head) / len(head) * scale[units["dram_rd"]]
    wr = sum(float(r[h["dram_wr"]]) for r in head) / len(head) * scale[units["dram_wr"]]
    res = {"uf100-430:65536": {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                               "ncu_time_us": sum(float(r[h["time"]]) for r in head) / len(head),
                               "source": f"{summary_csv} (ncu --set full, mean of {len(head)} launches of "
                                         f"env_kernel<256, MODE_STEP, OBS>)"}, **res}
res["_note"] = ("DRAM bytes per step launch measured by ncu (cold cache, serialised launches). For batches whose per-step "
                "working set fits the 126 MB L2 (uf50-218 x 4096, uf20-91, the 2,048 / 4,096-env shards) most observation "
                "writes are still in L2 when the kernel ends, so the DRAM write count is far below the algorithmic bytes.")
Path(__file__).with_name("traffic.json").write_text(json.dumps(res, indent=1))
print(json.dumps({k: v["dram_bytes_per_launch"] for k, v in res.items() if k != "_note"}, indent=1))
