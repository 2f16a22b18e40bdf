"""Per-source-line instruction counts / stall samples of one kernel from an ncu report captured with
`--set full --import-source on` (library built with -lineinfo).

    python profiles/hotlines.py REPORT.ncu-rep KERNEL_REGEX MANGLED_SUBSTRING [launch_skip] > out.csv

ncu's CLI exports the SASS view only; the SASS instructions are joined, in order, with the `//## File ..., line N`
annotations that `nvdisasm -g` prints for the same function of the shipped cubin."""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

rep, kregex, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
lib = Path(__file__).resolve().parent.parent / "marl_sat_b200" / "csrc" / "libmarlsat_b200.so"
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kregex}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci, si, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
sass = []
for r in rows[h + 1:]:
    if len(r) <= ci or not r[0].startswith("0x"):
        break                                   # end of the first kernel's block
    sass.append((r[si].strip(), int(r[ci] or 0), int(r[smp] or 0)))

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(lib)], cwd=td, capture_output=True)
    lines = []
    for cub in Path(td).glob("*.cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", str(cub)], capture_output=True, text=True).stdout
        if mangled in out:
            lines = out.splitlines()
            break
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and mangled in l and l.rstrip().endswith(":"))
cur, per_insn = ("?", 0), []
for l in lines[start + 1:]:
    if l.startswith("//---") or l.startswith("\t.section"):
        if per_insn:
            break
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (Path(m.group(1)).name, int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        per_insn.append(cur)
if len(per_insn) != len(sass):
    print(f"# warning: {len(per_insn)} disassembled instructions vs {len(sass)} profiled", file=sys.stderr)
agg = defaultdict(lambda: [0, 0])
for (f, ln), (_, n, s) in zip(per_insn, sass):
    agg[(f, ln)][0] += n
    agg[(f, ln)][1] += s
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[1] for v in agg.values()) or 1
text = {}
w = csv.writer(sys.stdout)
print(f"# {kregex} / {mangled}: warp instructions executed and stall samples per source line")
w.writerow(["total_warp_instructions", tot_i, "total_samples", tot_s])
w.writerow(["warp_inst", "pct_inst", "samples", "pct_samples", "file:line", "source"])
for (f, ln), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 45]:
    if f not in text:
        cand = list((lib.parent).glob(f)) + list((lib.parent.parent.parent / "include").glob(f))
        text[f] = cand[0].read_text().splitlines() if cand else []
    srcl = text[f][ln - 1].strip() if 0 < ln <= len(text[f]) else ""
    w.writerow([n, f"{100 * n / tot_i:.1f}", s, f"{100 * s / tot_s:.1f}", f"{f}:{ln}", srcl])
