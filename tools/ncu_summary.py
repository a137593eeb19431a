"""Summarise an .ncu-rep (raw metrics + stall samples by SASS window) into a small text file for profiles/."""
import csv
import re
import subprocess
import sys
from collections import Counter

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, unit, val = rows[0], rows[1], rows[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
lines = ["# %s" % rep, "# kernel: %s" % val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""]
for h, u, v in zip(hdr, unit, val):
    if h in want or h.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in h:
        lines.append("%s = %s %s" % (h, v, u))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h2, data = rows[1], rows[2:]
iS, isrc, iex = h2.index("# Samples"), h2.index("Source"), h2.index("Instructions Executed")
tot = sum(int(r[iS]) for r in data) or 1
lines.append("# stall samples by SASS window of 100 instructions (total %d samples, %d instructions)" % (tot, len(data)))
for w in range(0, len(data), 100):
    seg = data[w:w + 100]
    s = sum(int(r[iS]) for r in seg)
    ops = [(r[isrc].split()[1] if r[isrc].strip().startswith("@") else r[isrc].split()[0]) for r in seg if r[isrc].split()]
    c = Counter(o for o in ops if re.match(r"(UTC|LDTM|STTM|BAR|SYNCS|LDG|STS|LDS|STG|MUFU|CALL)", o))
    lines.append("%5d %5.1f%% exec<=%s %s" % (w, 100.0 * s / tot, max(int(r[iex]) for r in seg), dict(c)))
top = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:25]
lines.append("# top instructions by samples")
for i in sorted(top):
    lines.append("%5d %6s %9s %s" % (i, data[i][iS], data[i][iex], data[i][isrc].strip()[:100]))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
