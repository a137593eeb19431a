"""Count the Blackwell-specific SASS mnemonics per kernel of the built library (cuobjdump -sass): UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UBLKCP / UTMALDG (bulk / tensor TMA copies), LDGSTS (cp.async), SYNCS (mbarrier), HMMA (legacy
mma.sync: expected 0).  Usage: python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "eco-dqn_b200", "lib", "libecodqn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = [("UTC*MMA", r"\bUTC\w*MMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UBLKCP", r"\bUBLKCP"),
        ("UTMALDG", r"\bUTMALDG"), ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("DADD/DMUL/DFMA", r"\bD(ADD|MUL|FMA)\b")]
counts, size, name = collections.OrderedDict(), {}, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name).replace("eco::", "")
        counts[name] = collections.Counter()
        size[name] = 0
        continue
    if name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        size[name] += 1
        for key, pat in pats:
            if re.search(pat, line):
                counts[name][key] += 1
print("# %s  (sm_100a SASS; instruction counts per kernel)" % os.path.relpath(lib, ROOT))
print("%-46s %6s " % ("kernel", "instr") + " ".join("%8s" % k for k, _ in pats))
for k, c in counts.items():
    print("%-46s %6d " % (k[:46], size[k]) + " ".join("%8d" % c[key] for key, _ in pats))
