"""SASS opcode histogram per kernel of csn_b200/libcsn_b200.so (cuobjdump -sass), written to profiles/: the evidence that
the tensor-core kernels are tcgen05 (UTCHMMA) fed by TMA (UTMALDG / UTMASTG / UTMAREDG) with accumulators in TMEM
(LDTM / STTM), and contain no legacy HMMA / HGMMA.  Usage: python scripts/sass_histogram.py [out.csv]"""
import collections
import csv
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
out = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "profiles" / "r2_sass_histogram.csv")
sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "csn_b200" / "libcsn_b200.so")], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
KEY = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "HGMMA",
       "FFMA", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "ATOM", "RED", "DFMA"]
rows = []
cur, counts = None, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, counts))
        cur, counts = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        counts[base] += 1
        counts["_total"] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            counts["UTCHMMA.2CTA"] += 1
if cur:
    rows.append((cur, counts))
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "instructions"] + KEY)
    for name, c in sorted(rows, key=lambda r: -r[1]["_total"]):
        d = demangle(name)
        d = d.replace("(int)", "").replace("(bool)", "").replace("void ", "")
        d = re.sub(r"\(.*", "", d)
        w.writerow([d, c["_total"]] + [c.get(k, 0) for k in KEY])
tot = collections.Counter()
for _, c in rows:
    tot.update(c)
print(f"{len(rows)} kernels; UTCHMMA {tot['UTCHMMA']} (.2CTA {tot['UTCHMMA.2CTA']}), LDTM {tot['LDTM']}, STTM {tot['STTM']}, "
      f"UTMALDG {tot['UTMALDG']}, UTMASTG {tot['UTMASTG']}, UTMAREDG {tot['UTMAREDG']}, UBLKCP {tot['UBLKCP']}, HMMA {tot['HMMA']}, HGMMA {tot['HGMMA']}")
