"""Per-device-function share of executed instructions and stall samples of one kernel in an ncu report.

usage: ncu_by_function.py <sass_source_page.csv> <kernel index in the csv> <cubin> <kernel symbol substring>
The csv is `ncu -i X.ncu-rep --page source --csv --print-source sass`; the cubin comes from
`cuobjdump -xelf all libctrb200.so`.  Function boundaries are nvdisasm's labels."""
import csv
import re
import subprocess
import sys

src, kidx, cubin, ksym = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
dis = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
# locate the kernel's .text section, collect (offset, label) for every non-local label
labels, cur, inside = [], None, False
for line in dis:
    if line.startswith(".text."):
        inside = ksym in line
        continue
    if not inside:
        continue
    m = re.match(r"^(\$?[A-Za-z_][^:\s]*):\s*$", line)
    if m and not m.group(1).startswith(".L_"):
        cur = m.group(1)
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/", line)
    if m and cur is not None:
        labels.append((int(m.group(1), 16), cur))
        cur = None
labels.sort()

def short(lbl):
    m = re.findall(r"\d+([a-z_][a-z_0-9]*)E", lbl)
    return m[-1] if m else lbl[-40:]

rows = list(csv.reader(open(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
lo = starts[kidx]
hi = starts[kidx + 1] if kidx + 1 < len(starts) else len(rows)
hdr = rows[lo + 1]
ia, ii, ism = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
body = [r for r in rows[lo + 2:hi] if len(r) > ii]
base = int(body[0][ia], 16)
agg = {}
for r in body:
    off = int(r[ia], 16) - base
    name = "kernel body"
    for o, l in labels:
        if o <= off:
            name = short(l)
        else:
            break
    a = agg.setdefault(name, [0, 0, 0])
    a[0] += int(r[ii]); a[1] += int(r[ism]); a[2] += 1
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"{'function':28s} {'SASS':>6s} {'inst %':>7s} {'samples %':>9s}   total inst {ti:.3e}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} {a[2]:6d} {100 * a[0] / ti:7.2f} {100 * a[1] / max(ts, 1):9.2f}")
