"""Warp instructions and stall samples of one kernel of an .ncu-rep, summed per out-of-line routine.

The static kernels keep their building blocks out of line (`__noinline__`: the kernels run on instruction fetch), so the
kernel's SASS is a main body followed by the routines it calls.  ncu's source page lists every SASS instruction with its
execution count and stall samples but no routine names; this script takes the targets of the CALL instructions as routine
entry points and sums the counters between consecutive entry points.  Call counts identify the routines (the matrix-vector
products are called ~40 times per configuration, the node routine ~17 times, the Jacobian once, ...).

usage: ncu_by_routine.py X.ncu-rep <kernel name regex> <units per launch> > profiles/....txt"""
import bisect
import csv
import io
import re
import subprocess
import sys

rep, kernel, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) > 8 and r[0].startswith("0x")]
col = {n: hdr.index(n) for n in ("Address", "Source", "Instructions Executed", "# Samples", "Thread Instructions Executed")}
base = int(data[0][col["Address"]], 16)
ins = [(int(r[col["Address"]], 16) - base, r[col["Source"]].strip(), int(r[col["Instructions Executed"]]), int(r[col["# Samples"]]),
        int(r[col["Thread Instructions Executed"]])) for r in data]
calls = {}
for off, s, ex, smp, thr in ins:
    m = re.search(r"CALL\.\w+(?:\.\w+)*\s+(0x[0-9a-f]+)", s)
    if m:
        t = int(m.group(1), 16) - base
        calls[t] = calls.get(t, 0) + ex
starts = sorted(set(calls) | {0})
agg = {s: [0, 0, 0, 0] for s in starts}
for off, s, ex, smp, thr in ins:
    a = agg[starts[bisect.bisect_right(starts, off) - 1]]
    a[0] += 1
    a[1] += ex
    a[2] += smp
    a[3] += thr
tot_ex, tot_smp = sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values())
print(f"# {rows[0][1] if rows and len(rows[0]) > 1 else kernel}")
print(f"# {tot_ex / units:.0f} warp instructions per unit ({units:.0f} units per launch), {len(ins)} SASS instructions")
print(f"{'entry':>8} {'SASS':>6} {'inst/unit':>10} {'inst %':>7} {'samples %':>9} {'calls/unit':>10} {'lanes':>6}")
for s in starts:
    a = agg[s]
    print(f"{hex(s):>8} {a[0]:6d} {a[1] / units:10.1f} {100 * a[1] / tot_ex:7.1f} {100 * a[2] / tot_smp:9.1f} {calls.get(s, 0) / units:10.2f} "
          f"{a[3] / max(a[1], 1):6.1f}")
