"""Summarise an ncu report per barrier-delimited phase: python tools/ncu_phases.py rep.ncu-rep [launch_index]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:150])
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}


def num(r, k):
    try:
        return float(r[idx[k]])
    except (ValueError, KeyError, IndexError):
        return 0.0


seen, data = set(), []
for r in rows[2:]:
    if r[idx["Address"]] in seen:
        continue
    seen.add(r[idx["Address"]])
    data.append(r)
tot = sum(num(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
keys = ("stall_long_sb", "stall_barrier", "stall_short_sb", "stall_wait", "stall_mio", "stall_math", "stall_selected",
        "stall_not_selected", "stall_no_inst", "stall_branch_resolving", "stall_lg", "stall_dispatch")
seg, cur, acc, n, ex = [], 0, collections.Counter(), 0, 0
for r in data:
    cur += num(r, "# Samples")
    n += 1
    ex += num(r, "Instructions Executed")
    for k in keys:
        acc[k] += num(r, k)
    if "BAR.SYNC" in r[idx["Source"]]:
        seg.append((cur, n, ex, dict(acc)))
        cur, acc, n, ex = 0, collections.Counter(), 0, 0
seg.append((cur, n, ex, dict(acc)))
for i, (c, n, ex, a) in enumerate(seg):
    top = sorted(a.items(), key=lambda kv: -kv[1])[:4]
    print("phase %2d: samples %6d (%4.1f%%) sass %5d warp-instr %9d  %s" % (i, c, 100 * c / max(tot, 1), n, ex, " ".join("%s=%d" % (k[6:], v) for k, v in top)))
top = sorted(data, key=lambda r: -num(r, "# Samples"))[:12]
for r in top:
    print("%6d  %s" % (num(r, "# Samples"), r[idx["Source"]][:100]))
