"""Attribute ncu SASS-level samples to CUDA source lines: python tools/ncu_lines.py rep.ncu-rep obj.o kernel_substring [launch]
(joins `ncu --page source` rows with `nvdisasm -g` line info by instruction order)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, sub = sys.argv[1], sys.argv[2], sys.argv[3]
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
lines, fn, cur = [], None, None
for l in dis.splitlines():
    if l.lstrip().startswith(".section") and ".text." in l:
        fn = l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/", l) and fn and sub in fn:
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data, seen = [], set()
for r in rows[2:]:
    if len(r) < len(hdr) or r[idx["Address"]] in seen or r[idx["Address"]] == "Address":
        continue
    seen.add(r[idx["Address"]])
    data.append(r)
print(rows[0][1][:120], "| sass rows", len(data), "disasm instrs", len(lines))
n = min(len(data), len(lines))
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for r, ln in zip(data[:n], lines[:n]):
    a = agg[ln]
    a[0] += float(r[idx["# Samples"]] or 0)
    a[1] += float(r[idx["Instructions Executed"]] or 0)
    for k in keys:
        try:
            a[2][k[6:]] += float(r[idx[k]] or 0)
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values())
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print("%6d (%4.1f%%) exec %9d  %s:%s  %s" % (a[0], 100 * a[0] / max(tot, 1), a[1], ln[0] if ln else "?", ln[1] if ln else "?",
                                              " ".join("%s=%d" % kv for kv in a[2].most_common(3))))
