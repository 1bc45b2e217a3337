#!/usr/bin/env python
"""Top stall-sample instructions of one kernel of an `ncu --set full --import-source on` report (SASS view).
  ncu_hot_lines.py <report.ncu-rep> <kernel regex> [n_lines] [launch-skip]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(k for k, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[h]
print(rows[0][:2])
i_s, i_src, i_ex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
data = []
for k, r in enumerate(rows[h + 1:]):
    try:
        data.append((int(r[i_s] or 0), int(r[i_ex] or 0), k, r[i_src].strip()))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
print(f"total samples {tot}, {len(data)} SASS lines, {sum(d[1] for d in data)} warp instructions")
for s, e, k, src in sorted(data, reverse=True)[:top]:
    print(f"{s:6d} {100 * s / tot:5.1f}%  executed {e:8d}  #{k:5d}  {src}")
