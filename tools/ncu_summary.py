#!/usr/bin/env python
"""Summarise ncu output into the small text tables committed under profiles/.

  ncu_summary.py launches <launches.csv>            per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` log
  ncu_summary.py full <prof.ncu-rep> [--json out]   one line per captured launch of an `ncu --set full` report
                                                    (duration, DRAM bytes read / written, DRAM %, SM %, occupancy, regs)
Runs here (no GPU needed): `ncu -i` only reads the report.
"""
from __future__ import annotations

import collections
import csv
import io
import json
import re
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("heros::", "").replace("(anonymous namespace)::", "").replace("heros_impl::", "")
    return name if len(name) < 70 else name[:67] + "..."


def launches(path: str) -> None:
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    tot = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        k = short(r["Kernel Name"])
        n, s = tot.get(k, (0, 0.0))
        tot[k] = (n + 1, s + v)
    total = sum(s for _, s in tot.values())
    print(f"# {path}: {sum(n for n, _ in tot.values())} launches, {total / 1e3:.3f} ms of kernel time (cold-cache, serialised)")
    print(f"{'kernel':70s} {'launches':>8s} {'total us':>11s} {'avg us':>9s} {'share':>7s}")
    for k, (n, s) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:70s} {n:8d} {s:11.1f} {s / n:9.2f} {100 * s / total:6.1f}%")


COLS = [("gpu__time_duration.sum", "us", "dur_us"), ("dram__bytes_read.sum", "byte", "dram_rd_B"),
        ("dram__bytes_write.sum", "byte", "dram_wr_B"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", None, "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", None, "sm_pct"),
        ("lts__t_sector_hit_rate.pct", None, "l2_hit_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", None, "occ_pct"),
        ("launch__registers_per_thread", None, "regs"), ("launch__grid_size", None, "grid"),
        ("launch__block_size", None, "block")]


def full(path: str, json_out: str | None) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    idx = [(hdr.index(m) if m in hdr else -1, base, label) for m, base, label in COLS]
    recs = []
    for r in data:
        rec = {"kernel": short(r[ki])}
        for i, base, label in idx:
            if i < 0 or r[i] == "":
                rec[label] = None
                continue
            v = float(r[i].replace(",", ""))
            if base == "byte":
                v *= UNIT.get(units[i], 1.0)
            elif base == "us":
                v *= UNIT.get(units[i], 1.0)
            rec[label] = v
        rec["traffic_B"] = (rec["dram_rd_B"] or 0.0) + (rec["dram_wr_B"] or 0.0)
        recs.append(rec)
    print(f"# {path}: {len(recs)} captured launches (ncu --set full, --clock-control none)")
    print(f"{'kernel':58s} {'us':>8s} {'dramRd MB':>9s} {'dramWr MB':>9s} {'dram%':>6s} {'sm%':>6s} {'L2hit%':>6s} {'occ%':>6s} {'regs':>5s} {'grid':>6s} {'blk':>5s}")
    for r in recs:
        def f(x, fmt):
            return format(x, fmt) if x is not None else "-"
        print(f"{r['kernel'][:58]:58s} {f(r['dur_us'], '8.2f')} {f((r['dram_rd_B'] or 0) / 1e6, '9.3f')} "
              f"{f((r['dram_wr_B'] or 0) / 1e6, '9.3f')} {f(r['dram_pct'], '6.1f')} {f(r['sm_pct'], '6.1f')} "
              f"{f(r['l2_hit_pct'], '6.1f')} {f(r['occ_pct'], '6.1f')} {f(r['regs'], '5.0f')} {f(r['grid'], '6.0f')} "
              f"{f(r['block'], '5.0f')}")
    if json_out:
        with open(json_out, "w") as fo:
            json.dump(recs, fo, indent=1)


if __name__ == "__main__":
    if len(sys.argv) < 3:
        raise SystemExit(__doc__)
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[4] if len(sys.argv) > 4 and sys.argv[3] == "--json" else None)
