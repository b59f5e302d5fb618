#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python scripts/summarize_launches.py gpurun_out/launches.csv profiles/rN_launches.json [skip_first_n]

Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's live
numbers, not absolutes (B200_PROFILING.md)."""
import csv
import json
import re
import sys


def short(name):
    name = re.sub(r"^void\s+", "", name)
    m = re.match(r"(?:drba::)?(\w+)", name)
    if name.startswith("at::") or name.startswith("at_cuda") or "at::native" in name:
        m2 = re.search(r"(\w+Functor|\w+_kernel\w*|copy\w*)", name)
        return "torch:" + (m2.group(1) if m2 else name[:40])
    return m.group(1) if m else name[:60]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = []
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        rows.append((int(r["ID"]), short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], v))
    rows = [r for r in rows if r[0] >= skip]
    total = sum(r[4] for r in rows) or 1.0
    agg = {}
    for _, name, grid, block, ns in rows:
        d = agg.setdefault(name, {"launches": 0, "ns": 0.0, "max_ns": 0.0})
        d["launches"] += 1
        d["ns"] += ns
        d["max_ns"] = max(d["max_ns"], ns)
    out = {"source": src, "launches": len(rows), "total_us": round(total / 1e3, 1),
           "note": "ncu per-launch times are cold-cache and serialised; shares are comparable, absolutes are not",
           "kernels": {k: {"launches": v["launches"], "total_us": round(v["ns"] / 1e3, 1),
                           "avg_us": round(v["ns"] / 1e3 / v["launches"], 2), "max_us": round(v["max_ns"] / 1e3, 1),
                           "share": round(v["ns"] / total, 4)}
                       for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ns"])}}
    json.dump(out, open(dst, "w"), indent=1)
    for k, v in out["kernels"].items():
        print(f"{v['share']:7.3f} {v['launches']:5d} {v['avg_us']:9.2f} us  {k}")


if __name__ == "__main__":
    main()
