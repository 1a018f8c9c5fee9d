#!/usr/bin/env python
"""tools/traffic_sum.py out.csv [...] -- per-kernel and per-step sums of an ncu --csv launch list
(dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum)."""
import csv
import json
import re
import sys
from collections import defaultdict


def parse(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    return rows


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12,
            "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "second": 1e3}
    return v * mult.get(u, 1.0)


def main():
    for path in sys.argv[1:]:
        per = defaultdict(lambda: defaultdict(float))
        cnt = defaultdict(set)
        for r in parse(path):
            name = re.sub(r"<.*", "", r.get("Kernel Name", "?")).replace("void ", "").replace("adv::", "")
            per[name][r["Metric Name"]] += to_bytes(r["Metric Value"], r.get("Metric Unit", ""))
            cnt[name].add(r.get("ID"))
        out = {}
        tot = defaultdict(float)
        for k, v in per.items():
            out[k] = {"launches": len(cnt[k]), "dram_read_GB": round(v.get("dram__bytes_read.sum", 0) / 1e9, 4),
                      "dram_write_GB": round(v.get("dram__bytes_write.sum", 0) / 1e9, 4), "time_ms": round(v.get("gpu__time_duration.sum", 0), 4)}
            for kk, vv in out[k].items():
                tot[kk] += vv
        out["_total"] = {k: round(v, 4) for k, v in tot.items()}
        out["_total"]["dram_GB"] = round(tot["dram_read_GB"] + tot["dram_write_GB"], 4)
        print(path, json.dumps(out))


if __name__ == "__main__":
    main()
