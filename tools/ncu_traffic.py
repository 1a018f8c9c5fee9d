#!/usr/bin/env python
"""tools/ncu_traffic.py <report.ncu-rep> [out.json] -- DRAM bytes per launch of the four step kernels
from an `ncu --set full` capture of ONE step (read here, no GPU needed).  bench.py copies the result
(profiles/traffic_latest.json) into roofline.traffic."""
import csv
import io
import json
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]

    def val(r, name):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12,
                 "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}.get(u, 1.0)
        return v * scale

    res = {"report": rep, "kernels": {}, "dram_bytes_per_step": 0.0}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        short = name.split("<")[0].replace("void ", "").replace("adv::", "")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        ms = val(r, "gpu__time_duration.sum")
        res["kernels"][short] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "ncu_ms": ms,
                                 "registers": int(float(r[hdr.index("launch__registers_per_thread")])),
                                 "dram_pct_of_peak": float(r[hdr.index("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")]),
                                 "warps_active_pct": float(r[hdr.index("sm__warps_active.avg.pct_of_peak_sustained_active")]),
                                 "issue_active_pct": float(r[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
                                 "warp_instructions": float(r[hdr.index("smsp__inst_executed.sum")].replace(",", ""))}
        res["dram_bytes_per_step"] += rd + wr
    txt = json.dumps(res, indent=1)
    if out:
        open(out, "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
