#!/usr/bin/env python
"""Summarise an `ncu --csv` launch list of one bench step into profiles/*.md and profiles/rNN_dram_traffic.json.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:"pesq_|stoi_" -s 30 -c 10 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu
    python tools/ncu_launches_summary.py gpurun_out/launches.csv profiles/r02_ncu_launches

Writes <out>.csv (the input, ncu banner lines dropped), <out>.md (time share and DRAM bytes per launch) and
profiles/rNN_dram_traffic.json (NN = the round prefix of <out>), which bench.py reads -- newest round first -- for
`roofline.traffic` (keys = the library's profiler names), so the traffic figure comes from a launch list of the same
round's code.
"""
from __future__ import annotations

import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# ncu kernel name prefix -> name used by the library's own profiler / bench.py
ALIASES = {"pesq_filter_tiled_kernel": "pesq_filter_kernel", "pesq_filter_kernel": "pesq_filter_kernel",
           "stoi_resample85_kernel": "stoi_resample_kernel", "stoi_resample_kernel": "stoi_resample_kernel",
           "stoi_energy_from_hops_kernel": "stoi_energy_kernel"}


def main():
    src, out = sys.argv[1], sys.argv[2]
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    launches = {}
    for r in rows:
        e = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        e[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        e[r["Metric Name"] + ".unit"] = r["Metric Unit"]
    order = sorted(launches)
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
    total_us = 0.0
    table = []
    for i in order:
        e = launches[i]
        us = e["gpu__time_duration.sum"] * scale.get(e["gpu__time_duration.sum.unit"], 1.0)
        rd = e["dram__bytes_read.sum"] * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[e["dram__bytes_read.sum.unit"]]
        wr = e["dram__bytes_write.sum"] * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[e["dram__bytes_write.sum.unit"]]
        short = e["name"].replace("void ", "").split("(")[0]
        table.append((short, us, rd, wr))
        total_us += us
    with open(out + ".csv", "w") as f:
        f.writelines(lines)
    md = ["# ncu launch list of one step (8192 x 10 s, one B200): `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
          "dram__bytes_write.sum --clock-control none -k regex:\"pesq_|stoi_\" -s 30 -c 10 python bench.py --steps 1 "
          "--warmup 3 --no-e2e --no-cpu --no-parity --no-graph` (`--no-graph`: the same ten kernels launched directly on one "
          "stream, so that the launch order under ncu is the chain order; the timed bench replays them as one CUDA graph)", "",
          "Cold-cache, serialised launches: compare SHARES with the live CUDA-event times in the same round's bench JSON "
          "(`kernels`).", "", "| kernel | time (us) | share | DRAM read (GB) | DRAM write (GB) |", "|---|---|---|---|---|"]
    traffic = {}
    for short, us, rd, wr in table:
        md.append("| %s | %.1f | %.1f %% | %.3f | %.3f |" % (short, us, 100 * us / total_us, rd / 1e9, wr / 1e9))
        base = short.split("<")[0]
        key = ALIASES.get(base, base)
        traffic[key] = traffic.get(key, 0.0) + rd + wr
    md.append("| **total** | %.1f | | %.1f (read+write) | |" % (total_us, sum(traffic.values()) / 1e9))
    with open(out + ".md", "w") as f:
        f.write("\n".join(md) + "\n")
    import re
    m = re.match(r"(r\d\d)_", os.path.basename(out))
    rnd = m.group(1) if m else "r00"
    with open(os.path.join(ROOT, "profiles", rnd + "_dram_traffic.json"), "w") as f:
        json.dump({"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from %s.csv (ncu, 8192 x 10 s on one "
                               "B200); bench.py scales by items per GPU" % os.path.relpath(out, ROOT),
                   "items": 8192, "bytes_per_launch": traffic}, f, indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
