#!/usr/bin/env python
"""DRAM bytes per launch of every kernel from an ncu launch list taken with
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file <csv>
(tools/gpu_profiles.sh).  Writes the JSON that bench.py reads for `roofline.traffic`.

usage: ncu_traffic.py launches.csv out.json pairs_per_gpu
"""
import collections
import csv
import json
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}


def main():
    path, out, pairs = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    ix = {h: i for i, h in enumerate(rows[hi])}
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(ix):
            continue
        key = (r[ix["ID"]], r[ix["Kernel Name"]].split("(")[0])
        per.setdefault(key, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * UNIT[r[ix["Metric Unit"]]]
    # one whole step: from the first line-index launch of an encode (two per step, one per file) to the next encode's
    items = list(per.items())
    starts = [i for i, ((_, name), _) in enumerate(items) if name.startswith("k_index_lines")][::2]
    if len(starts) >= 2:
        items = items[starts[0]:starts[1]]
    elif len(starts) == 1:
        items = items[starts[0]:]                       # the capture ends with the last step: its encode and its decode
    agg = {}
    for (_, name), m in items:
        a = agg.setdefault(name, dict(launches=0, dram_read_bytes=0.0, dram_write_bytes=0.0, us=0.0))
        a["launches"] += 1
        a["dram_read_bytes"] += m.get("dram__bytes_read.sum", 0.0)
        a["dram_write_bytes"] += m.get("dram__bytes_write.sum", 0.0)
        a["us"] += m.get("gpu__time_duration.sum", 0.0)
    total_us = sum(a["us"] for a in agg.values())
    res = dict(pairs_per_gpu=pairs, source="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none (cold cache, serialised launches)",
               kernels={k: dict(launches=a["launches"], traffic_bytes_per_launch=(a["dram_read_bytes"] + a["dram_write_bytes"]) / a["launches"],
                                dram_read_bytes_per_launch=a["dram_read_bytes"] / a["launches"], dram_write_bytes_per_launch=a["dram_write_bytes"] / a["launches"],
                                us_per_launch=a["us"] / a["launches"], share_of_step=a["us"] / total_us) for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"])})
    json.dump(res, open(out, "w"), indent=1)
    for k, v in res["kernels"].items():
        print("%-26s x%-3d %9.1f us  %8.1f MB  share %.3f" % (k, v["launches"], v["us_per_launch"], v["traffic_bytes_per_launch"] / 1e6, v["share_of_step"]))


if __name__ == "__main__":
    main()
