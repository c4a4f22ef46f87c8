#!/usr/bin/env python
"""profiles/r2_roialign_dram.json from an ncu metrics pass over the ROIAlign sweep:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:roialign_staged --csv --log-file gpurun_out/ra_dram.csv python tools/bench_roialign.py --iters 1
    python tools/roialign_dram.py gpurun_out/ra_dram.csv gpurun_out/roialign_sweep.json profiles/r2_roialign_dram.json
Every case of the sweep launches the kernel iters + 2 = 3 times (L2 flushed before each); the LAST launch of a case is
taken.  Times measured under ncu are NOT bench values: only the byte counts are used (bench_roialign.py divides them by
the kernel time it measures itself)."""
import csv
import json
import sys


def main():
    ncu_csv, sweep_json, out = sys.argv[1:4]
    rows = [r for r in csv.reader(l for l in open(ncu_csv) if not l.startswith("==")) if r]
    hdr = rows[0]
    iid, iname, imetric, ival, iunit = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    launches = {}
    for r in rows[1:]:
        d = launches.setdefault(int(r[iid]), {"kernel": r[iname]})
        v = float(r[ival].replace(",", ""))
        u = r[iunit].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1.0)
        d[r[imetric]] = v * scale
    order = [launches[k] for k in sorted(launches)]
    cases = [r["case"] for r in json.load(open(sweep_json))["rows"]]
    per = len(order) // len(cases)
    assert per * len(cases) == len(order), (len(order), len(cases))
    res = {}
    for i, c in enumerate(cases):
        l = order[i * per + per - 1]
        res[c] = {"dram_bytes": l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"], "dram_read": l["dram__bytes_read.sum"],
                  "dram_write": l["dram__bytes_write.sum"], "kernel": l["kernel"].split("(")[0],
                  "us_under_ncu_not_a_bench_value": l["gpu__time_duration.sum"]}
    json.dump(res, open(out, "w"), indent=1)
    print(f"{len(res)} cases -> {out}")


if __name__ == "__main__":
    main()
