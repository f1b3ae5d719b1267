"""profiles/traffic.json from ncu CSV logs: per kernel and workload, dram__bytes_read.sum + dram__bytes_write.sum and
gpu__time_duration.sum of ONE launch (the mean over the launches of that kernel in the log).  bench.py reads the file for
the `traffic` field of its roofline objects.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \\
        --log-file gpurun_out/x.csv -k regex:<kernel> python bench.py ...          (on the GPU box)
    python tools/ncu_traffic.py <kernel-substring>@<shape>=gpurun_out/x.csv ...     (here)
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def parse(path, kernel_sub):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if kernel_sub in r.get("Kernel Name", ""):
            rows.append(r)
    per = {}
    for r in rows:
        key = r["ID"]
        val = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        per.setdefault(key, {})[r["Metric Name"]] = val
    launches = [v for v in per.values() if "dram__bytes_read.sum" in v]
    if not launches:
        raise SystemExit("no launches of %s in %s" % (kernel_sub, path))
    n = len(launches)
    return {"dram_bytes_read": sum(v["dram__bytes_read.sum"] for v in launches) / n,
            "dram_bytes_write": sum(v["dram__bytes_write.sum"] for v in launches) / n,
            "gpu_time_ms": sum(v.get("gpu__time_duration.sum", 0.0) for v in launches) / n,
            "launches_in_log": n, "source": os.path.relpath(path, ROOT)}


def main():
    out_path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(out_path) as f:
            out = json.load(f)
    except Exception:
        out = {}
    for arg in sys.argv[1:]:
        key, path = arg.split("=", 1)
        kernel = key.split("@")[0]
        out[key] = parse(path, kernel)
        print(key, out[key])
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
