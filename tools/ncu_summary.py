#!/usr/bin/env python
"""Text summaries of ncu captures for profiles/ (run here, after gpurun brought the files back).

    ncu_summary.py launches gpurun_out/launches.csv            per-kernel totals of a `--metrics gpu__time_duration.sum --csv` list
    ncu_summary.py full gpurun_out/prof.ncu-rep [regex]        selected metrics per launch of a `--set full` report
    ncu_summary.py traffic gpurun_out/prof.ncu-rep regex key [n] [source] [which]
                                                               dram bytes (read + write) of the LAST launch matching regex (which = -2: the
                                                               one before it, ...) -> profiles/ncu_traffic.json[key] (what bench.py reports
                                                               as roofline.traffic)
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg.per_second",
]


def launches(path):
    with open(path) as fh:
        text = fh.read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    total = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        value = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        value *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        n, t = total.get(name, (0, 0.0))
        total[name] = (n + 1, t + value)
    print("per-kernel totals over the captured launches (times are cold-cache and serialised; unit ns)")
    for name, (n, t) in sorted(total.items(), key=lambda kv: -kv[1][1]):
        print("%8d launches %14.1f  %s" % (n, t, name[-90:]))


def full(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    ki = head.index("Kernel Name")
    for r in rows[2:]:
        if pattern and not re.search(pattern, r[ki]):
            continue
        print("---- " + r[ki][:200])
        for m in METRICS:
            if m in head:
                print("  %-92s %s %s" % (m, r[head.index(m)], units[head.index(m)]))


def traffic(path, pattern, key, n=None, source=None, which=-1):
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    ki = head.index("Kernel Name")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    matches = []
    for r in rows[2:]:
        if re.search(pattern, r[ki]):
            total = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = head.index(m)
                total += float(r[i].replace(",", "")) * scale[units[i]]
            matches.append((r[ki], total))
    found = matches[int(which)] if len(matches) >= abs(int(which)) else None
    if found is None:
        sys.exit("no launch matches %r" % pattern)
    dest = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    data = {}
    if os.path.exists(dest):
        with open(dest) as fh:
            data = json.load(fh)
    entry = {"bytes": int(found[1]), "kernel": re.sub(r"\(.*", "", found[0])[-120:], "source": source or os.path.basename(path)}
    if n not in (None, "", "-"):
        entry["n"] = int(n)
    data[key] = entry
    with open(dest, "w") as fh:
        json.dump(data, fh, indent=1)
    print(key, entry)


if __name__ == "__main__":
    if len(sys.argv) >= 5 and sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else None,
                sys.argv[6] if len(sys.argv) > 6 else None, sys.argv[7] if len(sys.argv) > 7 else -1)
    elif len(sys.argv) >= 3 and sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif len(sys.argv) >= 3 and sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        sys.exit(__doc__)
