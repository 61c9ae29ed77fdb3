#!/usr/bin/env python
"""Instruction mix and stall samples per SASS opcode of one kernel of an `ncu --set full --import-source on` report.

    ncu_sass_mix.py report.ncu-rep kernel-regex [top]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main(path, pattern, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + pattern,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][1][:160])
    hdr = rows[1]
    i_src, i_s, i_ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) > i_ie and r[i_ie].isdigit():
            body.append(r)
    tot_s = sum(int(r[i_s] or 0) for r in body) or 1
    tot_i = sum(int(r[i_ie]) for r in body) or 1
    inst, samp = collections.Counter(), collections.Counter()
    stalls = collections.Counter()
    for r in body:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_src])
        k = m.group(2).split(".")[0] if m else r[i_src][:20]
        inst[k] += int(r[i_ie])
        samp[k] += int(r[i_s] or 0)
        for i in stall:
            if r[i]:
                stalls[hdr[i]] += int(r[i])
    print("warp instructions %d, samples %d" % (tot_i, tot_s))
    for k, v in inst.most_common(top):
        print("%-12s inst %10d (%4.1f%%)  samples %6d (%4.1f%%)" % (k, v, 100.0 * v / tot_i, samp[k], 100.0 * samp[k] / tot_s))
    print("stall reasons:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot_s) for k, v in stalls.most_common(8)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
