#!/usr/bin/env python
"""Instruction mix and stall samples per SASS opcode of one kernel of an `ncu --set full --import-source on` report.

    ncu_sass_mix.py report.ncu-rep kernel-regex [top]
    ncu_sass_mix.py lines report.ncu-rep kernel-regex file.sass mangled-name-fragment source.cu [top]
        executed warp instructions and stall samples per SOURCE LINE: file.sass = `nvdisasm -g -c` of the cubin
        (`cuobjdump -xelf all lib.so`) that was profiled; joins on the instruction address
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


def lines(path, pattern, sass, fragment, source, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + pattern,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    i_a, i_s, i_ie = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    body = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) > i_ie and r[i_ie].isdigit():
            body.append(r)
    base = int(body[0][i_a], 16)
    line_of, cur, inside = {}, None, False
    for ln in open(sass):
        if ln.startswith(".text.") or ln.startswith("\t.section"):
            inside = fragment in ln
        if not inside:
            continue
        m = re.search(r'//## File ".*", line (\d+)', ln)
        if m:
            cur = int(m.group(1))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    text = open(source).read().splitlines()
    inst, samp = collections.Counter(), collections.Counter()
    for r in body:
        l = line_of.get(int(r[i_a], 16) - base)
        inst[l] += int(r[i_ie])
        samp[l] += int(r[i_s] or 0)
    ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
    print("warp instructions %d, samples %d" % (ti, ts))
    for l, v in inst.most_common(top):
        src = text[l - 1].strip()[:100] if l and l <= len(text) else "?"
        print("%5s inst %9d (%4.1f%%) samples %5d (%4.1f%%) | %s" % (l, v, 100.0 * v / ti, samp[l], 100.0 * samp[l] / ts, src))


if __name__ == "__main__":
    if sys.argv[1] == "lines":
        lines(*sys.argv[2:7], **({"top": int(sys.argv[7])} if len(sys.argv) > 7 else {}))
        sys.exit(0)
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
