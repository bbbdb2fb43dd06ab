#!/usr/bin/env python
"""Dynamic SASS opcode mix of the first kernel in an ncu report (needs
--import-source on): executed warp instructions and stall samples per opcode.

    python tools/ncu_opmix.py report.ncu-rep [kernel-regex] [instances]
"""
import csv
import io
import re
import subprocess
import sys


def main():
    cmd = ["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "sass", "--csv"]
    if len(sys.argv) > 2 and sys.argv[2]:
        cmd += ["-k", "regex:" + sys.argv[2]]
    per = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
    hdr = None
    mix, stall = {}, {}
    for r in rows:
        if r and r[0] == "Address":
            if hdr is not None:
                break  # first kernel only
            hdr = r
            ii, si = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
            continue
        if hdr is None or len(r) <= ii:
            continue
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[1])
        if not m:
            continue
        op = m.group(2)
        mix[op] = mix.get(op, 0) + int(r[ii] or 0)
        stall[op] = stall.get(op, 0) + int(r[si] or 0)
    tot, tots = sum(mix.values()) or 1, sum(stall.values()) or 1
    print(f"total warp instructions {tot} ({tot / per:.1f} per instance), stall samples {tots}")
    for op, n in sorted(mix.items(), key=lambda kv: -kv[1])[:45]:
        print(f"  {op:12s} {n:12d} {100 * n / tot:5.1f}%  per-inst {n / per:8.1f}   stalls {100 * stall[op] / tots:5.1f}%")


if __name__ == "__main__":
    main()
