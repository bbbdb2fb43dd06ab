#!/usr/bin/env python
"""FP64-pipe cycles per CUDA source line (joins the cuda,sass and sass source
pages of an ncu report by instruction address).  DFMA/DADD/DMUL/DSETP count
2 pipe cycles per warp instruction (16 lanes/SMSP), DMMA.8x8x4 16.

    python tools/ncu_fp64_lines.py report.ncu-rep [instances] [top_n]
"""
import csv
import io
import re
import subprocess
import sys

COST = {"DFMA": 2, "DADD": 2, "DMUL": 2, "DSETP": 2, "DMMA": 16}


def page(rep, src):
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", src, "--csv"]
    return list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))


def main():
    rep = sys.argv[1]
    per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    addr_line = {}
    cur_file, line, src, kernel = None, None, None, None
    for r in page(rep, "cuda,sass"):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            kernel = kernel or r[1]
            active = r[1] == kernel
        elif r[0] == "Line No":
            pass
        elif r[0]:
            line, src = r[0], r[1]
        elif len(r) > 3 and r[2].startswith("0x") and active:
            addr_line.setdefault(r[2], (cur_file, int(line), src.strip()))
    agg, allinst = {}, {}
    hdr = None
    for r in page(rep, "sass"):
        if r and r[0] == "Address":
            if hdr is not None:
                break
            hdr = r
            ii = hdr.index("Instructions Executed")
            continue
        if hdr is None or len(r) <= ii:
            continue
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[1])
        if not m:
            continue
        key = addr_line.get(r[0], ("?", 0, "?"))
        n = int(r[ii] or 0)
        allinst[key] = allinst.get(key, 0) + n
        if m.group(2) in COST:
            agg[key] = agg.get(key, 0) + n * COST[m.group(2)]
    tot = sum(agg.values()) or 1
    print(f"kernel {kernel}\nFP64 pipe cycles {tot / per:.0f} per instance, warp instructions {sum(allinst.values()) / per:.0f} per instance")
    for key, c in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{c / per:8.1f} cyc {100 * c / tot:5.1f}%  inst {allinst[key] / per:7.1f}  {key[0]}:{key[1]}: {key[2][:80]}")


if __name__ == "__main__":
    main()
