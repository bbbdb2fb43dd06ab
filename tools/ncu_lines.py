#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: stall samples and executed
instructions (needs -lineinfo at compile time and --import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n] [kernel-regex]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
    if len(sys.argv) > 3:
        cmd += ["-k", "regex:" + sys.argv[3]]
    txt = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    cur_file, hdr, line_no, line_src = None, None, None, None
    agg = {}
    first_kernel = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            if first_kernel is None:
                first_kernel = r[1]
            active = r[1] == first_kernel
            continue
        if r[0] == "Line No":
            hdr = r
            si = hdr.index("Warp Stall Sampling (All Samples)")
            ie = hdr.index("Instructions Executed")
            stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or not active:
            continue
        if r[0]:
            line_no, line_src = r[0], r[1]
        if len(r) > ie and r[2]:
            try:
                s, n = int(r[si] or 0), int(r[ie] or 0)
            except ValueError:
                continue
            key = (cur_file, int(line_no))
            a = agg.setdefault(key, [0, 0, line_src, {}])
            a[0] += s
            a[1] += n
            for i, name in stall_cols:
                try:
                    a[3][name] = a[3].get(name, 0) + int(r[i] or 0)
                except (ValueError, IndexError):
                    pass
    tot = sum(a[0] for a in agg.values()) or 1
    toti = sum(a[1] for a in agg.values()) or 1
    print(f"kernel: {first_kernel}\ntotal stall samples {tot}, warp instructions {toti}")
    for (f, l), (s, n, src, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        why = ", ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4] if v)
        print(f"{s:7d} {100 * s / tot:5.1f}%  inst {n:9d} {100 * n / toti:5.1f}%  {f}:{l}: {src.strip()[:70]}\n{'':20s}[{why}]")


if __name__ == "__main__":
    main()
