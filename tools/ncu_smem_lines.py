#!/usr/bin/env python
"""Shared-memory wavefronts per CUDA source line of an ncu report (needs -lineinfo and --import-source on).

    python tools/ncu_smem_lines.py gpurun_out/prof.ncu-rep [top_n] [kernel-regex] [instances]

Prints, per source line: shared-memory wavefronts, the ideal count and the excess (bank conflicts);
with `instances` also per instance (= per warp of the solve kernel)."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
    if len(sys.argv) > 3 and sys.argv[3]:
        cmd += ["-k", "regex:" + sys.argv[3]]
    inst = float(sys.argv[4]) if len(sys.argv) > 4 else None
    rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
    cur_file, hdr, line_no, line_src, first, active = None, None, None, None, None, False
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            first = first or r[1]
            active = r[1] == first
            continue
        if r[0] == "Line No":
            hdr = r
            iw, ii, ie = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal"), hdr.index("Instructions Executed")
            continue
        if hdr is None or not active:
            continue
        if r[0]:
            line_no, line_src = r[0], r[1]
        if len(r) > iw and r[2]:
            try:
                w, i, n = int(r[iw] or 0), int(r[ii] or 0), int(r[ie] or 0)
            except ValueError:
                continue
            if w == 0:
                continue
            a = agg.setdefault((cur_file, int(line_no)), [0, 0, 0, line_src])
            a[0] += w
            a[1] += i
            a[2] += n
    tot, toti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    print(f"kernel: {first}\nshared wavefronts {tot} (ideal {toti}, excess {tot - toti} = {100 * (tot - toti) / max(tot, 1):.1f}%)"
          + (f"; per instance {tot / inst:.0f} (ideal {toti / inst:.0f})" if inst else ""))
    for (f, l), (w, i, n, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        per = f" per-inst {w / inst:7.1f} ideal {i / inst:7.1f}" if inst else ""
        print(f"{w:10d} {100 * w / tot:5.1f}%  ideal {i:10d}  inst {n:9d}{per}  {f}:{l}: {src.strip()[:80]}")


if __name__ == "__main__":
    main()
