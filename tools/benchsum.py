#!/usr/bin/env python
"""digest of a bench.py JSON line (file or '-' for stdin): the primary workload and every secondary"""
import json
import sys

src = sys.stdin.read() if sys.argv[1] == "-" else open(sys.argv[1]).read()
d = json.loads(src.strip().splitlines()[-1])


def row(r):
    ro, e, c = r.get("roofline") or {}, r.get("e2e") or {}, r.get("clocks") or {}
    name = r.get("workload") or (r.get("config") or {}).get("workload", "?")[:10]
    print("%-8s N=%s %s  value %.4g %s  ms/step %.4f  frac %s  e2e %.4g (%.3f ms)  clocks %s n=%s %s  cpu %s" % (
        name, r.get("n_gpus"), r.get("scaling"), r.get("value", float("nan")), r.get("unit", ""), r.get("ms_per_step", float("nan")),
        ("%.3f" % ro["frac"]) if ro.get("frac") is not None else None, e.get("value", float("nan")), e.get("ms_per_step", float("nan")),
        c.get("sm_mhz"), c.get("samples"), c.get("reasons"), (r.get("cpu_baseline") or {}).get("value")))
    if "strict_barrier" in r:
        pg = r["strict_barrier"]
        print("    strict per-step barrier: value %.4g  ms/step %.4f" % (pg["value"], pg["ms_per_step"]))
    for k in ("error", "parity"):
        if k in r:
            print("   ", k, r[k])


row(d)
for r in d.get("secondary", []):
    row(r)
if "peer_stress" in d:
    print("peer_stress", d["peer_stress"])
