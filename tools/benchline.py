#!/usr/bin/env python
"""one-line digest of a bench.py JSON line read from stdin"""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d.get("roofline", {})
e = d.get("e2e") or {}
print("%-42s value %.4g %s  ms/step %.4f  kernel_ms %.4f  frac %.3f  e2e %.4g" % (
    d["config"]["workload"][:42], d["value"], d["unit"], d["ms_per_step"], r.get("kernel_ms", float("nan")),
    r.get("frac") or float("nan"), e.get("value", float("nan"))))
