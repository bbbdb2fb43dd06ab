#!/bin/bash
mkdir -p gpurun_out
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_all_n1.json 2> gpurun_out/bench_all_n1.err ) 2>&1 | tail -3
tail -3 gpurun_out/bench_all_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_all_n1.json').read().strip().splitlines()[-1])
def row(r):
    ro=r.get('roofline') or {}; e=r.get('e2e') or {}; c=r.get('clocks') or {}
    print("%-8s value %.4g %s  ms/step %.4f  frac %s  e2e %.4g (%.3f ms)  clocks %s/%s n=%s %s  cpu %s" % (
        r.get('workload','c2'), r.get('value',float('nan')), r.get('unit',''), r.get('ms_per_step',float('nan')),
        ("%.3f"%ro['frac']) if ro.get('frac') is not None else None, e.get('value',float('nan')), e.get('ms_per_step',float('nan')),
        c.get('sm_mhz'), c.get('sm_max_mhz'), c.get('samples'), c.get('reasons'), (r.get('cpu_baseline') or {}).get('value')))
    if 'error' in r: print("   ERROR", r['error'])
    if 'parity' in r: print("   parity", r['parity'])
row(d)
for r in d.get('secondary',[]): row(r)
PY
python bench.py --workload dwa --steps 20 --warmup 3 2>&1 | tail -1 | python tools/benchline.py
python bench.py --workload collide --steps 20 --warmup 3 2>&1 | tail -1 | python tools/benchline.py
timeout 300 python -m pytest tests/test_cpp_adapter.py tests/test_gpu_dwa.py tests/test_gpu_golden_avoid.py -m gpu -x -q 2>&1 | tail -3
