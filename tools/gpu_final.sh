#!/bin/bash
# round-end rehearsal on one GPU: the driver's sequence (tests, smoke, reference arm, our arm)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err ) 2>&1 | grep real
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_ours.json 2> gpurun_out/final_ours.err ) 2>&1 | grep real
tail -c 600 gpurun_out/final_ref.json; echo
python tools/benchsum.py gpurun_out/final_ours.json
tail -3 gpurun_out/final_ours.err
