#!/bin/bash
# after the fused map-target kernel and the wide-CTA I/O staging: memcheck over the touched paths, ncu summary of the
# byte-source tile kernel
mkdir -p gpurun_out/prof
P=gpurun_out/prof
T="tests/test_gpu_map_target.py tests/test_gpu_control.py::test_wide_cta_path_single_wave_batch tests/test_gpu_control.py::test_batched_warm_state tests/test_gpu_phik.py::test_phik_matches_oracle"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest $T -m gpu -x -q > $P/sanitizer_memcheck_r02c.log 2>&1
echo "memcheck rc=$?"; tail -3 $P/sanitizer_memcheck_r02c.log
ncu --set full --clock-control none --import-source on -k regex:phik_tma_kernel -s 3 -c 1 -f -o gpurun_out/entropy_fused_r02 \
    python bench.py --workload entropy --steps 3 --warmup 3 > /dev/null 2>&1
python tools/ncu_lsu.py gpurun_out/entropy_fused_r02.ncu-rep 1 > $P/entropy_r02.txt 2>&1
echo "## dynamic SASS opcode mix" >> $P/entropy_r02.txt; python tools/ncu_opmix.py gpurun_out/entropy_fused_r02.ncu-rep 2>/dev/null | head -24 >> $P/entropy_r02.txt
echo "## stall samples per source line" >> $P/entropy_r02.txt; python tools/ncu_lines.py gpurun_out/entropy_fused_r02.ncu-rep 8 >> $P/entropy_r02.txt 2>&1
rm -f gpurun_out/*.ncu-rep
head -24 $P/entropy_r02.txt
