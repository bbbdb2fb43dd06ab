#!/bin/bash
# round-2 profile pass: launch list of the default bench, ncu --set full of the final kernels (summarised to text on
# the box: the reports themselves are too large to travel), racecheck
mkdir -p gpurun_out/prof
P=gpurun_out/prof
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $P/launches_r02.csv \
    python bench.py --steps 3 --warmup 3 --loop-steps 20 > $P/bench_under_ncu.log 2>&1
summarise() {  # report tag instances
  python tools/ncu_lsu.py $1 $3 > $P/$2.txt 2>&1
  echo "## dynamic SASS opcode mix" >> $P/$2.txt; python tools/ncu_opmix.py $1 2>/dev/null | head -40 >> $P/$2.txt
  echo "## shared-memory wavefronts per source line" >> $P/$2.txt; python tools/ncu_smem_lines.py $1 16 "" $3 >> $P/$2.txt 2>&1
  echo "## stall samples per source line" >> $P/$2.txt; python tools/ncu_lines.py $1 14 >> $P/$2.txt 2>&1
  rm -f $1
}
for w in c2:4096 c5:65536 c4:131072; do
  k=${w%%:*}; n=${w#*:}
  bash tools/ncu_solve.sh r02 $k > /dev/null 2>&1
  summarise gpurun_out/solve_${k}_r02.ncu-rep solve_${k}_r02 $n
done
ncu --set full --clock-control none --import-source on -k regex:phik_tma_kernel -s 3 -c 1 -f -o gpurun_out/phik_c3_r02 \
    env EB_PTIME_ALGOS=4 python tools/ptime.py 8192 32 > $P/ncu_phik_r02.log 2>&1
summarise gpurun_out/phik_c3_r02.ncu-rep phik_c3_r02 1
ncu --set full --clock-control none -k regex:entropy_density -s 2 -c 1 -f -o gpurun_out/entropy_r02 python bench.py --workload entropy --steps 3 --warmup 3 > /dev/null 2>&1
python tools/ncu_lsu.py gpurun_out/entropy_r02.ncu-rep > $P/entropy_r02.txt 2>&1; rm -f gpurun_out/entropy_r02.ncu-rep
rm -f gpurun_out/*.ncu-rep gpurun_out/ncu_*.log
T="tests/test_gpu_control.py::test_batched_warm_state tests/test_gpu_control.py::test_replay_memory_branches tests/test_gpu_phik.py::test_phik_matches_oracle tests/test_gpu_collision.py"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 99 python -m pytest $T -m gpu -x -q > $P/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "Race reported|hazard|ERROR SUMMARY|RACECHECK SUMMARY" $P/sanitizer_racecheck.log | sort | uniq -c | sort -rn | head -20
ls -la $P; du -sh gpurun_out
