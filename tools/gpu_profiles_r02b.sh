#!/bin/bash
# round-2 profile refresh after the phi_k final-sum / PDL changes: launch list of the default bench, ncu --set full of
# the phi_k TMA kernel and of the C2 solve kernel (summarised to text on the box)
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
ncu --set full --clock-control none --import-source on -k regex:phik_tma_kernel -s 3 -c 1 -f -o gpurun_out/phik_c3_r02 \
    env EB_PTIME_ALGOS=4 python tools/ptime.py 8192 32 > $P/ncu_phik_r02.log 2>&1
summarise gpurun_out/phik_c3_r02.ncu-rep phik_c3_r02 1
bash tools/ncu_solve.sh r02 c2 > /dev/null 2>&1
summarise gpurun_out/solve_c2_r02.ncu-rep solve_c2_r02 4096
rm -f gpurun_out/*.ncu-rep gpurun_out/ncu_*.log
EB_LIB_PATH=variants/lib_pttrace.so python tools/phik_trace.py 8192 2>&1 | tail -9 > $P/phik_timeline_r02.txt
ls -la $P
