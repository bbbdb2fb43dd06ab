#!/bin/bash
# registers / spills of every solve_kernel instantiation: tools/regs.sh [extra nvcc flags]
cd "$(dirname "$0")/../ergodic_exploration_b200/csrc"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xptxas -v "$@" -c -o /dev/null ergodic_b200.cu 2>&1 |
  awk '/Compiling entry function/ {name=$0; sub(/.*function ./,"",name); sub(/. for.*/,"",name)}
       /spill stores/ {sp=$0; sub(/^ +/,"",sp)}
       /Used [0-9]+ registers/ {if (name ~ /solve_kernel|phik_dmma/) {r=$0; sub(/.*Used /,"",r); sub(/ registers.*/,"",r); print name, "regs", r, "|", sp}}' |
  sed 's/_ZN2eb12solve_kernelILi\([01]\)ELi\([0-9]*\)EEEvNS_11SolveParamsE/solve<model \1, NB \2>/'
