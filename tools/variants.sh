#!/bin/bash
# build tuning variants of the library: tools/variants.sh name "flags" [name "flags" ...]
# -> variants/lib_<name>.so (git-ignored; travels to the GPU box); run with EB_LIB_PATH=...
cd "$(dirname "$0")/.."
mkdir -p variants
while [ $# -ge 2 ]; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $2 -shared \
    -o variants/lib_$1.so ergodic_exploration_b200/csrc/ergodic_b200.cu -lcudart &
  shift 2
done
wait
ls -la variants
