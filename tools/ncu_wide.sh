#!/bin/bash
# ncu --set full of the num_basis > 32 kernel (nb 64, 4096 instances), summarised to text on the box
mkdir -p gpurun_out/prof
P=gpurun_out/prof
cat > /tmp/wide_one.py <<'PY'
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import ergodic_exploration_b200 as eb
R, umin, umax = np.diag([1.0, 1.0, 2.0]), [-1.0, -1.0, -2.0], [1.0, 1.0, 2.0]
nb, B = 64, 4096
rng = np.random.default_rng(nb)
ctl = eb.ErgodicControl(eb.Omni(), 0.1, 5.0, 0.1, 1.0, nb, 1000, 100, R, umin, umax, batch=B)
ctl.setTarget([eb.Gaussian([2.5, 2.5], [1.5, 1.5]), eb.Gaussian([8.5, 2.5], [1.5, 1.5])])
ctl.keep_ck(False)
x = np.column_stack([rng.uniform(0.5, 9.5, B), rng.uniform(0.5, 9.5, B), rng.uniform(-np.pi, np.pi, B)])
ctl.set_ut(rng.uniform(umin, umax, size=(B, ctl.steps, 3)) * 0.5)
xd = torch.from_numpy(x).cuda()
u0 = torch.empty((B, 3), dtype=torch.float64, device="cuda")
for _ in range(4):
    ctl.control((0.0, 10.0, 0.0, 10.0), xd, u0=u0)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:solve_kernel_big -s 2 -c 1 -f -o gpurun_out/solve_wide_r02 \
    python /tmp/wide_one.py > $P/ncu_wide.log 2>&1
R=gpurun_out/solve_wide_r02.ncu-rep
python tools/ncu_lsu.py $R 4096 > $P/solve_wide_r02.txt 2>&1
echo "## dynamic SASS opcode mix" >> $P/solve_wide_r02.txt; python tools/ncu_opmix.py $R 2>/dev/null | head -30 >> $P/solve_wide_r02.txt
echo "## shared-memory wavefronts per source line" >> $P/solve_wide_r02.txt; python tools/ncu_smem_lines.py $R 12 "" 4096 >> $P/solve_wide_r02.txt 2>&1
echo "## stall samples per source line" >> $P/solve_wide_r02.txt; python tools/ncu_lines.py $R 14 >> $P/solve_wide_r02.txt 2>&1
rm -f $R
tail -5 $P/ncu_wide.log; head -40 $P/solve_wide_r02.txt
