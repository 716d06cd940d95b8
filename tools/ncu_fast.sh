#!/bin/bash
# one ncu --set full capture of k_fast_cells (a 128-frame launch of the bench workload) -> gpurun_out/<tag>_fast.ncu-rep
tag=${1:-q}
cmd="python bench.py --steps 2 --warmup 1 --no-cpu --no-rgbd --no-refgpu --no-cfg5 --no-configs --no-parity --sustain-s 0"
ncu --set full --import-source on --clock-control none --kernel-name regex:'k_fast_cells' --launch-skip 6 --launch-count 1 \
    -f -o gpurun_out/${tag}_fast $cmd > gpurun_out/${tag}_ncu_fast.log 2>&1
tail -3 gpurun_out/${tag}_ncu_fast.log
