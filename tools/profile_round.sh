#!/bin/bash
# End-of-round evidence run (on the GPU box, via gpurun): bench without ncu first, then the ncu passes of the same
# command.  Outputs land in gpurun_out/ and are summarised into profiles/ by tools/ncu_summary.py / ncu_lines.py.
#   usage: bash tools/profile_round.sh <tag>      e.g. r01m
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
python bench.py --impl reference > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench.err
cmd="python bench.py --steps 2 --warmup 1 --no-cpu --no-rgbd --no-refgpu"
$cmd > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launch_list.csv $cmd > /dev/null 2>&1
ncu --set full --import-source on --clock-control none \
    --kernel-name regex:'k_level0|k_resize_rows|k_fast_cells|k_octree|k_blur|k_angle_orb' --launch-skip 48 --launch-count 14 \
    -f -o $out/${tag}_all $cmd > $out/${tag}_ncu_all.log 2>&1
python tools/single_frame_probe.py 300 > $out/${tag}_single_frame.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/${tag}_single_frame_launches.csv \
    python tools/single_frame_probe.py 3 > /dev/null 2>&1
ls -la $out | tail -12
