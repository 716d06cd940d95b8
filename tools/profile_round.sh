#!/bin/bash
# End-of-round evidence run (on the GPU box, via gpurun): bench without ncu first, then the ncu passes of the same
# command.  Outputs land in gpurun_out/ and are summarised into profiles/ by tools/ncu_summary.py / ncu_lines.py /
# ncu_wavefronts.py.   usage: bash tools/profile_round.sh <tag>      e.g. r02
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
python bench.py --impl reference > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench.err
cmd="python bench.py --steps 2 --warmup 1 --no-cpu --no-rgbd --no-refgpu --no-cfg5 --no-configs --no-parity --sustain-s 0"
$cmd > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launch_list.csv $cmd > /dev/null 2>&1
# one full capture of every extraction kernel of a timed step (the stage-interface pass launches each kernel over all 256 frames)
ncu --set full --import-source on --clock-control none \
    --kernel-name regex:'k_level0|k_resize_rows|k_fast_cells|k_octree|k_blur|k_angle_orb' --launch-skip 48 --launch-count 14 \
    -f -o $out/${tag}_all $cmd > $out/${tag}_ncu_all.log 2>&1
# the matcher (cfg 5 geometry: one rank's 257 k descriptors against the 50 k map)
# (the default kernel: tcgen05 form; ORBB_MATCH_UMMA=0 / ORBB_MATCH_POPC=1 + regex k_match_imma / '^k_match$' for the other two)
ncu --set full --import-source on --clock-control none --kernel-name regex:'k_match_umma' --launch-count 1 \
    -f -o $out/${tag}_match python tools/matcher_probe.py > $out/${tag}_ncu_match.log 2>&1
UMMA_PROBE_NOTIME=1 python tools/umma_probe.py > $out/${tag}_umma_probe.txt 2>&1
[ -x tools/_build/pipe_probe ] && ./tools/_build/pipe_probe > $out/${tag}_pipe_probe.txt 2>&1
python tools/single_frame_probe.py 300 > $out/${tag}_single_frame.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/${tag}_single_frame_launches.csv \
    python tools/single_frame_probe.py 3 > /dev/null 2>&1
python tools/latency_probe.py > $out/${tag}_latency_probe.txt 2>&1
python tools/stage_latency_probe.py > $out/${tag}_stage_latency.txt 2>&1
ORBB_STAGE_PROF=1 python tools/stage_latency_probe.py 2>&1 | grep 'stage prof' | tail -4 >> $out/${tag}_stage_latency.txt
ls -la $out | tail -14
