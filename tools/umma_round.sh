#!/bin/bash
# GPU box: tcgen05 matcher probe (both epilogue forms against the POPC matcher), then -- with the first form that
# matched -- the matcher tests of the GPU tier and a short bench run (matcher block, cfg 5 time and result hash).
out=gpurun_out
tag=${1:-r02h}
mkdir -p $out
timeout 420 python tools/umma_probe.py > $out/${tag}_umma_probe.txt 2>&1
cat $out/${tag}_umma_probe.txt
V=0
grep -q "^umma: all" $out/${tag}_umma_probe.txt && V=1
[ $V = 0 ] && grep -q "^umma(plain epilogue): all" $out/${tag}_umma_probe.txt && V=2
echo "variant $V"
if [ $V != 0 ]; then
  ORBB_MATCH_UMMA=$V timeout 200 python -m pytest tests -m gpu -q -k "match" > $out/${tag}_umma_tests.log 2>&1
  tail -4 $out/${tag}_umma_tests.log
  ORBB_MATCH_UMMA=$V timeout 240 python bench.py --no-cpu --no-rgbd --no-refgpu --no-configs --sustain-s 0 > $out/${tag}_umma_bench.json 2> $out/${tag}_umma_bench.err
  python - <<PY
import json
for line in open("$out/${tag}_umma_bench.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("value", round(d["value"]), "matcher", json.dumps(d.get("matcher"))[:400])
        print("cfg5", json.dumps(d.get("cfg5"))[:1200])
PY
fi
if [ $V != 0 ]; then
  ORBB_MATCH_UMMA=$V timeout 150 ncu --set full --import-source on --clock-control none -k regex:k_match_umma -c 1 -f -o $out/${tag}_match_umma python tools/matcher_probe.py > $out/${tag}_ncu_umma.log 2>&1
  tail -3 $out/${tag}_ncu_umma.log
fi
