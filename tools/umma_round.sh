#!/bin/bash
# GPU box: tcgen05 matcher probe (both descriptor variants against the POPC matcher), then -- with the variant that
# matched -- the matcher tests of the GPU tier and a short bench run (matcher block, cfg 5 time and result hash).
out=gpurun_out
mkdir -p $out
timeout 420 python tools/umma_probe.py > $out/r02g_umma_probe.txt 2>&1
cat $out/r02g_umma_probe.txt
V=0
grep -q "umma(lbo=plane,sbo=128): all" $out/r02g_umma_probe.txt && V=1
[ $V = 0 ] && grep -q "umma(lbo=128,sbo=plane): all" $out/r02g_umma_probe.txt && V=2
echo "variant $V"
if [ $V != 0 ]; then
  ORBB_MATCH_UMMA=$V timeout 200 python -m pytest tests -m gpu -q -k "match" > $out/r02g_umma_tests.log 2>&1
  tail -4 $out/r02g_umma_tests.log
  ORBB_MATCH_UMMA=$V timeout 240 python bench.py --no-cpu --no-rgbd --no-refgpu --no-configs --sustain-s 0 > $out/r02g_umma_bench.json 2> $out/r02g_umma_bench.err
  python - <<PY
import json
for line in open("$out/r02g_umma_bench.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("value", round(d["value"]), "matcher", json.dumps(d.get("matcher"))[:400])
        print("cfg5", json.dumps(d.get("cfg5"))[:1200])
PY
fi
