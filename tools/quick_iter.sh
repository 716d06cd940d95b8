#!/bin/bash
# Quick GPU iteration (via gpurun): the parity tier that covers the extraction kernels, then the stage times of the
# bench workload.   usage: bash tools/quick_iter.sh <tag> [pytest -k expression]
set -u
tag=${1:-q}
kexpr=${2:-}
out=gpurun_out
mkdir -p $out
if [ -n "$kexpr" ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_contract.py -x -q -m gpu -k "$kexpr" > $out/${tag}_tests.log 2>&1
else
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_contract.py -x -q -m gpu > $out/${tag}_tests.log 2>&1
fi
tail -3 $out/${tag}_tests.log
timeout 600 python bench.py --no-cpu --no-rgbd --no-refgpu --no-cfg5 --no-configs --sustain-s 0 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python - <<PY
import json
for line in open("$out/${tag}_bench.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "stages", {k: round(v, 4) for k, v in d["stages_ms"].items()}, "parity", d.get("parity", {}).get("failed"))
PY
