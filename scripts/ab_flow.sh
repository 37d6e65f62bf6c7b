#!/bin/bash
# Run ON THE GPU BOX: parity of the dataflow chain, then same-box A/B of ms_per_step (alternating runs).
# Usage: scripts/ab_flow.sh <tag> "<name>:<ENV=VAL> ..." [rounds] [workload]
set -u
TAG=${1:-t}
VARIANTS=${2:-"flow2:SIB_HUBERT_FLOW=2 flow0:SIB_HUBERT_FLOW=0"}
ROUNDS=${3:-2}
WORK=${4:-cfg2}
OUT=gpurun_out
mkdir -p $OUT
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 1200 python -m pytest tests/test_gpu_bf16.py -x -q -k "dataflow or flow_chain or test_hubert_bf16" > $OUT/flow_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -5 $OUT/flow_tests_$TAG.log
fi
for i in $(seq 1 $ROUNDS); do
  for v in $VARIANTS; do
    name=${v%%:*}; setting=${v#*:}
    env $setting timeout 400 python bench.py --workload $WORK --steps 20 --warmup 5 --no-cpu-baseline > $OUT/ab_${TAG}_${name}_$i.json 2> $OUT/ab_${TAG}_${name}_$i.err
    python - "$OUT/ab_${TAG}_${name}_$i.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms_per_step", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), "sm_mhz", d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
  done
done
