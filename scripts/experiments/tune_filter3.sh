#!/bin/bash
# Build-time tuning sweep of k_filter_hits3 on the GPU box: each variant is rebuilt there and
# benched (short run, no CPU baseline).  Usage: bash scripts/tune_filter3.sh <tag>
TAG=${1:-tune}
OUT=gpurun_out
mkdir -p $OUT
run() { # name, nvcc extra flags, env
  name=$1; flags=$2; shift 2
  SEGALIGN_B200_NVCC_EXTRA="$flags" python -c "from segalign_b200.build import build_backend; build_backend(force=True)" || { echo "$name build failed"; return; }
  env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_$name.json"))
    print("$name", "value", d["value"], "ms/step", d["ms_per_step"], "filter ms/launch", d["roofline"]["avg_launch_ms"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("$name FAILED", e)
PY
}
run base "" X=1
run t288_s6_q48 "-DSA_SCR_THREADS=288 -DSA_SCR_STAGE_STRIDE=6 -DSA_SCR_Q_CAP=48" X=1
run t256_s6 "-DSA_SCR_STAGE_STRIDE=6" X=1
run t256_ctas2 "" SEGALIGN_B200_FILTER_CTAS=2
run t256_q64 "-DSA_SCR_Q_CAP=64" X=1
run streams "" SEGALIGN_B200_STREAMS=6
# restore the default build
python -c "from segalign_b200.build import build_backend; build_backend(force=True)"
