#!/bin/bash
# 2-GPU visit: (1) the default bench on one GPU with every secondary workload (configs[0], [1], [3]-scale) and the
# reference-kernel byte-compares, (2) torchrun --strong: ONE query block, its SeedAndFilter calls split statically
# over the ranks, union checked against one GPU.  Usage: bash scripts/gpu_strong.sh <tag> <N> [steps]
TAG=${1:-rX}; N=${2:-2}; STEPS=${3:-3}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n1_extras.json 2> $OUT/${TAG}_bench_n1_extras.err
RC=$?; echo "bench extras exit $RC"; tail -3 $OUT/${TAG}_bench_n1_extras.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_n1_extras.json"))
    print("value", d["value"], "frac", d["roofline"]["frac"], "reference_gpu", {k: d["reference_gpu"].get(k) for k in ("seconds","ours_seconds","speedup","identical")})
    for k, v in d["extra"].items(): print(k, {a: b for a, b in v.items() if a != "workload"})
except Exception as e:
    print("no result", e)
PY
LEAN="--no-cpu-baseline --no-reference-gpu --no-extra"
for W in syn500 ce11; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --workload $W --gpus $N --strong --steps $STEPS --warmup 3 $LEAN > $OUT/${TAG}_bench_${W}_strong_n$N.json 2> $OUT/${TAG}_bench_${W}_strong_n$N.err
  echo "strong $W exit $?"; tail -3 $OUT/${TAG}_bench_${W}_strong_n$N.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_${W}_strong_n$N.json"))
    print("$W strong n=$N value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "scaling", d["scaling"], d.get("sharded_vs_one_gpu"))
except Exception as e:
    print("$W strong no result", e)
PY
done
