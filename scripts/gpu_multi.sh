#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the in-process pool tests, one process with N GPUs (--inproc, the reference's
# model, strong scaling of one fixed query block), and N processes under torchrun (weak scaling, the driver's run).
# Usage: bash scripts/gpu_multi.sh <tag> <N> [steps]
TAG=${1:-rX}; N=${2:-2}; STEPS=${3:-3}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $OUT/${TAG}_gpus.txt; nproc >> $OUT/${TAG}_gpus.txt
nvidia-smi topo -m >> $OUT/${TAG}_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu_inproc.py -m gpu -x -q > $OUT/${TAG}_pytest_multi.log 2>&1
RC=$?; echo "pytest multi exit $RC"; tail -5 $OUT/${TAG}_pytest_multi.log
[ $RC -ne 0 ] && exit 1
LEAN="--no-cpu-baseline --no-reference-gpu --no-extra"
for W in syn500 ce11; do
  timeout 600 python bench.py --workload $W --gpus $N --inproc --steps $STEPS --warmup 3 $LEAN > $OUT/${TAG}_bench_${W}_inproc_n$N.json 2> $OUT/${TAG}_bench_${W}_inproc_n$N.err
  echo "inproc $W exit $?"; tail -2 $OUT/${TAG}_bench_${W}_inproc_n$N.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --workload $W --gpus $N --steps $STEPS --warmup 3 $LEAN > $OUT/${TAG}_bench_${W}_torchrun_n$N.json 2> $OUT/${TAG}_bench_${W}_torchrun_n$N.err
  echo "torchrun $W exit $?"; tail -2 $OUT/${TAG}_bench_${W}_torchrun_n$N.err
  python - <<PY
import json
for mode in ("inproc", "torchrun"):
    try:
        d=json.load(open("$OUT/${TAG}_bench_${W}_%s_n$N.json" % mode))
        print("$W", mode, "n=$N value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "vec", d["e2e"]["vector_abi"]["value"], "scaling", d["scaling"], d.get("calls_per_gpu"), d["setup_ms"])
    except Exception as e:
        print("$W", mode, "no result", e)
PY
done
