#!/bin/bash
# One GPU-box visit: parity tests, the bench on both workloads, ncu launch list + full capture of the
# dominant kernel.  Usage (from the repo root on the box): bash scripts/gpu_round.sh <tag> [steps]
# Env: SKIP_TESTS=1, SKIP_NCU=1, FULL_BENCH=1 (reference-GPU leg, secondary workloads, LASTZ baseline),
#      NCU_WORKLOADS="syn500 ce11", TESTS="<pytest selection>", PRE="<command run first>"
# Every step has its own timeout and the visit stops at the first failing step: a sticky CUDA error
# otherwise lets the profiler steps hang until gpurun's own limit (40 GPU-minutes lost that way once).
TAG=${1:-rX}; STEPS=${2:-3}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt
free -g | head -2 >> $OUT/${TAG}_gpu.txt; nproc >> $OUT/${TAG}_gpu.txt
if [ -n "$PRE" ]; then bash -c "$PRE" || { echo "PRE step failed"; exit 1; }; fi
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest ${TESTS:-tests} -m gpu -x -q --durations=8 > $OUT/${TAG}_pytest.log 2>&1
  RC=$?
  echo "pytest exit $RC" | tee -a $OUT/${TAG}_pytest.log
  tail -14 $OUT/${TAG}_pytest.log
  [ $RC -ne 0 ] && { grep -E "Error|error|FAILED" $OUT/${TAG}_pytest.log | head -20; exit 1; }
fi
LEAN="--no-cpu-baseline --no-reference-gpu --no-extra"
if [ -n "$FULL_BENCH" ]; then
  timeout 1200 python bench.py --steps $STEPS --warmup 3 > $OUT/${TAG}_bench_syn500.json 2> $OUT/${TAG}_bench_syn500.err
else
  timeout 600 python bench.py --steps $STEPS --warmup 3 $LEAN > $OUT/${TAG}_bench_syn500.json 2> $OUT/${TAG}_bench_syn500.err
fi
RC=$?; echo "bench syn500 exit $RC"; cut -c1-1200 $OUT/${TAG}_bench_syn500.json; tail -3 $OUT/${TAG}_bench_syn500.err
[ $RC -ne 0 ] && exit 1
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench_syn500.json")); r=d["roofline"]
print("SYN500 value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "vec", d["e2e"]["vector_abi"]["value"],
      "frac", r["frac"], "launch_ms", r["avg_launch_ms"], "whole", r["whole_step"]["frac"], "hits/s", d["rates"]["hits_per_s"])
for k in ("reference_gpu", "extra", "cpu_baseline"):
    if d.get(k): print(k, json.dumps(d[k])[:1500])
PY
timeout 600 python bench.py --workload ce11 --steps $STEPS --warmup 3 $LEAN > $OUT/${TAG}_bench_ce11.json 2> $OUT/${TAG}_bench_ce11.err
RC=$?; echo "bench ce11 exit $RC"; tail -3 $OUT/${TAG}_bench_ce11.err
[ $RC -ne 0 ] && exit 1
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench_ce11.json")); r=d["roofline"]
print("CE11 value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "vec", d["e2e"]["vector_abi"]["value"],
      "frac", r["frac"], "launch_ms", r["avg_launch_ms"], "whole", r["whole_step"]["frac"], "hits/s", d["rates"]["hits_per_s"])
PY
if [ -n "$VARIANTS" ]; then   # "name:nvcc flags;name:flags": rebuild on the box and bench each variant (lean), then restore
  IFS=';' read -ra VARS <<< "$VARIANTS"
  for V in "${VARS[@]}"; do
    NAME=${V%%:*}; FLAGS=${V#*:}
    SEGALIGN_B200_NVCC_EXTRA="$FLAGS" python -c "from segalign_b200.build import build_backend; build_backend(force=True)" || exit 1
    for W in syn500 ce11; do
      timeout 400 python bench.py --workload $W --steps $STEPS --warmup 3 $LEAN > $OUT/${TAG}_bench_${W}_$NAME.json 2> $OUT/${TAG}_bench_${W}_$NAME.err || { echo "variant $NAME $W failed"; tail -3 $OUT/${TAG}_bench_${W}_$NAME.err; continue; }
      python -c "
import json
d=json.load(open('$OUT/${TAG}_bench_${W}_$NAME.json')); r=d['roofline']
print('VARIANT $NAME $W value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', r['frac'], 'launch_ms', r['avg_launch_ms'], 'whole', r['whole_step']['frac'])"
    done
  done
  python -c "from segalign_b200.build import build_backend; build_backend(force=True)"
fi
if [ -z "$SKIP_NCU" ]; then
  for W in ${NCU_WORKLOADS:-syn500}; do
    QMB=8; [ "$W" = "syn500" ] && QMB=4
    ARGS="--workload $W --steps 1 --warmup 1 $LEAN --query-mb $QMB --roofline-launches 8 --acct-launches 2 --vector-steps 1"
    timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_${W}_launches.csv \
        python bench.py $ARGS > $OUT/${TAG}_${W}_ncu_bench.log 2>&1 || { echo "ncu launch list failed"; exit 1; }
    python profiles/launch_summary.py $OUT/${TAG}_${W}_launches.csv > $OUT/${TAG}_${W}_launches_summary.txt; head -8 $OUT/${TAG}_${W}_launches_summary.txt
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_filter_hits3 -s 12 -c 2 -f -o $OUT/${TAG}_${W}_filter \
        python bench.py $ARGS > $OUT/${TAG}_${W}_ncu_full.log 2>&1 || { echo "ncu full capture failed"; exit 1; }
    ncu -i $OUT/${TAG}_${W}_filter.ncu-rep --page raw --csv > $OUT/${TAG}_${W}_filter_raw.csv 2>/dev/null
    python profiles/ncu_extract.py $OUT/${TAG}_${W}_filter_raw.csv > $OUT/${TAG}_${W}_k_filter_hits3_ncu_summary.txt; cat $OUT/${TAG}_${W}_k_filter_hits3_ncu_summary.txt
  done
fi
