#!/bin/bash
# One GPU-box visit: parity tests, bench per filter kernel, ncu launch list + full capture of the
# dominant kernel.  Usage (from the repo root on the box): bash scripts/gpu_round.sh <tag> [kernels...]
TAG=${1:-rX}; shift
KERNELS=${@:-3 2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_gpu.txt
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
if [ -n "$CHECK_LARGE" ]; then
  timeout 900 python tests/golden/check_large.py $CHECK_LARGE > $OUT/${TAG}_check_large.json 2> $OUT/${TAG}_check_large.err
  echo "check_large exit $?"; cat $OUT/${TAG}_check_large.json | cut -c1-700
fi
for K in $KERNELS; do
  SEGALIGN_B200_FILTER_KERNEL=$K timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
      > $OUT/${TAG}_bench_k$K.json 2> $OUT/${TAG}_bench_k$K.err
  echo "bench k=$K exit $?"; cat $OUT/${TAG}_bench_k$K.json | cut -c1-600
done
if [ -n "$EXTRA_BUILDS" ]; then   # "name:flags;name:flags": rebuild on the box and bench each variant
  IFS=';' read -ra VARS <<< "$EXTRA_BUILDS"
  for V in "${VARS[@]}"; do
    NAME=${V%%:*}; FLAGS=${V#*:}
    SEGALIGN_B200_NVCC_EXTRA="$FLAGS" python -c "from segalign_b200.build import build_backend; build_backend(force=True)"
    timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$NAME.json 2> $OUT/${TAG}_bench_$NAME.err
    echo "variant $NAME ($FLAGS): $(python -c "import json;d=json.load(open('$OUT/${TAG}_bench_$NAME.json'));print(d['value'],d['ms_per_step'],d['roofline']['avg_launch_ms'],d['e2e']['value'])")"
  done
  python -c "from segalign_b200.build import build_backend; build_backend(force=True)"
fi
K=${KERNELS%% *}
export SEGALIGN_B200_FILTER_KERNEL=$K
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --query-mb 8 --roofline-launches 8 > $OUT/${TAG}_ncu_bench.log 2>&1
python profiles/launch_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt; cat $OUT/${TAG}_launches_summary.txt | head -8
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_filter_hits -s 20 -c 2 -f -o $OUT/${TAG}_filter \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --query-mb 8 --roofline-launches 8 > $OUT/${TAG}_ncu_full.log 2>&1
ncu -i $OUT/${TAG}_filter.ncu-rep --page raw --csv > $OUT/${TAG}_filter_raw.csv 2>/dev/null
python profiles/ncu_extract.py $OUT/${TAG}_filter_raw.csv > $OUT/${TAG}_filter_ncu_summary.txt; cat $OUT/${TAG}_filter_ncu_summary.txt
