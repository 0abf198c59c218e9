#!/bin/bash
# End-of-round evidence on one GPU box: parity tests, the default bench (with the LASTZ CPU baseline),
# the reference arm, the ncu launch list + one full capture of the dominant kernel.
# Usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference > $OUT/${TAG}_bench_reference_arm_n1.json 2> $OUT/${TAG}_bench_reference_arm_n1.err; echo "reference arm exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --query-mb 8 --roofline-launches 8 > $OUT/${TAG}_ncu_bench.log 2>&1
python profiles/launch_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt; head -8 $OUT/${TAG}_launches_summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_filter_hits -s 20 -c 2 -f -o $OUT/${TAG}_filter \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --query-mb 8 --roofline-launches 8 > $OUT/${TAG}_ncu_full.log 2>&1
ncu -i $OUT/${TAG}_filter.ncu-rep --page raw --csv > $OUT/${TAG}_filter_raw.csv 2>/dev/null
python profiles/ncu_extract.py $OUT/${TAG}_filter_raw.csv > $OUT/${TAG}_filter_ncu_summary.txt
python -c "
import json
for n in ('bench_n1','bench_reference_arm_n1'):
    d=json.load(open('$OUT/${TAG}_'+n+'.json')); print(n, d['value'], d.get('ms_per_step'), d.get('e2e'), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'))
"
