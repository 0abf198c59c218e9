#!/bin/bash
# End-of-round evidence on one GPU box: every GPU test, smoke(), compute-sanitizer (memcheck on smoke, racecheck on the
# repeat / self-alignment goldens), the whole-genome driver at configs[4]/10 scale, the default bench (all legs) and the
# reference arm as the driver runs them.   Usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; RC=$?; echo "pytest exit $RC"; tail -2 $OUT/${TAG}_pytest.log
[ $RC -ne 0 ] && { grep -E "Error|FAILED" $OUT/${TAG}_pytest.log | head; exit 1; }
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/${TAG}_smoke.log
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
  echo "---- racecheck: repeats_entropy + self_align goldens, fast path (k_filter_hits3, k_extend_wide, k_extend_hits, k_finalize_small)"
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x \
      -k "test_device_seeding_matches_reference_golden and (repeats_entropy or self_align)" 2>&1 | tail -12
  echo "---- memcheck: runs of N under --ambiguous (zero-run planes of stage B), repeat-masker case against the CPU restatement"
  # (not the live cases: they start oracle_runner, and the reference's own find_hsps reads its count_del[] array out of
  #  frame -- memcheck follows child processes and stops on that; --target-processes application-only does not
  #  attach to the venv's python at all)
  timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_repeat_masker.py -m gpu -q -x \
      -k "across_n_runs and (default or lane_pair)" 2>&1 | tail -6
  if [ -n "$RACECHECK_N_RUNS" ]; then   # ~9 GPU-minutes
    timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_live_reference_gpu.py -m gpu -q -x \
        -k "n_runs_iupac and (default or WIDE)" 2>&1 | tail -8
  fi ) > $OUT/${TAG}_compute_sanitizer.txt 2>&1
echo "sanitizer done"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/${TAG}_compute_sanitizer.txt
timeout 900 python scripts/run_cli_scale.py > $OUT/${TAG}_cli_scale.json 2> $OUT/${TAG}_cli_scale.err; echo "cli scale exit $?"; cut -c1-900 $OUT/${TAG}_cli_scale.json
timeout 1200 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference_arm_n1.json 2> $OUT/${TAG}_bench_reference_arm_n1.err; echo "reference arm exit $?"
python -c "
import json
d=json.load(open('$OUT/${TAG}_bench_n1.json')); r=d['roofline']
print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'vec', d['e2e']['vector_abi']['value'], 'frac', r['frac'], r['avg_launch_ms'], 'whole', r['whole_step']['frac'], d['clocks'])
print('reference_gpu', {k:d['reference_gpu'][k] for k in ('seconds','ours_seconds','speedup','identical')})
print('extra', {k:(v['value'], v['ms_per_step']) for k,v in d['extra'].items()})
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
d=json.load(open('$OUT/${TAG}_bench_reference_arm_n1.json')); print('reference arm', d['value'], d.get('wall_s'), d['cpu_baseline']['sample'][:160])
"
