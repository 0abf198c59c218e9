run() { name=$1; flags=$2; shift 2
  SEGALIGN_B200_NVCC_EXTRA="$flags" python -c "from segalign_b200.build import build_backend; build_backend(force=True)" || { echo "$name build failed"; return; }
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1v_$name.json 2> gpurun_out/r1v_$name.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r1v_$name.json')); print('$name', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['e2e']['value'])
except Exception as e: print('$name FAILED', e)"
}
run t192_c3 "-DSA_SCR_THREADS=192 -DSA_SCR_MIN_CTAS=4 -DSA_SCR_STAGE_STRIDE=6 -DSA_SCR_Q_CAP=64" SEGALIGN_B200_FILTER_CTAS=3
run t192_c2 "-DSA_SCR_THREADS=192 -DSA_SCR_MIN_CTAS=4 -DSA_SCR_STAGE_STRIDE=6 -DSA_SCR_Q_CAP=64" SEGALIGN_B200_FILTER_CTAS=2
run t192_c4 "-DSA_SCR_THREADS=192 -DSA_SCR_MIN_CTAS=4 -DSA_SCR_STAGE_STRIDE=6 -DSA_SCR_Q_CAP=64" SEGALIGN_B200_FILTER_CTAS=4
run t256_c2_s24 "" SEGALIGN_B200_STREAMS=8 
python -m pytest tests/test_parity_gpu.py -q -m gpu -k "tiny" 2>&1 | tail -3
