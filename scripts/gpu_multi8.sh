TAG=r3n; N=8; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/${TAG}_gpus.txt; nproc >> $OUT/${TAG}_gpus.txt; free -g | head -2 >> $OUT/${TAG}_gpus.txt
timeout 300 python -m pytest tests/test_multi_gpu_inproc.py -m gpu -x -q > $OUT/${TAG}_pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -3 $OUT/${TAG}_pytest_multi.log
LEAN="--no-cpu-baseline --no-reference-gpu --no-extra"
timeout 400 python bench.py --workload syn500 --gpus $N --inproc --steps 3 --warmup 2 $LEAN > $OUT/${TAG}_bench_syn500_inproc_n$N.json 2> $OUT/${TAG}_bench_syn500_inproc_n$N.err; echo "inproc syn500 exit $?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --workload syn500 --gpus $N --steps 3 --warmup 2 $LEAN > $OUT/${TAG}_bench_syn500_torchrun_n$N.json 2> $OUT/${TAG}_bench_syn500_torchrun_n$N.err; echo "torchrun syn500 exit $?"
timeout 300 python bench.py --workload ce11 --gpus $N --inproc --steps 3 --warmup 2 $LEAN > $OUT/${TAG}_bench_ce11_inproc_n$N.json 2> $OUT/${TAG}_bench_ce11_inproc_n$N.err; echo "inproc ce11 exit $?"
python - <<PY
import json
for w,mode in (("syn500","inproc"),("syn500","torchrun"),("ce11","inproc")):
    try:
        d=json.load(open("$OUT/${TAG}_bench_%s_%s_n$N.json" % (w,mode)))
        print(w, mode, "n=$N value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "vec", d["e2e"]["vector_abi"]["value"], d["scaling"], d.get("calls_per_gpu"), d["setup_ms"], d["config"]["host_threads"])
    except Exception as e:
        print(w, mode, "no result", e)
PY
