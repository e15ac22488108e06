timeout 300 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q > gpurun_out/r02_peer_test.txt 2>&1
tail -3 gpurun_out/r02_peer_test.txt
for store in fresh shared; do
GSX_BENCH_STORE=$store timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 3 --no-icp --no-raw --no-extra-configs > gpurun_out/r02_bench_n2_$store.json 2> gpurun_out/r02_bench_n2_$store.err
tail -c 300 gpurun_out/r02_bench_n2_$store.err
python - <<PY
import json
for l in open("gpurun_out/r02_bench_n2_$store.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$store", d["value"], d["ms_per_step"], d["timed_regions_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
done
