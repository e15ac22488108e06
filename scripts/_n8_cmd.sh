N=${NG:-8}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --repeats 3 --no-icp --no-raw --no-extra-configs > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 600 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
for l in open("gpurun_out/r02_bench_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N", d["value"], d["ms_per_step"], d["timed_regions_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("host_cpus"))
PY
