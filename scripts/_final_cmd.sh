timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_final.txt 2>&1
tail -2 gpurun_out/r02_gputest_final.txt
timeout 60 python __graft_entry__.py smoke > gpurun_out/r02_smoke.txt 2>&1
tail -1 gpurun_out/r02_smoke.txt
timeout 200 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err
tail -c 200 gpurun_out/r02_bench_n1_final.err
python - <<PY
import json
for l in open("gpurun_out/r02_bench_n1_final.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --repeats 1 --no-icp --no-raw --no-extra-configs --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_final_b.log 2>&1
tail -2 gpurun_out/r02_ncu_final_b.log | cut -c1-200
