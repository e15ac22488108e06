timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_l.txt 2>&1
tail -3 gpurun_out/r02_gputest_l.txt
timeout 300 python scripts/config3_breakdown.py > gpurun_out/r02_config3_breakdown2.txt 2>&1
grep -E "forward|ms$" gpurun_out/r02_config3_breakdown2.txt | head -30
timeout 600 python bench.py > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err
tail -c 300 gpurun_out/r02_bench_n1_b.err
python - <<PY
import json
for l in open("gpurun_out/r02_bench_n1_b.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d.get("config2_b1_l32",{}) and d["config2_b1_l32"].get("ms_per_step"))
PY
