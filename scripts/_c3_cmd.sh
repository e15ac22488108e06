timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_k.txt 2>&1
tail -3 gpurun_out/r02_gputest_k.txt
timeout 300 python scripts/config3_breakdown.py > gpurun_out/r02_config3_breakdown.txt 2>&1
grep -E "forward|ms$" gpurun_out/r02_config3_breakdown.txt | head -30
