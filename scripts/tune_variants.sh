#!/bin/bash
# Runs bench.py once per kernel-tuning variant (gradslam_b200/_lib/variants/*.so) and prints the per-kernel table.
for so in gradslam_b200/_lib/libgsx.so gradslam_b200/_lib/variants/*.so; do
  echo "== $so"
  GSX_LIB_PATH=$PWD/$so python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value %.0f frames/s  ms/step %.3f  e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
for k,v in d['kernels'].items(): print('  %-26s avg %.1f us  %.0f GB/s' % (k, v['avg_us'], v['algorithmic_GB_per_s']))
"
done
