"""Builds tuning variants of libgsx.so (same ABI, different -D launch-configuration macros) into
gradslam_b200/_lib/variants/ for scripts/tune.py.  Usage: python scripts/build_variants.py name=-DX=1,-DY=2 ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gradslam_b200 import build as b

out_dir = os.path.join(ROOT, "gradslam_b200", "_lib", "variants")
os.makedirs(out_dir, exist_ok=True)
for f in os.listdir(out_dir):
    os.remove(os.path.join(out_dir, f))
procs = []
for spec in sys.argv[1:]:
    name, flags = spec.split("=", 1)
    out = os.path.join(out_dir, "libgsx_%s.so" % name)
    cmd = ["/usr/local/cuda/bin/nvcc"] + b.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-o", out] + b.sources()
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    out, _ = p.communicate()
    print(name, "ok" if p.returncode == 0 else "FAILED\n" + out)
