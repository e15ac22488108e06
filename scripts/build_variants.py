"""Builds tuning variants of libgsx.so into gradslam_b200/_lib/variants/ (picked up by scripts/tune.py).

    python scripts/build_variants.py                      # the round-2 candidates below
    python scripts/build_variants.py name=-DFLAG=1,-DX=2  # ad-hoc variants

Variants are compile-time switches of the same sources (see the macros at the top of csrc/gsx_fusion.cu); the product
library gradslam_b200/_lib/libgsx.so is not touched.  To check that a variant keeps the maps bit-identical run the GPU
tests against it:   GSX_LIB_PATH=gradslam_b200/_lib/variants/libgsx_<name>.so python -m pytest tests -m gpu -q"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gradslam_b200.build import NVCC_FLAGS, SRC_DIR  # noqa: E402

CANDIDATES = {
    "fasttest": ["-DGSX_K2_FASTTEST=1"],                       # decision-exact shortcuts in K2
    "fasttest_mb3": ["-DGSX_K2_FASTTEST=1", "-DGSX_K2_MINB=3"],  # the same without spills (85 registers)
    "geo_mb3": ["-DGSX_K4_GEO_MINB=3"],                        # geo32 layout study, K4 without spills
    "noexp_nodiv": ["-DGSX_K4_FAST_EXP=0", "-DGSX_K4_CTA_DIV=0"],  # the round-1 baseline of K4's instruction cuts
}


def build(name, flags):
    out_dir = os.path.join(ROOT, "gradslam_b200", "_lib", "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libgsx_%s.so" % name)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(flags) + ["-o", out] + sorted(glob.glob(os.path.join(SRC_DIR, "*.cu")))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed for variant " + name)
    return out


if __name__ == "__main__":
    wanted = dict(CANDIDATES)
    if len(sys.argv) > 1:
        wanted = {}
        for arg in sys.argv[1:]:
            name, _, flags = arg.partition("=")
            wanted[name] = [f for f in flags.split(",") if f]
    with ThreadPoolExecutor(max_workers=4) as pool:
        for path in pool.map(lambda kv: build(*kv), wanted.items()):
            print(path)
