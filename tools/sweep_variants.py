"""Builds variants of libhge_b200.so (compile-time knobs) here, or runs them on the GPU box.

    python tools/sweep_variants.py build     # in the dev container (nvcc, no GPU)
    python tools/sweep_variants.py run       # on the GPU box (gpurun)
"""
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VAR_DIR = os.path.join(ROOT, "hypergraphembedding_b200", "_variants")

VARIANTS = {
    "acc%d_mb%d" % (acc, mb): ["HGE_ACCUM_MODE=%d" % acc, "HGE_MIN_BLOCKS=%d" % mb]
    for acc, mb in itertools.product((0, 1, 2), (2, 3, 4))
}
TUNINGS = [(64, 256, 0), (32, 256, 0), (128, 256, 0), (64, 128, 0), (64, 512, 0), (16, 128, 0)]


def build():
  from hypergraphembedding_b200 import build as b
  os.makedirs(VAR_DIR, exist_ok=True)
  for name, defines in VARIANTS.items():
    out = os.path.join(VAR_DIR, "libhge_%s.so" % name)
    b.build(force=True, defines=defines, out=out, tag="_" + name)
    print("built", out)


def run():
  results = []
  for name in sorted(VARIANTS):
    env = dict(os.environ, HGE_LIB_PATH=os.path.join(VAR_DIR, "libhge_%s.so" % name))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "time_variant.py"), "c2"],
                       env=env, capture_output=True, text=True)
    line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-400:]
    print(name, line, flush=True)
    results.append((name, line))
  best = "acc1_mb3"
  for tuning in TUNINGS:
    env = dict(os.environ, HGE_LIB_PATH=os.path.join(VAR_DIR, "libhge_%s.so" % best))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "time_variant.py"), "c2"] +
                       [str(v) for v in tuning], env=env, capture_output=True, text=True)
    line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-400:]
    print(best, tuning, line, flush=True)


if __name__ == "__main__":
  {"build": build, "run": run}[sys.argv[1]]()
