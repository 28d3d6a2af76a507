"""tools/build_variant.py NAME [-DFLAG=V ...] -- builds an experiment variant of the library into
scratch/libk4lz4_NAME.so (never the shipped in-tree .so); tools/dbench.py --lib loads it."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from k4os.compression.lz4_b200 import build as b
name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "scratch", f"libk4lz4_{name}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
cmd = ["/usr/local/cuda/bin/nvcc"] + b.NVCC_FLAGS + flags + ["-o", out, os.path.join(b.CSRC, "k4lz4_api.cu")]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode != 0:
    sys.exit("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
print(out)
