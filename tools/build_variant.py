"""Tuning builds: python tools/build_variant.py NAME [-DX=Y ...]  ->  5g-nr-randomaccess_b200/tune/NAME.so
(git-ignored; travels to the GPU box; select with RACH_GPU_LIB=... or tools/ab*.py)."""
import importlib, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
b = importlib.import_module("5g-nr-randomaccess_b200.build")
name, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(b.HERE, "tune", name + ".so")
os.makedirs(os.path.dirname(out), exist_ok=True)
subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + extra + b.SOURCES + ["-o", out])
print(out)
