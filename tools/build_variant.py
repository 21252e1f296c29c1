"""Kernel-tuning helper: link libctensor_b200 with ONE source recompiled under extra -D flags.
usage: build_variant.py NAME source.cu -DX=1 -DY=2 ...   ->  lol_b200/csrc/build/variants/NAME.so
Run a script against it with LOLB_LIBRARY=<that path>."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lol_b200 import build
name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
build.build_library()
vdir = os.path.join(build.OBJ_DIR, "variants"); os.makedirs(vdir, exist_ok=True)
obj = os.path.join(vdir, name + ".o")
subprocess.run(["nvcc", *[f for f in build.NVCC_FLAGS if f not in ("-Xptxas", "-v")], *flags, "-c", os.path.join(build.CSRC, src), "-o", obj], check=True)
objs = [os.path.join(build.OBJ_DIR, f) for f in os.listdir(build.OBJ_DIR) if f.endswith(".o") and f != src[:-3] + ".o"] + [obj]
out = os.path.join(vdir, name + ".so")
subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs], check=True)
print(out)
