"""Build a VARIANT of libdfd_b200.so with extra -D flags on selected sources (kernel tuning sweeps on the GPU box).

    python tools/build_variant.py <name> <source.cu>[,<source.cu>...] -DFOO=1 [-DBAR=2 ...]

Objects of the other sources are taken from build/obj (run the normal build first).  Output: build/variants/libdfd_<name>.so,
selected at run time with DFD_LIB_PATH (deepfake_video_detection_b200/_lib.py) — never by a switch inside the product code."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepfake_video_detection_b200 import _build as B

name, srcs, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
out_dir = os.path.join(ROOT, "build", "variants")
os.makedirs(out_dir, exist_ok=True)
B.build()
objs = []
for s in B.SOURCES:
    obj = os.path.join(B.OBJ, s[:-3] + ".o")
    if s in srcs:
        obj = os.path.join(out_dir, f"{s[:-3]}_{name}.o")
        subprocess.check_call([B._nvcc(), *B.NVCC_FLAGS, *flags, "-c", os.path.join(B.CSRC, s), "-o", obj])
    objs.append(obj)
lib = os.path.join(out_dir, f"libdfd_{name}.so")
subprocess.check_call([B._nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
print(lib)
