"""Build and run the CPU emulation of mbconv_fused_kernel: extracts the kernel text between the DFD_FUSED_KERNEL markers of
csrc/mbconv_fused.cu (unchanged) and compiles it with tools/host_emul/emul_mbconv_fused.cpp.
    python tools/host_emul/run.py [quick] [tsan]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = open(os.path.join(ROOT, "deepfake_video_detection_b200", "csrc", "mbconv_fused.cu")).read()
kernel = src[src.index("// DFD_FUSED_KERNEL_BEGIN"):src.index("// DFD_FUSED_KERNEL_END")]
out = os.path.join(ROOT, "build", "host_emul")
os.makedirs(out, exist_ok=True)
open(os.path.join(out, "mbconv_fused_kernel.inc"), "w").write(kernel)
tsan = "tsan" in sys.argv
exe = os.path.join(out, "emul_mbconv_fused" + ("_tsan" if tsan else ""))
cmd = ["g++", "-std=c++20", "-O1", "-g", "-pthread", "-I", out, os.path.join(ROOT, "tools", "host_emul", "emul_mbconv_fused.cpp"), "-o", exe]
if tsan:
    cmd += ["-fsanitize=thread"]
subprocess.check_call(cmd)
sys.exit(subprocess.call([exe] + (["quick"] if "quick" in sys.argv else [])))
