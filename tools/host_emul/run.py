"""Build and run the CPU emulations of the CUDA-core kernels: the kernel text between the emulation markers of the .cu file
is extracted UNCHANGED and compiled with the matching harness in this directory.
    python tools/host_emul/run.py [fused|se|poolhead|prepstem|march|resnet|rnn|all] [quick] [tsan]"""
import hashlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "build", "host_emul")
os.makedirs(OUT, exist_ok=True)
CASES = {"fused": ("mbconv_fused.cu", "DFD_FUSED_KERNEL", "mbconv_fused_kernel.inc", "emul_mbconv_fused.cpp"),
         "se": ("se.cu", "DFD_SE1_KERNEL", "se_kernel_v1.inc", "emul_se.cpp"),
         "poolhead": ("poolhead.cu", "DFD_POOLHEAD_KERNEL", "pool_head_kernel.inc", "emul_pool_head.cpp"),
         "march": ("dwconv_march.cu", "DFD_MARCH_KERNEL", "dwconv_march_kernel.inc", "emul_dwconv_march.cpp"),
         "resnet": ("resnet.cu", "DFD_RESNET_SMALL_KERNELS", "resnet_small_kernels.inc", "emul_resnet_small.cpp"),
         "rnn": ("rnn.cu", "DFD_RNN_KERNELS", "rnn_kernels.inc", "emul_rnn.cpp"),
         "prepstem": ("preprocess.cu", "DFD_PREP_KERNEL", "preprocess_kernel.inc", "emul_prep_stem.cpp")}
which = [a for a in sys.argv[1:] if a in CASES] or (list(CASES) if "all" in sys.argv else ["fused"])
tsan = "tsan" in sys.argv
rc = 0
for name in which:
    cu, marker, inc, cpp = CASES[name]
    if not os.path.exists(os.path.join(ROOT, "tools", "host_emul", cpp)):
        continue
    src = open(os.path.join(ROOT, "deepfake_video_detection_b200", "csrc", cu)).read()
    text = src[src.index(f"// {marker}_BEGIN"):src.index(f"// {marker}_END")]
    if name == "march":                                # the kernel's one inline-asm statement (a 32-bit shared-memory load)
        asm_line = 'asm volatile("ld.shared.b32 %0, [%1];" : "=r"(raw) : "r"(sb_c + jj * pix_b));'
        assert text.count(asm_line) == 1
        text = text.replace(asm_line, "raw = lds32(sb_c + jj * pix_b);")
    if name == "rnn":
        assert text.count("extern __shared__ float sm[];") == 1
        text = text.replace("extern __shared__ float sm[];", "")
    if name == "poolhead":                             # the one dynamic shared-memory array becomes a host buffer
        assert "extern __shared__ float s_f[];" in text
        text = text.replace("extern __shared__ float s_f[];", "float* s_f = g_dyn_smem;")
    open(os.path.join(OUT, inc), "w").write(text)
    if name == "prepstem":
        st = open(os.path.join(ROOT, "deepfake_video_detection_b200", "csrc", "stem.cu")).read()
        open(os.path.join(OUT, "stem_kernel.inc"), "w").write(st[st.index("// DFD_STEM_KERNEL_BEGIN"):st.index("// DFD_STEM_KERNEL_END")])
    quick = "quick" in sys.argv or ("stem" in sys.argv and name == "fused")
    flags = ["-std=c++20", "-O1", "-pthread"] + (["-DEMUL_QUICK"] if quick and name == "fused" else []) + (["-g", "-fsanitize=thread"] if tsan else [])
    cpp_path = os.path.join(ROOT, "tools", "host_emul", cpp)
    incs = "".join(open(os.path.join(OUT, f)).read() for f in sorted(os.listdir(OUT)) if f.endswith(".inc"))
    key = hashlib.sha1((open(cpp_path).read() + incs + " ".join(flags)).encode()).hexdigest()[:12]
    exe = os.path.join(OUT, f"{cpp[:-4]}_{key}")
    if not os.path.exists(exe):                        # same sources + flags: reuse the binary of an earlier invocation
        subprocess.check_call(["g++", *flags, "-I", OUT, cpp_path, "-o", exe])
    rc |= subprocess.call([exe] + (["quick"] if "quick" in sys.argv else []) + (["stem"] if "stem" in sys.argv and name == "fused" else []))
sys.exit(rc)
