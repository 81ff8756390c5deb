"""CPU emulation of the CUDA-core kernels (tools/host_emul/): the UNCHANGED kernel text of the .cu files is compiled with
g++ against host stand-ins for the device helpers — one std::thread per CUDA thread, std::barrier for __syncthreads, mma.sync /
ldmatrix / shfl by their PTX definitions, lazy and eager cp.async models, NaN-filled shared memory — and checked against a
plain fp32 / double reference.  This pins the kernels' indexing and barrier placement on the CPU; the parity tests proper
run on the GPU (tests/test_gpu_kernels.py)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20)")


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "host_emul", "run.py"), *args], capture_output=True, text=True, timeout=900)
    if r.returncode != 0 and ("Resource temporarily unavailable" in r.stderr or "std::system_error" in r.stderr):
        pytest.skip("this machine does not allow the ~1000 threads a CTA emulation needs")
    assert r.returncode == 0 and "MISMATCH" not in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_fused_expand_depthwise_kernel_source_on_cpu():
    out = _run("fused", "quick")                     # two small-map geometries of the template + one bf16 case, both cp.async models
    assert out.count("-> ok") == 6


def test_se_gate_kernel_source_on_cpu():
    assert _run("se").count("-> ok") == 8               # squeeze-excite gate: 8 shapes incl. a partial last CTA and an odd rd


def test_pool_head_kernel_source_on_cpu():
    """Rows a6 / a7 (attention pool + head, pretrained_detector.py:123-141): the default kernel's source against the reference
    arithmetic in double, ragged videos incl. an empty one, both pooling modes, both feature widths (1280: efficientnet_b0,
    2048: the resnet50 member, row a9)."""
    assert _run("poolhead").count("-> ok") == 4


def test_preprocess_and_stem_kernel_sources_on_cpu():
    """Row a1 (K1 tensor prep: bit-exact with the reference's fp32 arithmetic rounded once) and row a3 (CUDA-core stem for the fp32
    NCHW input forward() receives, and for uint8 crops with the prep fused): the default kernels' sources on CPU threads."""
    out = _run("prepstem")
    assert out.count("-> ok") == 3 and "0 of 3072 values differ" in out


def test_depthwise_march_kernel_source_on_cpu():
    """Row a4, the dominant kernel of the step (depthwise kxk + BN + SiLU + squeeze-excite sums): the default kernel's source on CPU
    threads for five of the network's own shapes (compile-time geometry) and two run-time-geometry shapes, both cp.async models."""
    assert _run("march").count("-> ok") == 14


def test_resnet_member_small_kernel_sources_on_cpu():
    """Row a9 (`resnet50` ensemble member): its max-pool and average-pool kernels on CPU threads (its attention pool + head is
    the shared kernel of poolhead.cu, emulated above)."""
    assert _run("resnet").count("-> ok") == 2


def test_rnn_head_kernel_sources_on_cpu():
    """Row a10 (LogicRNNLSTM head, config 3): LogicCell gate math with the length mask, attention over time + classifier."""
    assert _run("rnn").count("-> ok") == 2
