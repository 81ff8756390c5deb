"""Per-kernel parity through the C ABI (include/dfd_b200_kernels.h) against fp32 PyTorch ops on the same
16-bit-rounded inputs.  Tolerances are one output rounding of the storage type plus accumulation noise:
fp16 2^-10, bf16 2^-7 relative (written next to each check)."""
import ctypes as C
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DT = {"fp16": (1, torch.float16, 2.0 ** -10), "bf16": (0, torch.bfloat16, 2.0 ** -7)}


@pytest.fixture(scope="module")
def lib():
    from deepfake_video_detection_b200 import _lib
    return _lib.load()


def chk(lib, rc):
    assert rc == 0, lib.dfd_last_error().decode()


def stream():
    return int(torch.cuda.current_stream().cuda_stream)


def close(out, ref, rel, abs_=None):
    out, ref = out.float(), ref.float()
    scale = ref.abs().max().item() + 1e-6
    err = (out - ref).abs().max().item()
    tol = rel * scale if abs_ is None else max(rel * scale, abs_)
    assert err <= tol, f"max abs err {err:.3e} > {tol:.3e} (scale {scale:.3e})"


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_preprocess_bit_exact(lib, prec):
    from oracle import effnet_b0_oracle as O
    code, tdt, _ = DT[prec]
    g = torch.Generator().manual_seed(1)
    u8 = torch.randint(0, 256, (5, 224, 224, 3), dtype=torch.uint8, generator=g)
    u8[0, 0, :, :] = torch.arange(224 * 3, dtype=torch.int64).remainder(256).to(torch.uint8).view(224, 3)   # every byte value, every channel
    u8[0, 1, :, :] = torch.arange(224 * 3, dtype=torch.int64).add(85).remainder(256).to(torch.uint8).view(224, 3)
    u8[0, 2, :, :] = torch.arange(224 * 3, dtype=torch.int64).add(170).remainder(256).to(torch.uint8).view(224, 3)
    out = torch.empty((5, 3, 224, 224), dtype=tdt, device="cuda")
    u8d = u8.cuda()
    chk(lib, lib.dfd_preprocess_u8hwc_to_nchw(u8d.data_ptr(), out.data_ptr(), 5, 224, 224, code, stream()))
    ref = O.prep_u8_hwc(u8).to(tdt)            # reference fp32 arithmetic, one rounding to the storage type
    assert torch.equal(out.cpu().view(torch.int16), ref.view(torch.int16))


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_stem(lib, prec, kind):
    from oracle import effnet_b0_oracle as O
    code, tdt, rel = DT[prec]
    g = torch.Generator().manual_seed(2)
    u8 = torch.randint(0, 256, (3, 64, 96, 3), dtype=torch.uint8, generator=g)
    x32 = O.prep_u8_hwc(u8)
    w = torch.randn(32, 3, 3, 3, generator=g) * 0.3
    b = torch.randn(32, generator=g) * 0.2
    wp = w.permute(2, 3, 1, 0).reshape(27, 32).contiguous().cuda()
    if kind == 0:
        xin, xref = u8.cuda(), x32
    elif kind == 1:
        xin, xref = x32.contiguous().cuda(), x32
    else:
        xin = x32.contiguous().to(tdt).cuda(); xref = xin.float().cpu()
    out = torch.empty((3, 32, 48, 32), dtype=tdt, device="cuda")
    bd = b.cuda()
    chk(lib, lib.dfd_k_stem(xin.data_ptr(), kind, wp.data_ptr(), bd.data_ptr(), out.data_ptr(), 3, 64, 96, code, stream()))
    ref = F.silu(F.conv2d(xref, w, b, 2, 1)).permute(0, 2, 3, 1)
    close(out.cpu(), ref, rel)


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("frames,H,W", [(3, 64, 96), (2, 224, 224), (1, 32, 32), (1, 32, 288), (2, 34, 48)])   # 288: wider than one row tile; 34: odd OH (one row per tile)
def test_stem_tcgen05(lib, prec, frames, H, W):
    """uint8 stem as tcgen05 implicit GEMM with hi/lo split operands: must agree with the fp32 conv to ~1 output rounding."""
    from oracle import effnet_b0_oracle as O
    code, tdt, rel = DT[prec]
    g = torch.Generator().manual_seed(7)
    u8 = torch.randint(0, 256, (frames, H, W, 3), dtype=torch.uint8, generator=g)
    w = torch.randn(32, 3, 3, 3, generator=g) * 0.3
    b = torch.randn(32, generator=g) * 0.2
    wp = w.permute(2, 3, 1, 0).reshape(27, 32).contiguous()               # host fp32 [(ky*3+kx)*3+c][oc]
    out = torch.full((frames, H // 2, W // 2, 32), float("nan"), dtype=tdt, device="cuda")
    u8d, bd = u8.cuda(), b.cuda()
    chk(lib, lib.dfd_k_stem_tc(u8d.data_ptr(), wp.data_ptr(), bd.data_ptr(), out.data_ptr(), frames, H, W, code, stream()))
    ref = F.silu(F.conv2d(O.prep_u8_hwc(u8), w, b, 2, 1)).permute(0, 2, 3, 1)
    close(out.cpu(), ref, rel)
    if prec == "fp16":      # the split keeps ~fp32 accuracy: error is one fp16 output rounding, not an input/weight rounding
        assert (out.cpu().float() - ref.to(tdt).float()).abs().max().item() <= 2 * 2.0 ** -10 * ref.abs().max().item() * 0.51


DW_SHAPES = [(32, 3, 1, 112), (96, 3, 2, 112), (144, 3, 1, 56), (144, 5, 2, 56), (240, 5, 1, 28), (240, 3, 2, 28),
             (480, 3, 1, 14), (480, 5, 1, 14), (672, 5, 1, 14), (672, 5, 2, 14), (1152, 5, 1, 7), (1152, 3, 1, 7)]


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("C_,k,s,H", DW_SHAPES)
def test_dwconv_and_squeeze(lib, prec, C_, k, s, H):
    code, tdt, rel = DT[prec]
    g = torch.Generator().manual_seed(C_ + k + s)
    frames = 3
    x = torch.randn(frames, H, H, C_, generator=g).to(tdt)
    w = torch.randn(C_, 1, k, k, generator=g) * (1.0 / k)
    b = torch.randn(C_, generator=g) * 0.2
    OH = (H + 2 * (k // 2) - k) // s + 1
    nparts = lib.dfd_k_dw_num_partials(OH, OH, C_, k, s)
    wp = w.reshape(C_, k * k).t().contiguous().cuda()
    out = torch.empty((frames, OH, OH, C_), dtype=tdt, device="cuda")
    parts = torch.full((frames, nparts, C_), float("nan"), device="cuda")
    xd, bd = x.cuda(), b.cuda()
    chk(lib, lib.dfd_k_dwconv(xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), out.data_ptr(), parts.data_ptr(),
                              frames, H, H, C_, k, s, code, stream()))
    ref = F.silu(F.conv2d(x.float().permute(0, 3, 1, 2), w, b, s, k // 2, 1, C_))
    close(out.cpu(), ref.permute(0, 2, 3, 1), rel)
    sums = parts.sum(1).cpu()
    assert torch.isfinite(sums).all()
    close(sums, ref.sum((2, 3)), 1e-5, abs_=1e-3)


@pytest.mark.parametrize("C_,rd,nparts,frames", [(32, 8, 49, 3), (96, 4, 37, 9), (1152, 48, 8, 17), (672, 28, 19, 8), (240, 10, 7, 5),
                                                  (480, 20, 2, 33), (672, 28, 1, 16), (1152, 48, 2, 40), (1152, 48, 1, 1)])
def test_se_gate(lib, C_, rd, nparts, frames):
    g = torch.Generator().manual_seed(C_)
    parts = torch.randn(frames, nparts, C_, generator=g)
    w1 = torch.randn(rd, C_, generator=g) * 0.1; b1 = torch.randn(rd, generator=g) * 0.1
    w2 = torch.randn(C_, rd, generator=g) * 0.3; b2 = torch.randn(C_, generator=g) * 0.3
    gate = torch.empty((frames, C_), device="cuda")
    inv = 1.0 / 123.0
    dev = [t.cuda() for t in (parts, w1, b1, w2.t().contiguous(), b2)]
    chk(lib, lib.dfd_k_se(dev[0].data_ptr(), nparts, C.c_float(inv), dev[1].data_ptr(), dev[2].data_ptr(),
                          dev[3].data_ptr(), dev[4].data_ptr(), gate.data_ptr(), frames, C_, rd, stream()))
    mean = parts.sum(1) * inv
    ref = torch.sigmoid(F.linear(F.silu(F.linear(mean, w1, b1)), w2, b2))
    close(gate.cpu(), ref, 1e-5, abs_=2e-6)


# (K, N, HW): the 19 distinct pointwise shapes of SURVEY.md App. A
PW_SHAPES = [(32, 16, 12544), (16, 96, 12544), (96, 24, 3136), (24, 144, 3136), (144, 24, 3136), (144, 40, 784),
             (40, 240, 784), (240, 40, 784), (240, 80, 196), (80, 480, 196), (480, 80, 196), (480, 112, 196),
             (112, 672, 196), (672, 112, 196), (672, 192, 49), (192, 1152, 49), (1152, 192, 49), (1152, 320, 49)]


def _gemm_case(lib, prec, K, N, HW, frames, gate, res, act, impl, seed=0):
    code, tdt, rel = DT[prec]
    g = torch.Generator().manual_seed(seed + K * 7 + N)
    M = frames * HW
    A = (torch.randn(M, K, generator=g)).to(tdt)
    Wt = (torch.randn(N, K, generator=g) * (1.0 / K ** 0.5)).to(tdt)
    bias = torch.randn(N, generator=g) * 0.3
    G = torch.rand(frames, K, generator=g) if gate else None
    R = torch.randn(M, N, generator=g).to(tdt) if res else None
    D = torch.full((M, N), float("nan"), dtype=tdt, device="cuda")
    Ad, Wd, bd = A.cuda(), Wt.cuda(), bias.cuda()
    Gd = G.cuda() if gate else None
    Rd = R.cuda() if res else None
    chk(lib, lib.dfd_k_gemm(Ad.data_ptr(), Wd.data_ptr(), bd.data_ptr(), Gd.data_ptr() if gate else None,
                            Rd.data_ptr() if res else None, D.data_ptr(), M, K, N, HW, int(act), code, impl, stream()))
    torch.cuda.synchronize()
    a = A.float()
    if gate:
        a = (a.view(frames, HW, K) * G.view(frames, 1, K)).to(tdt).float().view(M, K)   # operand is re-rounded
    ref = a.double() @ Wt.double().t() + bias.double()
    if act:
        ref = F.silu(ref)
    if res:
        ref = ref + R.double()
    close(D.cpu(), ref.float(), rel)


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("K,N,HW", PW_SHAPES)
def test_gemm_tcgen05_all_shapes(lib, prec, K, N, HW):
    frames = 2 if HW >= 3136 else 5          # M not a multiple of 128 for the small maps
    is_project = K > N or (K, N) == (32, 16)
    _gemm_case(lib, prec, K, N, HW, frames, gate=is_project, res=False, act=not is_project, impl=0)


@pytest.mark.parametrize("K,N,HW", [(144, 24, 3136), (240, 40, 784), (480, 80, 196), (672, 112, 196), (1152, 192, 49)])
def test_gemm_tcgen05_gate_residual(lib, K, N, HW):
    _gemm_case(lib, "fp16", K, N, HW, 3, gate=True, res=True, act=False, impl=0)


@pytest.mark.parametrize("K,N,HW,res", [(32, 16, 12544, False), (96, 24, 3136, False), (144, 24, 3136, True), (144, 40, 784, False), (240, 40, 784, True)])
def test_gemm_tcgen05_per_frame_weights(lib, K, N, HW, res):
    """Project convs of the big maps: SE gate folded into per-frame weights, frame-aligned tiles (partial last tile)."""
    code, tdt, rel = DT["fp16"]
    g = torch.Generator().manual_seed(K + N)
    frames = 3
    M = frames * HW
    A = torch.randn(M, K, generator=g).to(tdt); Wt = (torch.randn(N, K, generator=g) / K ** 0.5).to(tdt)
    bias = torch.randn(N, generator=g) * 0.3; G = torch.rand(frames, K, generator=g)
    R = torch.randn(M, N, generator=g).to(tdt) if res else None
    D = torch.full((M, N), float("nan"), dtype=tdt, device="cuda")
    dev = [t.cuda() for t in (A, Wt, bias, G)] + ([R.cuda()] if res else [])
    chk(lib, lib.dfd_k_gemm(dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), dev[3].data_ptr(), dev[4].data_ptr() if res else None,
                            D.data_ptr(), M, K, N, HW, 0, code, 2, stream()))
    wf = (Wt.float().view(1, N, K) * G.view(frames, 1, K)).to(tdt).double()                  # weights are re-rounded, not activations
    ref = torch.einsum("fmk,fnk->fmn", A.double().view(frames, HW, K), wf).reshape(M, N) + bias.double()
    if res:
        ref = ref + R.double()
    close(D.cpu(), ref.float(), rel)


@pytest.mark.parametrize("M_frames,HW", [(1, 49), (1, 1), (3, 127), (2, 129)])
def test_gemm_tcgen05_ragged_m(lib, M_frames, HW):
    _gemm_case(lib, "fp16", 96, 24, HW, M_frames, gate=True, res=False, act=False, impl=0)
    _gemm_case(lib, "fp16", 24, 144, HW, M_frames, gate=False, res=False, act=True, impl=0)


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("frames,HW", [(1, 49), (2, 49), (7, 49), (301, 49), (9, 64), (5, 36), (6, 100)])
def test_head_gemm_pool(lib, prec, frames, HW, impl=0):
    """conv_head + BN + SiLU + average pool: the transposed tcgen05 kernel (csrc/head_pool_tc.cu; maps of up to 64 pixels; 301 frames =
    more work units than SMs, a partial last tile) and the row-major fallback (100 pixels)."""
    code, tdt, rel = DT[prec]
    g = torch.Generator().manual_seed(frames)
    K, N = 320, 1280
    A = torch.randn(frames * HW, K, generator=g).to(tdt)
    Wt = (torch.randn(N, K, generator=g) * (1.0 / K ** 0.5)).to(tdt)
    bias = torch.randn(N, generator=g) * 0.3
    feat = torch.full((frames, N), float("nan"), device="cuda")
    Ad, Wd, bd = A.cuda(), Wt.cuda(), bias.cuda()
    chk(lib, lib.dfd_k_gemm_pool(Ad.data_ptr(), Wd.data_ptr(), bd.data_ptr(), feat.data_ptr(),
                                 frames * HW, K, N, HW, code, impl, stream()))
    ref = F.silu(A.double() @ Wt.double().t() + bias.double()).view(frames, HW, N).mean(1)
    close(feat.cpu(), ref.float(), 1e-5, abs_=1e-5)


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("frames,H,W,K,C,N", [(2, 7, 7, 256, 128, 64), (3, 14, 14, 64, 64, 64), (2, 56, 56, 64, 64, 64),
                                              (1, 28, 28, 512, 128, 128), (5, 7, 7, 2048, 512, 512), (1, 9, 5, 64, 64, 8)])
def test_conv1x1_conv3x3_implicit(lib, prec, frames, H, W, K, C, N):
    """resnet50 bottleneck conv1 -> conv2 through the zero-haloed map (implicit 3x3 GEMM): against F.conv2d on the same
    16-bit-rounded operands, with the intermediate rounded to the storage type as the kernel stores it."""
    code, tdt, rel = DT[prec]
    g = torch.Generator().manual_seed(H * 1000 + C)
    x = torch.randn(frames, H, W, K, generator=g).to(tdt)
    w1 = (torch.randn(C, K, generator=g) / K ** 0.5).to(tdt); b1 = torch.randn(C, generator=g) * 0.2
    w2 = (torch.randn(N, 3, 3, C, generator=g) / (9 * C) ** 0.5 * 2).to(tdt); b2 = torch.randn(N, generator=g) * 0.2
    rows = lib.dfd_k_conv3x3_maps(frames, H, W, C // 64, None, None, None, None)
    pad = torch.full((rows * C,), float("nan"), dtype=tdt, device="cuda")          # the entry point zeroes it itself
    out = torch.full((frames * H * W, N), float("nan"), dtype=tdt, device="cuda")
    dev = [t.cuda() for t in (x, w1, b1, w2.reshape(N, 9 * C).contiguous(), b2)]
    chk(lib, lib.dfd_k_conv1x1_conv3x3(dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), dev[3].data_ptr(), dev[4].data_ptr(),
                                       out.data_ptr(), frames, H, W, K, C, N, code, pad.data_ptr(), pad.numel() * 2, stream()))
    mid = torch.relu(x.double().reshape(-1, K) @ w1.double().t() + b1.double()).to(tdt)          # stored 16-bit
    mid = mid.double().view(frames, H, W, C).permute(0, 3, 1, 2)
    ref = torch.relu(F.conv2d(mid, w2.double().permute(0, 3, 1, 2), b2.double(), padding=1)).permute(0, 2, 3, 1).reshape(-1, N)
    assert torch.isfinite(out).all()
    close(out.cpu(), ref.float(), 2 * rel)
    halo = pad.view(rows, C).float().cpu()
    interior = torch.zeros(rows, dtype=torch.bool)
    import numpy as np
    pr = np.zeros(frames * H * W, np.int64)
    lib.dfd_k_conv3x3_maps(frames, H, W, C // 64, pr.ctypes.data, None, None, None)
    interior[torch.from_numpy(pr)] = True
    assert (halo[~interior] == 0).all()                                             # halo and guards untouched by the scatter


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("cin,mid,H,k,s", [(16, 96, 112, 3, 2), (24, 144, 56, 3, 1), (24, 144, 56, 5, 2)])
def test_mbconv_fused_expand_depthwise(lib, prec, cin, mid, H, k, s):
    """Expand 1x1 + SiLU fused into the marching depthwise kernel (mbconv_fused.cu): against fp32 PyTorch with the expanded
    tensor rounded to the storage type (the rounding point the kernel keeps), and against the two verified kernels it replaces."""
    code, tdt, rel = DT[prec]
    assert lib.dfd_k_mbconv_fused_supported(H, H, cin, mid, k, s) == 1 and lib.dfd_k_mbconv_fused_supported(H, H, cin, mid, k, 3 - s) == 0
    g = torch.Generator().manual_seed(cin * 7 + k + s)
    frames = 3
    x = torch.randn(frames, H, H, cin, generator=g).to(tdt)
    we = (torch.randn(mid, cin, generator=g) / cin ** 0.5).to(tdt); be = torch.randn(mid, generator=g) * 0.3
    w = torch.randn(mid, 1, k, k, generator=g) * (1.0 / k); b = torch.randn(mid, generator=g) * 0.2
    OH = (H + 2 * (k // 2) - k) // s + 1
    nparts = lib.dfd_k_dw_num_partials(OH, OH, mid, k, s)
    wp = w.reshape(mid, k * k).t().contiguous().cuda()
    xd, wed, bed, bd = x.cuda(), we.cuda(), be.cuda(), b.cuda()
    out = torch.full((frames, OH, OH, mid), float("nan"), dtype=tdt, device="cuda")
    parts = torch.full((frames, nparts, mid), float("nan"), device="cuda")
    chk(lib, lib.dfd_k_mbconv_fused(xd.data_ptr(), wed.data_ptr(), bed.data_ptr(), wp.data_ptr(), bd.data_ptr(), out.data_ptr(),
                                    parts.data_ptr(), frames, H, H, cin, mid, k, s, code, stream()))
    e = F.silu(x.double().reshape(-1, cin) @ we.double().t() + be.double()).to(tdt).float().view(frames, H, H, mid)
    ref = F.silu(F.conv2d(e.permute(0, 3, 1, 2), w, b, s, k // 2, 1, mid))
    assert torch.isfinite(out).all()
    close(out.cpu(), ref.permute(0, 2, 3, 1), 2 * rel)
    close(parts.sum(1).cpu(), ref.sum((2, 3)), 1e-3, abs_=5e-2)
    # the unfused pair: expand GEMM (tcgen05) -> marching depthwise
    E = torch.empty((frames * H * H, mid), dtype=tdt, device="cuda")
    chk(lib, lib.dfd_k_gemm(xd.data_ptr(), wed.data_ptr(), bed.data_ptr(), None, None, E.data_ptr(), frames * H * H, cin, mid, H * H, 1, code, 0, stream()))
    out2 = torch.empty_like(out); parts2 = torch.empty_like(parts)
    chk(lib, lib.dfd_k_dwconv(E.data_ptr(), wp.data_ptr(), bd.data_ptr(), out2.data_ptr(), parts2.data_ptr(), frames, H, H, mid, k, s, code, stream()))
    close(out.cpu(), out2.cpu(), rel)                                               # same rounding points; only the MMA accumulation order differs
    close(parts.cpu(), parts2.cpu(), 1e-3, abs_=5e-2)


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("M,K,N,act", [
    (591, 768, 768, 0),          # 3 row tiles of 256, the last one partial (TMA zero fill on load, clipped store)
    (100, 768, 256, 2),          # fewer rows than one CTA of the pair holds: the peer CTA's rows are all out of range
    (256 * 80 + 37, 768, 2304, 0),   # 729 tiles on 74 pairs: both accumulator buffers, every stage, the staging panels reused
    (4096, 3072, 768, 0),        # 48 k-blocks per tile (fc2 shape)
    (3000, 768, 3072, 2),        # fc1 shape with the GELU epilogue
    (1500, 72, 512, 2),          # K tail inside a 64-wide k-block
])
def test_gemm_cta_pairs(lib, prec, M, K, N, act):
    """The CTA-pair GEMM (csrc/gemm_pair.cu, `tcgen05.mma.cta_group::2`, TMA-store epilogue) against an fp64 matmul of the same
    16-bit-rounded operands; GELU is the exact erf form (timm's nn.GELU, reference src/models.py:93)."""
    code, tdt, rel = DT[prec]
    g = torch.Generator().manual_seed(M + K + N)
    A = (torch.randn(M, K, generator=g)).to(tdt)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(tdt)
    b = torch.randn(N, generator=g) * 0.5
    Ad, Wd, bd = A.cuda(), W.cuda(), b.cuda()
    D = torch.full((M, N), float("nan"), dtype=tdt, device="cuda")
    guard = torch.full((4096,), 7.0, dtype=tdt, device="cuda")          # allocated right behind D on a fresh allocator segment or not: checked either way
    chk(lib, lib.dfd_k_gemm(Ad.data_ptr(), Wd.data_ptr(), bd.data_ptr(), None, None, D.data_ptr(), M, K, N, 1, act, code, 3, stream()))
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t() + b.double()
    if act == 2:
        ref = F.gelu(ref)
    assert torch.isfinite(D).all()
    close(D.cpu(), ref, rel)
    assert (guard == 7.0).all()


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("M,K,N", [(591, 768, 768), (256 * 80 + 37, 768, 768), (4096, 3072, 768), (100, 72, 256)])
def test_gemm_cta_pairs_fp32_residual(lib, prec, M, K, N):
    """X += A W^T + bias on the fp32 residual stream, in place (the ViT attention-projection / fc2 epilogue of csrc/gemm_pair.cu):
    the accumulator tile goes through shared memory so that the read-modify-write of X is coalesced."""
    code, tdt, _ = DT[prec]
    g = torch.Generator().manual_seed(M + K + N + 1)
    A = torch.randn(M, K, generator=g).to(tdt)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(tdt)
    b = torch.randn(N, generator=g) * 0.5
    X0 = torch.randn(M + 1, N, generator=g) * 3.0                                    # one guard row behind the matrix
    Ad, Wd, bd, X = A.cuda(), W.cuda(), b.cuda(), X0.cuda()
    chk(lib, lib.dfd_k_gemm(Ad.data_ptr(), Wd.data_ptr(), bd.data_ptr(), None, X.data_ptr(), X.data_ptr(), M, K, N, 1, 0, code, 3, stream()))
    torch.cuda.synchronize()
    ref = X0[:M].double() + A.double() @ W.double().t() + b.double()
    err = (X[:M].cpu().double() - ref).abs().max().item()
    assert err <= 2e-5 * max(1.0, ref.abs().max().item()), err                       # fp32 accumulation and adds only
    assert torch.equal(X[M].cpu(), X0[M])


@pytest.mark.parametrize("C_,k,HW_,N", [(672, 5, 14, 112), (240, 3, 28, 40)])
def test_dependent_launches_without_synchronisation(lib, C_, k, HW_, N):
    """Programmatic dependent launch (csrc/common.cuh): every kernel is launched with the programmatic-stream-serialisation
    attribute and waits (`griddepcontrol.wait`) before it touches what the kernel ahead of it produced.  One MBConv tail —
    depthwise + SE sums -> SE gate -> gated project GEMM, each consuming the output of the launch before it — enqueued back to back WITHOUT any host synchronisation, 30 times over the same buffers (so a missing
    wait shows up as a write racing with the previous round's reads, too), must give the bits of the fully synchronised sequence."""
    code, tdt, _ = DT["fp16"]
    g = torch.Generator().manual_seed(C_ + N)
    frames, rd = 24, max(8, C_ // 24)
    x = torch.randn(frames, HW_, HW_, C_, generator=g).to(tdt).cuda()
    w = (torch.randn(C_, 1, k, k, generator=g) / k).reshape(C_, k * k).t().contiguous().cuda()
    b = (torch.randn(C_, generator=g) * 0.2).cuda()
    w1 = (torch.randn(rd, C_, generator=g) * 0.1).cuda(); b1 = (torch.randn(rd, generator=g) * 0.1).cuda()
    w2t = (torch.randn(rd, C_, generator=g) * 0.3).cuda(); b2 = (torch.randn(C_, generator=g) * 0.3).cuda()
    Wp = (torch.randn(N, C_, generator=g) / C_ ** 0.5).to(tdt).cuda(); bp = (torch.randn(N, generator=g) * 0.3).cuda()
    nparts = lib.dfd_k_dw_num_partials(HW_, HW_, C_, k, 1)
    torch.cuda.synchronize()

    def tail(sync):
        out = torch.empty((frames, HW_, HW_, C_), dtype=tdt, device="cuda")
        parts = torch.empty((frames, nparts, C_), device="cuda")
        gate = torch.empty((frames, C_), device="cuda")
        D = torch.empty((frames * HW_ * HW_, N), dtype=tdt, device="cuda")
        torch.cuda.synchronize()
        rounds = 1 if sync else 30
        for _ in range(rounds):
            chk(lib, lib.dfd_k_dwconv(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), parts.data_ptr(), frames, HW_, HW_, C_, k, 1, code, stream()))
            if sync: torch.cuda.synchronize()
            chk(lib, lib.dfd_k_se(parts.data_ptr(), nparts, C.c_float(1.0 / (HW_ * HW_)), w1.data_ptr(), b1.data_ptr(), w2t.data_ptr(), b2.data_ptr(),
                                  gate.data_ptr(), frames, C_, rd, stream()))
            if sync: torch.cuda.synchronize()
            chk(lib, lib.dfd_k_gemm(out.data_ptr(), Wp.data_ptr(), bp.data_ptr(), gate.data_ptr(), None, D.data_ptr(), frames * HW_ * HW_, C_, N, HW_ * HW_,
                                    0, code, 0, stream()))
            if sync: torch.cuda.synchronize()
        torch.cuda.synchronize()
        return out, gate, D

    ref = tail(True)
    got = tail(False)
    for a_, b_ in zip(ref, got):
        assert torch.isfinite(a_.float()).all() and torch.equal(a_, b_)
