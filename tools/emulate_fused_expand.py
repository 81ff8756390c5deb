"""Design check (CPU, numpy) of the shared-memory addressing of mbconv_fused.cu's expand step: the formulas below are
transcribed from `expand_row` and from the x-row / weight stagers; the mma.sync m16n8k16 fragment semantics are those the
GPU-verified attention kernel (vit.cu) relies on.  For every (warp, lane, tile) the emulation gathers the A / B fragments
from byte-addressed shared-memory images, applies the MMA by definition and scatters the result with the kernel's store
addresses; the expanded slot must equal silu(x @ We^T + be) at (pixel + PAD) * CB + channel and stay zero elsewhere.
Run: python tools/emulate_fused_expand.py"""
import numpy as np


def check(KS, S, CIN, C, W, CB):
    PAD, TW = KS // 2, 7
    OW = (W + 2 * PAD - KS) // S + 1
    strips = OW // TW
    THREADS = (strips * (CB // 2) + 31) // 32 * 32; WARPS = THREADS // 32
    pixw = max((strips * TW - 1) * S + KS, W + 2 * PAD)
    KP = (CIN + 15) & ~15; KSTEPS = KP // 16; XP = KP + 8
    PXT = (W + 15) // 16; NTL = CB // 8
    rsb, xsb, wsb = pixw * CB * 2, PXT * 16 * XP * 2, CB * XP * 2
    rng = np.random.default_rng(W + CIN)
    x = rng.standard_normal((W, CIN)).astype(np.float16)
    we = (rng.standard_normal((CB, CIN)) / CIN ** 0.5).astype(np.float16)
    be = rng.standard_normal(CB).astype(np.float32)
    # shared-memory images in halves (address / 2)
    xs = np.zeros(xsb // 2, np.float16); ws = np.zeros(wsb // 2, np.float16); es = np.zeros(rsb // 2, np.float16)
    XCH = W * (CIN // 8)
    for i in range(XCH):                                   # x stager: chunk i -> (px, sub)
        px, sub = divmod(i, CIN // 8)
        s_off = (px * XP + sub * 8) * 2
        xs[s_off // 2: s_off // 2 + 8] = x[px, sub * 8: sub * 8 + 8]
    for i in range(CB * (CIN // 8)):                       # weight stager
        r, q = divmod(i, CIN // 8)
        off = (r * XP + q * 8) * 2
        ws[off // 2: off // 2 + 8] = we[r, q * 8: q * 8 + 8]
    ld = lambda img, addr: img[addr // 2: addr // 2 + 2].astype(np.float32)     # one 32-bit word = 2 halves
    written = np.zeros(rsb // 2, bool)
    for warp in range(WARPS):
        for tile in range(warp, PXT * NTL, WARPS):
            pt, nt = divmod(tile, NTL)
            A = np.zeros((16, KP), np.float32); B = np.zeros((KP, 8), np.float32)
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                ar = ((pt * 16 + g) * XP + 2 * t) * 2
                br = ((nt * 8 + g) * XP + 2 * t) * 2
                for ks in range(KSTEPS):
                    k0 = ks * 16 + 2 * t
                    A[g, k0:k0 + 2] = ld(xs, ar + ks * 32); A[g + 8, k0:k0 + 2] = ld(xs, ar + ks * 32 + 8 * XP * 2)
                    A[g, k0 + 8:k0 + 10] = ld(xs, ar + ks * 32 + 16); A[g + 8, k0 + 8:k0 + 10] = ld(xs, ar + ks * 32 + 8 * XP * 2 + 16)
                    B[k0:k0 + 2, g] = ld(ws, br + ks * 32); B[k0 + 8:k0 + 10, g] = ld(ws, br + ks * 32 + 16)
            Cm = A @ B                                      # [16 px][8 ch]
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                px0, px1 = pt * 16 + g, pt * 16 + g + 8
                ea = ((px0 + PAD) * CB + nt * 8 + 2 * t) * 2
                bb = be[nt * 8 + 2 * t: nt * 8 + 2 * t + 2]
                silu = lambda v: v / (1 + np.exp(-v))
                if px0 < W:
                    es[ea // 2: ea // 2 + 2] = silu(Cm[g, 2 * t:2 * t + 2] + bb); written[ea // 2: ea // 2 + 2] = True
                if px1 < W:
                    e1 = ea + 8 * CB * 2
                    es[e1 // 2: e1 // 2 + 2] = silu(Cm[g + 8, 2 * t:2 * t + 2] + bb); written[e1 // 2: e1 // 2 + 2] = True
    ref = x.astype(np.float32) @ we.astype(np.float32).T + be
    ref = ref / (1 + np.exp(-ref))
    slot = es.reshape(pixw, CB).astype(np.float32)
    err = np.abs(slot[PAD:PAD + W] - ref).max()
    assert err < 2e-2, err
    wr = written.reshape(pixw, CB)
    assert wr[PAD:PAD + W].all() and not wr[:PAD].any() and not wr[PAD + W:].any()
    # the depthwise phase reads columns ox0*S + jj, jj < (TW-1)*S + KS, of the slot: all inside it
    assert (strips * TW - 1) * S + KS <= pixw
    print(f"k{KS} s{S} cin{CIN} mid{C} W{W} CB{CB}: {WARPS} warps, {PXT * NTL} tiles, slot {pixw}x{CB}, max err {err:.2e} ok")


for spec in [(3, 2, 16, 96, 112, 48), (3, 1, 24, 144, 56, 48), (5, 2, 24, 144, 56, 48),
             (3, 2, 16, 96, 112, 96), (3, 1, 24, 144, 56, 72), (5, 2, 24, 144, 56, 144),       # DFD_FUSE_CB=1 alternatives
             (3, 2, 40, 240, 28, 48), (3, 1, 80, 480, 14, 96), (5, 2, 112, 672, 14, 96), (3, 1, 192, 1152, 7, 128)]:   # level 2
    check(*spec)
