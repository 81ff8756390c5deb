"""Design check (CPU, numpy) of vit_attention_v2_kernel's fragment addressing (vit.cu, DFD_VIT_ATTN_V2): ldmatrix is
emulated from its PTX definition (lane l supplies the address of row l%8 of matrix l/8; after the load, lane i holds the
32-bit word (row i/4, columns 2(i%4), 2(i%4)+1) of each matrix, or with .trans the elements (2(i%4), i/4), (2(i%4)+1, i/4));
mma.m16n8k16 from the fragment layout the GPU-verified first attention kernel relies on.  The address arithmetic is
transcribed from the kernel.  One warp, every q tile, one head: must equal softmax(Q K^T / 8) V.
Run: python tools/emulate_attention_v2.py"""
import numpy as np

TOK, TOKPAD, HD, STRIDE = 197, 208, 64, 72
ROWB = STRIDE * 2


def ldsm_x4(img, lane_addr, trans):
    """img: fp16 array indexed in halves; lane_addr[32] byte addresses -> regs[32][4][2] (two halves per register)."""
    out = np.zeros((32, 4, 2), np.float32)
    for m in range(4):
        rows = np.stack([img[lane_addr[m * 8 + r] // 2: lane_addr[m * 8 + r] // 2 + 8] for r in range(8)]).astype(np.float32)   # 8x8
        for i in range(32):
            g, t = i >> 2, i & 3
            out[i, m] = (rows[2 * t, g], rows[2 * t + 1, g]) if trans else (rows[g, 2 * t], rows[g, 2 * t + 1])
    return out


def mma(c, a, b0, b1):
    """c[32][4] += A(16x16) B(16x8); a[32][4][2], b0/b1[32][2] per the m16n8k16 fragment layout."""
    A = np.zeros((16, 16), np.float32); B = np.zeros((16, 8), np.float32)
    for i in range(32):
        g, t = i >> 2, i & 3
        A[g, 2 * t:2 * t + 2] = a[i, 0]; A[g + 8, 2 * t:2 * t + 2] = a[i, 1]
        A[g, 2 * t + 8:2 * t + 10] = a[i, 2]; A[g + 8, 2 * t + 8:2 * t + 10] = a[i, 3]
        B[2 * t:2 * t + 2, g] = b0[i]; B[2 * t + 8:2 * t + 10, g] = b1[i]
    D = A @ B
    for i in range(32):
        g, t = i >> 2, i & 3
        c[i] += (D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1])


rng = np.random.default_rng(0)
Q, K, V = (rng.standard_normal((TOK, HD)).astype(np.float16) for _ in range(3))
img = np.zeros(3 * TOKPAD * STRIDE, np.float16)                       # [sQ | sK | sV], padded rows zero
for which, M in enumerate((Q, K, V)):
    for tok in range(TOK):
        img[which * TOKPAD * STRIDE + tok * STRIDE: which * TOKPAD * STRIDE + tok * STRIDE + HD] = M[tok]
sq, sk, sv = 0, TOKPAD * STRIDE * 2, 2 * TOKPAD * STRIDE * 2          # byte offsets
lanes = np.arange(32)
a_off = (lanes & 15) * ROWB + (lanes >> 4) * 16
k_off = (lanes & 7) * ROWB + (lanes >> 3) * 16
v_off = ((lanes & 7) + ((lanes >> 3) & 1) * 8) * ROWB + (lanes >> 4) * 16
scale = 0.125 * 1.4426950408889634
out = np.zeros((TOKPAD, HD), np.float32)
for qt in range(TOKPAD // 16):
    q0 = qt * 16
    aq = [ldsm_x4(img, sq + q0 * ROWB + a_off + ks * 32, False) for ks in range(4)]
    m = np.full((32, 2), -np.inf, np.float32); l = np.zeros((32, 2), np.float32)
    acc = np.zeros((8, 32, 4), np.float32)
    for kb0 in range(0, TOKPAD, 64):
        nkt = min(64, TOKPAD - kb0) // 8
        s = np.zeros((8, 32, 4), np.float32)
        for nt in range(nkt):
            ka = sk + (kb0 + nt * 8) * ROWB + k_off
            b0, b1 = ldsm_x4(img, ka, False), ldsm_x4(img, ka + 64, False)
            mma(s[nt], aq[0], b0[:, 0], b0[:, 1]); mma(s[nt], aq[1], b0[:, 2], b0[:, 3])
            mma(s[nt], aq[2], b1[:, 0], b1[:, 1]); mma(s[nt], aq[3], b1[:, 2], b1[:, 3])
        for nt in range(8):
            for i in range(32):
                t = i & 3
                for j in range(2):
                    ok = nt < nkt and kb0 + nt * 8 + 2 * t + j < TOK
                    s[nt, i, j] = s[nt, i, j] * scale if ok else -np.inf
                    s[nt, i, 2 + j] = s[nt, i, 2 + j] * scale if ok else -np.inf
        bm = np.stack([s[:, :, :2].max(axis=(0, 2)), s[:, :, 2:].max(axis=(0, 2))], 1)          # per lane
        for i in range(32):                                                                      # quad reduction (xor 1, 2)
            quad = [i & ~3 | q for q in range(4)]
            bm[i] = bm[quad].max(0) if False else bm[i]
        bmq = np.stack([bm[(i & ~3):(i & ~3) + 4].max(0) for i in range(32)])
        n = np.maximum(m, bmq)
        with np.errstate(invalid="ignore"):
            al = np.exp2(m - n)
        m = n
        p = np.zeros_like(s)
        p[:, :, :2] = np.exp2(s[:, :, :2] - n[None, :, 0:1]); p[:, :, 2:] = np.exp2(s[:, :, 2:] - n[None, :, 1:2])
        l = l * al + np.stack([p[:, :, :2].sum(axis=(0, 2)), p[:, :, 2:].sum(axis=(0, 2))], 1)
        acc[:, :, :2] *= al[None, :, 0:1]; acc[:, :, 2:] *= al[None, :, 1:2]
        p16 = p.astype(np.float16).astype(np.float32)
        for kk in range(4):
            if 2 * kk < nkt:
                ap = np.zeros((32, 4, 2), np.float32)
                ap[:, 0] = p16[2 * kk][:, 0:2]; ap[:, 1] = p16[2 * kk][:, 2:4]
                ap[:, 2] = p16[2 * kk + 1][:, 0:2]; ap[:, 3] = p16[2 * kk + 1][:, 2:4]
                va = sv + (kb0 + kk * 16) * ROWB + v_off
                for dp in range(4):
                    bv = ldsm_x4(img, va + dp * 32, True)
                    mma(acc[2 * dp], ap, bv[:, 0], bv[:, 1]); mma(acc[2 * dp + 1], ap, bv[:, 2], bv[:, 3])
    lq = np.stack([l[(i & ~3):(i & ~3) + 4].sum(0) for i in range(32)])
    for i in range(32):
        g, t = i >> 2, i & 3
        for dt in range(8):
            out[q0 + g, dt * 8 + 2 * t: dt * 8 + 2 * t + 2] = acc[dt, i, 0:2] / lq[i, 0]
            out[q0 + g + 8, dt * 8 + 2 * t: dt * 8 + 2 * t + 2] = acc[dt, i, 2:4] / lq[i, 1]
S = Q.astype(np.float32) @ K.astype(np.float32).T / 8
P = np.exp(S - S.max(1, keepdims=True)); P /= P.sum(1, keepdims=True)
ref = P @ V.astype(np.float32)
err = np.abs(out[:TOK] - ref).max()
print(f"attention v2 emulation: max abs err {err:.2e} (scale {np.abs(ref).max():.2f})")
assert err < 5e-3
