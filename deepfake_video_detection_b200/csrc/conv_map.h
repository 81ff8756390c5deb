// Row arithmetic of the zero-haloed map used by the implicit 3x3 convolution (gemm_tc.cu, CONV variants).  Host and device
// code share these functions, so the CPU test of the scheme (tests/test_host.py, through dfd_k_conv3x3_maps) exercises
// exactly what the kernels compute.
//
// Layout: [guard: W+3 rows][frames x (H+2) x (W+2) padded pixels][guard: W+3 rows], C channels per row.  Interior pixel
// (f, y, x) is padded pixel (f, y+1, x+1).  For the tile of padded pixels starting at logical row m0, filter tap (ky, kx)
// reads physical rows m0 + ky*(W+2) + kx ... : the guard makes every tap row non-negative and the whole box in bounds, and
// a tap of an INTERIOR pixel never leaves its own frame's padded block (guard contents only reach dropped halo outputs).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define DFD_HD __host__ __device__ __forceinline__
#else
#define DFD_HD inline
#endif

namespace dfd {

// physical row (guard included) of interior pixel number m = (f*H + y)*W + x
DFD_HD int64_t conv_pad_row(uint32_t m, uint32_t H, uint32_t W) {
    const uint32_t hw = H * W, f = m / hw, rem = m - f * hw, y = rem / W, x = rem - y * W, pw = W + 2u;
    return (int64_t)(pw + 1u) + (int64_t)f * ((H + 2u) * pw) + (int64_t)((y + 1u) * pw + x + 1u);
}

// logical padded pixel p = (f*(H+2) + py)*(W+2) + px -> output row (f*H + py-1)*W + px-1; false for halo pixels
DFD_HD bool conv_unpad_row(uint32_t p, uint32_t H, uint32_t W, int64_t* out_row) {
    const uint32_t pw = W + 2u, phw = (H + 2u) * pw, f = p / phw, rem = p - f * phw, py = rem / pw, px = rem - py * pw;
    *out_row = (int64_t)f * (H * W) + (int64_t)((py - 1u) * W + (px - 1u));
    return py >= 1u && py <= H && px >= 1u && px <= W;
}

// k-block walk of one tile: kb = tap * cpk + ck; the A box of a k-block is (column ck*64, physical row `row`)
struct ConvTapIter {
    int ck, kx, row;
    DFD_HD void init(int m0) { ck = 0; kx = 0; row = m0; }
    DFD_HD void next(int cpk, int W) {
        if (++ck == cpk) {                       // next tap: one pixel to the right, or the start of the next padded row
            ck = 0;
            if (++kx == 3) { kx = 0; row += W; } else ++row;
        }
    }
};

}  // namespace dfd
