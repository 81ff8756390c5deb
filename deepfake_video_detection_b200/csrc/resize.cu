// K0 — crop + resize of face boxes to SxS uint8 crops, the step in front of the scoring path:
// reference `pil.crop((x1, y1, x2, y2)).resize((face_size, face_size))` at app.py:1964-1978 and
// src/data_prepare.py:54-56.  Pillow's `Image.resize` default for RGB is BICUBIC; the arithmetic (Pillow is an
// un-vendored dependency, requirements.txt:17 `pillow>=11.0.0`) is restated from its published algorithm
// (src/libImaging/Resample.c) and reproduced BIT-EXACTLY:
//   * separable: horizontal pass, 8-bit rounded + clipped intermediate, vertical pass;
//   * antialiased windows: support = 2 * max(in/out, 1), [int(c - support + 0.5), int(c + support + 0.5)) clipped to the
//     crop, c = (xx + 0.5) * in/out; bicubic kernel a = -0.5; weights normalised in double, converted to 22-bit fixed
//     point (round half away from zero); int32 accumulation from 1 << 21, >> 22, clip to [0, 255].
// Coefficient tables are built on the host in double precision with Pillow's operation order (one table per distinct
// input extent), the two passes run on the GPU straight from the full video frames (no intermediate crop copy).
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/dfd_b200.h"
#include "../../include/dfd_b200_kernels.h"

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

struct CropDev {                 // per-crop record in the workspace
    long long src_off;           // byte offset of pixel (x1, y1) of the frame inside d_frames
    int pitch;                   // bytes per frame row
    int cw, ch;                  // crop extent
    int hb_off, hk_off, hks;     // horizontal table: bounds[S][2], coefficients[S][hks] (int32 offsets into the table area)
    int vb_off, vk_off, vks;
    long long tmp_off;           // byte offset of the [ch][S][3] intermediate inside the tmp area
};

double bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for one axis; appends bounds then coefficients to `tab`
void build_coeffs(int in_size, int out_size, std::vector<int>& tab, int& b_off, int& k_off, int& ksize) {
    double scale = (double)in_size / out_size, filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    ksize = (int)std::ceil(support) * 2 + 1;
    b_off = (int)tab.size();
    tab.resize(tab.size() + (size_t)out_size * 2);
    k_off = (int)tab.size();
    tab.resize(tab.size() + (size_t)out_size * ksize, 0);
    const double ss = 1.0 / filterscale;
    std::vector<double> k(ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) { k[x] = bicubic((x + xmin - center + 0.5) * ss); ww += k[x]; }
        for (int x = 0; x < xmax; ++x) {
            const double v = ww != 0.0 ? k[x] / ww : k[x];
            tab[(size_t)k_off + (size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << kPrecisionBits)) : (int)(0.5 + v * (1 << kPrecisionBits));
        }
        tab[(size_t)b_off + 2 * xx] = xmin; tab[(size_t)b_off + 2 * xx + 1] = xmax;
    }
}

__device__ __forceinline__ unsigned char clip8(int acc) {
    const int v = acc >> kPrecisionBits;                     // arithmetic shift, as Pillow's clip8 lookup
    return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: grid (crop, crop row), threads over (xx, channel)
__global__ void resize_h_kernel(const unsigned char* __restrict__ frames, const CropDev* __restrict__ crops,
                                const int* __restrict__ tab, unsigned char* __restrict__ tmp, int S) {
    const CropDev c = crops[blockIdx.x];
    const int yy = blockIdx.y;
    if (yy >= c.ch) return;
    const unsigned char* row = frames + c.src_off + (long long)yy * c.pitch;
    unsigned char* dst = tmp + c.tmp_off + (long long)yy * S * 3;
    for (int i = threadIdx.x; i < S * 3; i += blockDim.x) {
        const int xx = i / 3, ch = i - xx * 3;
        const int x0 = tab[c.hb_off + 2 * xx], n = tab[c.hb_off + 2 * xx + 1];
        const int* k = tab + c.hk_off + xx * c.hks;
        int acc = 1 << (kPrecisionBits - 1);
        for (int x = 0; x < n; ++x) acc += (int)row[(x0 + x) * 3 + ch] * k[x];
        dst[i] = clip8(acc);
    }
}

// vertical pass: grid (crop, output row), threads over (xx, channel)
__global__ void resize_v_kernel(const CropDev* __restrict__ crops, const int* __restrict__ tab,
                                const unsigned char* __restrict__ tmp, unsigned char* __restrict__ out, int S) {
    const CropDev c = crops[blockIdx.x];
    const int yy = blockIdx.y;
    const int y0 = tab[c.vb_off + 2 * yy], n = tab[c.vb_off + 2 * yy + 1];
    const int* k = tab + c.vk_off + yy * c.vks;
    const unsigned char* src = tmp + c.tmp_off + (long long)y0 * S * 3;
    unsigned char* dst = out + ((long long)blockIdx.x * S + yy) * S * 3;
    for (int i = threadIdx.x; i < S * 3; i += blockDim.x) {
        int acc = 1 << (kPrecisionBits - 1);
        for (int y = 0; y < n; ++y) acc += (int)src[(long long)y * S * 3 + i] * k[y];
        dst[i] = clip8(acc);
    }
}

thread_local std::string g_rs_err;
int rsfail(int code, const std::string& m) { g_rs_err = m; return code; }
size_t up256(size_t b) { return (b + 255) & ~size_t(255); }

struct Layout { size_t crops_bytes, tab_bytes, tmp_bytes; };

// validates the boxes and builds the host-side records + tables (tables shared between crops / axes of equal extent)
int plan(const dfd_crop_box* boxes, int64_t n, int S, std::vector<CropDev>* recs, std::vector<int>* tab, Layout& L) {
    if (!boxes || n <= 0 || S <= 0 || S > 1024) return rsfail(DFD_EINVAL, "dfd_crop_resize: bad argument");
    std::map<int, int> seen;                                   // extent -> index into `meta`
    struct Meta { int b_off, k_off, ks; };
    std::vector<Meta> meta;
    size_t tmp = 0, ints = 0;
    auto table = [&](int extent) {
        auto it = seen.find(extent);
        if (it != seen.end()) return meta[it->second];
        Meta m{};
        if (tab) build_coeffs(extent, S, *tab, m.b_off, m.k_off, m.ks);
        else { double fs = (double)extent / S; if (fs < 1.0) fs = 1.0; m.ks = (int)std::ceil(2.0 * fs) * 2 + 1; }
        ints += (size_t)S * (2 + m.ks);
        seen[extent] = (int)meta.size(); meta.push_back(m);
        return m;
    };
    for (int64_t i = 0; i < n; ++i) {
        const dfd_crop_box& b = boxes[i];
        if (b.frame_w <= 0 || b.frame_h <= 0 || b.x1 < 0 || b.y1 < 0 || b.x2 > b.frame_w || b.y2 > b.frame_h || b.x2 <= b.x1 || b.y2 <= b.y1 || b.frame_offset < 0)
            return rsfail(DFD_EINVAL, "dfd_crop_resize: box " + std::to_string(i) + " is empty or outside its frame (clamp as app.py:1968-1973 does)");
        const Meta h = table(b.x2 - b.x1), v = table(b.y2 - b.y1);
        if (recs) {
            CropDev c{};
            c.src_off = b.frame_offset + ((long long)b.y1 * b.frame_w + b.x1) * 3;
            c.pitch = b.frame_w * 3; c.cw = b.x2 - b.x1; c.ch = b.y2 - b.y1;
            c.hb_off = h.b_off; c.hk_off = h.k_off; c.hks = h.ks; c.vb_off = v.b_off; c.vk_off = v.k_off; c.vks = v.ks;
            c.tmp_off = (long long)tmp;
            recs->push_back(c);
        }
        tmp += up256((size_t)(b.y2 - b.y1) * S * 3);
    }
    L.crops_bytes = up256((size_t)n * sizeof(CropDev)); L.tab_bytes = up256(ints * 4); L.tmp_bytes = tmp;
    return DFD_OK;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

const char* dfd_resize_last_error(void) { return g_rs_err.c_str(); }

int dfd_crop_resize_workspace_bytes(const dfd_crop_box* h_boxes, int64_t n, int out_size, size_t* bytes) {
    if (!bytes) return rsfail(DFD_EINVAL, "dfd_crop_resize_workspace_bytes: null pointer");
    Layout L{};
    const int rc = plan(h_boxes, n, out_size, nullptr, nullptr, L);
    if (rc) return rc;
    *bytes = L.crops_bytes + L.tab_bytes + L.tmp_bytes + 512;
    return DFD_OK;
}

int dfd_crop_resize_u8(const uint8_t* d_frames, const dfd_crop_box* h_boxes, int64_t n, int out_size, uint8_t* d_out,
                       void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!d_frames || !d_out || !d_workspace) return rsfail(DFD_EINVAL, "dfd_crop_resize_u8: null pointer");
    std::vector<CropDev> recs; std::vector<int> tab; Layout L{};
    const int rc = plan(h_boxes, n, out_size, &recs, &tab, L);
    if (rc) return rc;
    if (workspace_bytes < L.crops_bytes + L.tab_bytes + L.tmp_bytes + 512) return rsfail(DFD_ENOMEM, "dfd_crop_resize_u8: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(d_workspace) + 255) & ~uintptr_t(255));
    CropDev* d_crops = reinterpret_cast<CropDev*>(ws);
    int* d_tab = reinterpret_cast<int*>(ws + L.crops_bytes);
    uint8_t* d_tmp = ws + L.crops_bytes + L.tab_bytes;
    // pageable host memory: cudaMemcpyAsync stages the bytes before returning, so the vectors may die with this call
    cudaError_t e = cudaMemcpyAsync(d_crops, recs.data(), recs.size() * sizeof(CropDev), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return rsfail(DFD_ECUDA, std::string("dfd_crop_resize_u8: table upload: ") + cudaGetErrorString(e));
    int max_h = 0;
    for (const CropDev& c : recs) max_h = c.ch > max_h ? c.ch : max_h;
    for (int64_t i0 = 0; i0 < n; i0 += 32768) {                      // grid.x = crops (chunks keep grid.y * grid.x sane)
        const unsigned nb = (unsigned)((n - i0) < 32768 ? (n - i0) : 32768);
        resize_h_kernel<<<dim3(nb, (unsigned)max_h), 256, 0, s>>>(d_frames, d_crops + i0, d_tab, d_tmp, out_size);
        resize_v_kernel<<<dim3(nb, (unsigned)out_size), 256, 0, s>>>(d_crops + i0, d_tab, d_tmp, d_out + (size_t)i0 * out_size * out_size * 3, out_size);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return rsfail(DFD_ECUDA, std::string("dfd_crop_resize_u8: launch: ") + cudaGetErrorString(e));
    return DFD_OK;
}

int dfd_k_resize_coeffs(int in_size, int out_size, int32_t* h_bounds, int32_t* h_coeffs, int* ksize) {
    if (in_size <= 0 || out_size <= 0 || !ksize) return rsfail(DFD_EINVAL, "dfd_k_resize_coeffs: bad argument");
    std::vector<int> tab; int b_off = 0, k_off = 0, ks = 0;
    build_coeffs(in_size, out_size, tab, b_off, k_off, ks);
    *ksize = ks;
    if (h_bounds) memcpy(h_bounds, tab.data() + b_off, (size_t)out_size * 2 * sizeof(int));
    if (h_coeffs) memcpy(h_coeffs, tab.data() + k_off, (size_t)out_size * ks * sizeof(int));
    return ks;
}

#pragma GCC visibility pop
}  // extern "C"
