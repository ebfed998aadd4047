#include "imgops.cuh"

#include <math.h>

namespace locr {

namespace {

constexpr int kOutW = 100, kOutH = 32, kPrec = 22;

__device__ __forceinline__ double bicubic_filter(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for one output index (libImaging/Resample.c).
__device__ void pil_coeffs(int in_size, int out_size, int xx, int ksize, int32_t* k, int32_t* bounds) {
    const double scale = (double)in_size / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 2.0 * filterscale;
    const double ss = 1.0 / filterscale;
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) ww += bicubic_filter((x + xmin - center + 0.5) * ss);
    for (int x = 0; x < ksize; ++x) {
        double w = 0.0;
        if (x < xmax) {
            w = bicubic_filter((x + xmin - center + 0.5) * ss);
            if (ww != 0.0) w /= ww;
        }
        k[x] = w < 0 ? (int)(-0.5 + w * (1 << kPrec)) : (int)(0.5 + w * (1 << kPrec));
    }
    bounds[0] = xmin;
    bounds[1] = xmax;
}

__device__ __forceinline__ int clip8(int v) {
    v >>= kPrec;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__device__ __forceinline__ int gray_at(const CropDesc& d, int y, int x) {
    const uint8_t* p = d.src + (long)y * d.stride + (long)x * d.channels;
    if (d.channels == 1) return p[0];
    // cv2 BGR2GRAY, 8-bit: 15-bit fixed point
    return (int)(((unsigned)p[0] * 3735u + (unsigned)p[1] * 19235u + (unsigned)p[2] * 9798u + 16384u) >> 15);
}

__global__ void __launch_bounds__(256)
crop_resize_kernel(const CropDesc* __restrict__ descs, int32_t* __restrict__ coef_scratch,
                   uint8_t* __restrict__ inter_scratch, float* __restrict__ out_x, uint8_t* __restrict__ out_u8) {
    const CropDesc d = descs[blockIdx.x];
    const int t = threadIdx.x;
    float* ox = out_x + (long)blockIdx.x * kOutW * kOutH;
    uint8_t* ou = out_u8 ? out_u8 + (long)blockIdx.x * kOutW * kOutH : nullptr;
    if (d.h <= 0 || d.w <= 0) {
        for (int i = t; i < kOutW * kOutH; i += 256) {
            ox[i] = 0.f;
            if (ou) ou[i] = 0;
        }
        return;
    }
    int32_t* kh = coef_scratch + d.coef_off;            // [100][ksh]
    int32_t* kv = kh + kOutW * d.ksh;                    // [32][ksv]
    int32_t* bh = kv + kOutH * d.ksv;                    // [100][2]
    int32_t* bv = bh + kOutW * 2;                        // [32][2]
    uint8_t* inter = inter_scratch + d.inter_off;        // [h][100]
    const bool need_h = d.w != kOutW, need_v = d.h != kOutH;
    if (t < kOutW) {
        if (need_h) pil_coeffs(d.w, kOutW, t, d.ksh, kh + t * d.ksh, bh + t * 2);
    } else if (t < kOutW + kOutH) {
        const int yy = t - kOutW;
        if (need_v) pil_coeffs(d.h, kOutH, yy, d.ksv, kv + yy * d.ksv, bv + yy * 2);
    }
    __syncthreads();
    // horizontal pass (uint8 intermediate, Pillow ImagingResampleHorizontal_8bpc)
    for (int i = t; i < d.h * kOutW; i += 256) {
        const int y = i / kOutW, xx = i - y * kOutW;
        int v;
        if (need_h) {
            const int xmin = bh[xx * 2], xmax = bh[xx * 2 + 1];
            const int32_t* k = kh + xx * d.ksh;
            int ss0 = 1 << (kPrec - 1);
            for (int x = 0; x < xmax; ++x) ss0 += gray_at(d, y, x + xmin) * k[x];
            v = clip8(ss0);
        } else {
            v = gray_at(d, y, xx);
        }
        inter[i] = (uint8_t)v;
    }
    __syncthreads();
    // vertical pass + ToTensor + (x - 0.5) / 0.5
    for (int i = t; i < kOutH * kOutW; i += 256) {
        const int yy = i / kOutW, xx = i - yy * kOutW;
        int v;
        if (need_v) {
            const int ymin = bv[yy * 2], ymax = bv[yy * 2 + 1];
            const int32_t* k = kv + yy * d.ksv;
            int ss0 = 1 << (kPrec - 1);
            for (int y = 0; y < ymax; ++y) ss0 += (int)inter[(y + ymin) * kOutW + xx] * k[y];
            v = clip8(ss0);
        } else {
            v = inter[yy * kOutW + xx];
        }
        if (ou) ou[i] = (uint8_t)v;
        ox[i] = ((float)v / 255.0f - 0.5f) / 0.5f;
    }
}

// cv2.resize INTER_LINEAR, 8-bit, 3 channels: 11-bit fixed-point coefficients (imgproc/src/resize.cpp).
__device__ __forceinline__ void linear_tab(int d, double scale, int in_size, bool clamp, int& s, int& w0, int& w1) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int si = (int)floorf(f);
    f -= (float)si;
    if (clamp && si < 0) { si = 0; f = 0.f; }
    if (clamp && si >= in_size - 1) { si = in_size - 1; f = 0.f; }
    s = si;
    w0 = __float2int_rn((1.f - f) * 2048.f);
    w1 = __float2int_rn(f * 2048.f);
}

__global__ void __launch_bounds__(256)
resize_linear_kernel(const uint8_t* __restrict__ src, int B, int sh, int sw, uint8_t* __restrict__ dst, int dh,
                     int dw) {
    const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    const long total = (long)B * dh * dw;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int dx = (int)(i % dw), dy = (int)((i / dw) % dh), b = (int)(i / ((long)dw * dh));
        int sx, ax0, ax1, sy, by0, by1;
        linear_tab(dx, scale_x, sw, true, sx, ax0, ax1);
        linear_tab(dy, scale_y, sh, false, sy, by0, by1);
        const int x1 = sx + 1 < sw ? sx + 1 : sw - 1;
        int y0 = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
        int y1 = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
        const uint8_t* r0 = src + ((long)b * sh + y0) * sw * 3;
        const uint8_t* r1 = src + ((long)b * sh + y1) * sw * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int S0 = r0[sx * 3 + c] * ax0 + r0[x1 * 3 + c] * ax1;
            const int S1 = r1[sx * 3 + c] * ax0 + r1[x1 * 3 + c] * ax1;
            int v = (((by0 * (S0 >> 4)) >> 16) + ((by1 * (S1 >> 4)) >> 16) + 2) >> 2;
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
            dst[i * 3 + c] = (uint8_t)v;
        }
    }
}

}  // namespace

void crop_scratch_sizes(int h, int w, int* ksh, int* ksv, long* coef_ints, long* inter_bytes) {
    auto ksize = [](int in, int out) {
        double scale = (double)in / out;
        if (scale < 1.0) scale = 1.0;
        return (int)ceil(2.0 * scale) * 2 + 1;
    };
    *ksh = ksize(w > 0 ? w : 1, kOutW);
    *ksv = ksize(h > 0 ? h : 1, kOutH);
    *coef_ints = (long)kOutW * *ksh + (long)kOutH * *ksv + kOutW * 2 + kOutH * 2;
    *inter_bytes = ((long)(h > 0 ? h : 0) * kOutW + 15) / 16 * 16;
}

void launch_crop_resize(const CropDesc* d_descs, int n, int32_t* coef_scratch, uint8_t* inter_scratch, float* out_x,
                        uint8_t* out_u8, cudaStream_t s) {
    if (n <= 0) return;
    crop_resize_kernel<<<n, 256, 0, s>>>(d_descs, coef_scratch, inter_scratch, out_x, out_u8);
}

void launch_resize_linear_bgr(const uint8_t* src, int B, int sh, int sw, uint8_t* dst, int dh, int dw,
                              cudaStream_t s) {
    const long total = (long)B * dh * dw;
    long g = (total + 255) / 256;
    if (g > 148L * 16) g = 148L * 16;
    resize_linear_kernel<<<(int)g, 256, 0, s>>>(src, B, sh, sw, dst, dh, dw);
}

}  // namespace locr
