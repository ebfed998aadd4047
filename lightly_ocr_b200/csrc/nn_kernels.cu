#include "nn_kernels.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

namespace locr {

namespace {

__device__ __forceinline__ float act2f(uint16_t u, int f16) {
    if (f16) return __half2float(__ushort_as_half(u));
    return __bfloat162float(__ushort_as_bfloat16(u));
}
// saturating (F2FP.SATFINITE): beyond the fp16 range a value becomes +-65504, never an infinity
__device__ __forceinline__ uint16_t f2act(float v, int f16) {
    uint16_t r;
    if (f16) asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
    else asm("cvt.rn.satfinite.bf16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8], int f16) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = act2f((uint16_t)(w[i] & 0xffff), f16);
        f[2 * i + 1] = act2f((uint16_t)(w[i] >> 16), f16);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], int f16) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        w[i] = (uint32_t)f2act(f[2 * i], f16) | ((uint32_t)f2act(f[2 * i + 1], f16) << 16);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------------------------------- direct conv (tiny Cin)
// One thread = one output pixel x CPT output channels (CPT = 32): the 9*CIN inputs are loaded and converted once and
// reused for 32 accumulators.  In u8 mode the weights arrive pre-divided by the per-channel std, so the normalisation
// (x - mean)/std costs one subtraction per input (canvas padding reads as raw 0, i.e. -mean/std, conv padding as 0).
// POOL: the thread owns one pixel of the MaxPool2d(2, 2) output instead and reduces its four conv results in registers
// (conv + BN + ReLU + max-pool of TPS_STN.py:38-43 in one pass; the full-resolution tensor is never written).
template <int CIN, bool U8, bool POOL>
__global__ void __launch_bounds__(256)
direct_conv3x3_kernel(const void* __restrict__ in, int B, int H, int W, int img_h, int img_w, long row_stride,
                      long img_stride, const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                      uint16_t* __restrict__ out, long out_pitch, int relu, int f16, int split) {
    constexpr int CPT = 32;
    extern __shared__ float sw[];  // [9*CIN][Cout] then bias[Cout]
    const int nw = 9 * CIN * Cout;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[nw + i] = bias[i];
    __syncthreads();
    const uint32_t groups = (uint32_t)Cout / CPT;
    const int OH = POOL ? H / 2 : H, OW = POOL ? W / 2 : W;
    const uint32_t total = (uint32_t)B * OH * OW * groups;   // < 2^31, checked by the launcher
    const float mean[3] = {(float)(0.485 * 255.0), (float)(0.456 * 255.0), (float)(0.406 * 255.0)};
    for (uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += gridDim.x * blockDim.x) {
        const uint32_t cg = gid % groups;
        const uint32_t pix = gid / groups;
        const int ox = (int)(pix % (uint32_t)OW);
        const uint32_t rest = pix / (uint32_t)OW;
        const int oy = (int)(rest % (uint32_t)OH);
        const int b = (int)(rest / (uint32_t)OH);
        float res[CPT];
#pragma unroll 1
        for (int sub = 0; sub < (POOL ? 4 : 1); ++sub) {
            const int x = POOL ? 2 * ox + (sub & 1) : ox;
            const int y = POOL ? 2 * oy + (sub >> 1) : oy;
            float acc[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[j] = sw[nw + cg * CPT + j];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = y + ky - 1;
                if (yy < 0 || yy >= H) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = x + kx - 1;
                    if (xx < 0 || xx >= W) continue;
                    float v[CIN];
                    if (U8) {
                        const bool inside = (yy < img_h) && (xx < img_w);
                        const uint8_t* p =
                            reinterpret_cast<const uint8_t*>(in) + (long)b * img_stride + (long)yy * row_stride + xx * 3;
#pragma unroll
                        for (int c = 0; c < CIN; ++c) v[c] = (inside ? (float)p[c] : 0.0f) - mean[c];
                    } else {
                        v[0] = __ldg(reinterpret_cast<const float*>(in) + ((long)(b * H + yy) * W + xx));
                    }
                    const float* wt = sw + ((ky * 3 + kx) * CIN) * Cout + cg * CPT;
#pragma unroll
                    for (int c = 0; c < CIN; ++c) {
#pragma unroll
                        for (int j = 0; j < CPT; j += 4) {
                            const float4 w4 = *reinterpret_cast<const float4*>(wt + c * Cout + j);
                            acc[j] = fmaf(v[c], w4.x, acc[j]);
                            acc[j + 1] = fmaf(v[c], w4.y, acc[j + 1]);
                            acc[j + 2] = fmaf(v[c], w4.z, acc[j + 2]);
                            acc[j + 3] = fmaf(v[c], w4.w, acc[j + 3]);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < CPT; ++j) res[j] = sub == 0 ? acc[j] : fmaxf(res[j], acc[j]);
        }
        uint16_t* op = out + (long)pix * out_pitch + cg * CPT;
#pragma unroll
        for (int q = 0; q < CPT / 8; ++q) {
            float r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = relu ? fmaxf(res[q * 8 + j], 0.f) : res[q * 8 + j];
            const uint4 hi = pack8(r, f16);
            *reinterpret_cast<uint4*>(op + q * 8) = hi;
            if (split) {  // split precision: lo = v - hi goes to channel Cout + n
                float h[8];
                unpack8(hi, h, f16);
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] -= h[j];
                *reinterpret_cast<uint4*>(op + Cout + q * 8) = pack8(r, f16);
            }
        }
    }
}

// conv3x3 (Cin = 1, fp32 input) + BN + ReLU + MaxPool2d(2, 2) (TPS_STN.py:38-43), one thread = one POOLED pixel x 16 output
// channels: the 4 x 4 input patch is loaded once, and every 16-byte weight load from shared memory feeds the four
// positions of the pooling window (16 FMAs per LDS.128 instead of 4: the per-position form above is bound by the
// shared-memory pipe).  Taps outside the image multiply a zero, which leaves the fp32 sums exactly as the skipping form
// computes them.
// POOL = false: the same thread writes its 2 x 2 block of conv outputs instead of their maximum (conv0_1 of the ResNet).
template <bool POOL>
__global__ void __launch_bounds__(256)
direct_conv3x3_pool4_kernel(const float* __restrict__ in, int B, int H, int W, const float* __restrict__ w,
                            const float* __restrict__ bias, int Cout, uint16_t* __restrict__ out, long out_pitch,
                            int relu, int f16, int split, int out_row_px) {
    constexpr int CPT = 16;
    extern __shared__ float sw[];  // [9][Cout] then bias[Cout]
    const int nw = 9 * Cout;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[nw + i] = bias[i];
    __syncthreads();
    const uint32_t groups = (uint32_t)Cout / CPT;
    const int OH = H / 2, OW = W / 2;
    const uint32_t total = (uint32_t)B * OH * OW * groups;   // < 2^31, checked by the launcher
    for (uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += gridDim.x * blockDim.x) {
        const uint32_t cg = gid % groups;
        const uint32_t pix = gid / groups;
        const int ox = (int)(pix % (uint32_t)OW);
        const uint32_t rest = pix / (uint32_t)OW;
        const int oy = (int)(rest % (uint32_t)OH);
        const int b = (int)(rest / (uint32_t)OH);
        float patch[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int yy = 2 * oy - 1 + r;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int xx = 2 * ox - 1 + c;
                patch[r][c] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(in + ((long)(b * H + yy) * W + xx)) : 0.f;
            }
        }
        float acc[4][CPT];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4)
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[s4][j] = sw[nw + cg * CPT + j];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float* wt = sw + (ky * 3 + kx) * Cout + cg * CPT;
#pragma unroll
                for (int j = 0; j < CPT; j += 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wt + j);
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) {
                        const float v = patch[(s4 >> 1) + ky][(s4 & 1) + kx];
                        acc[s4][j] = fmaf(v, w4.x, acc[s4][j]);
                        acc[s4][j + 1] = fmaf(v, w4.y, acc[s4][j + 1]);
                        acc[s4][j + 2] = fmaf(v, w4.z, acc[s4][j + 2]);
                        acc[s4][j + 3] = fmaf(v, w4.w, acc[s4][j + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int s4 = 0; s4 < (POOL ? 1 : 4); ++s4) {
            // out_row_px > 0: row-padded output (one zero pixel left, two right of every row: conv_tc.cuh x_row_px)
            const long opix = POOL ? (long)pix
                                   : (out_row_px > 0 ? (long)(b * H + 2 * oy + (s4 >> 1)) * out_row_px + 2 * ox + (s4 & 1) + 1
                                                     : (long)(b * H + 2 * oy + (s4 >> 1)) * W + 2 * ox + (s4 & 1));
            uint16_t* op = out + opix * out_pitch + cg * CPT;
#pragma unroll
            for (int q = 0; q < CPT / 8; ++q) {
                float r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float m = POOL ? fmaxf(fmaxf(acc[0][q * 8 + j], acc[1][q * 8 + j]),
                                                 fmaxf(acc[2][q * 8 + j], acc[3][q * 8 + j]))
                                         : acc[s4][q * 8 + j];
                    r[j] = relu ? fmaxf(m, 0.f) : m;
                }
                const uint4 hi = pack8(r, f16);
                *reinterpret_cast<uint4*>(op + q * 8) = hi;
                if (split) {  // split precision: lo = v - hi goes to channel Cout + n
                    float h[8];
                    unpack8(hi, h, f16);
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] -= h[j];
                    *reinterpret_cast<uint4*>(op + Cout + q * 8) = pack8(r, f16);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------- CRAFT image -> NHWC16
// normalizeMeanVariance on the zero-padded canvas (reference imgproc.py:19-25, :58-60) written as a 16-channel NHWC
// tensor (channels 3..15 zero) so that the first convolution runs on the tensor cores like every other layer.
__global__ void __launch_bounds__(256)
preproc_nhwc16_kernel(const uint8_t* __restrict__ in, int B, int H, int W, int img_h, int img_w, long row_stride,
                      long img_stride, uint16_t* __restrict__ out, int f16) {
    const float mean[3] = {(float)(0.485 * 255.0), (float)(0.456 * 255.0), (float)(0.406 * 255.0)};
    const float stdv[3] = {(float)(0.229 * 255.0), (float)(0.224 * 255.0), (float)(0.225 * 255.0)};
    // output rows are padded to W + 3 pixels: column 0 and columns W + 1, W + 2 are zero (conv_tc.cuh: x_row_px)
    const uint32_t Wp = (uint32_t)W + 3;
    // two threads per pixel (16 bytes each), so that a warp's store is one contiguous 512-byte run
    const uint32_t total = (uint32_t)B * H * Wp * 2u;   // < 2^32, checked by the caller
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const uint32_t pix = idx >> 1;
        const bool upper = idx & 1u;                    // channels 8..15: always zero
        const int xp = (int)(pix % Wp);
        const uint32_t rest = pix / Wp;
        const int y = (int)(rest % (uint32_t)H);
        const int b = (int)(rest / (uint32_t)H);
        const int x = xp - 1;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (!upper && x >= 0 && x < W) {
            const bool inside = (y < img_h) && (x < img_w);
            const uint8_t* p = in + (long)b * img_stride + (long)y * row_stride + x * 3;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = ((inside ? (float)p[c] : 0.0f) - mean[c]) / stdv[c];
            val = pack8(v, f16);
        }
        reinterpret_cast<uint4*>(out)[idx] = val;
    }
}

// Same with 8 channels per pixel (3 used): one thread and one 16-byte store per pixel.
__global__ void __launch_bounds__(256)
preproc_nhwc8_kernel(const uint8_t* __restrict__ in, int B, int H, int W, int img_h, int img_w, long row_stride,
                     long img_stride, uint16_t* __restrict__ out, int f16) {
    const float mean[3] = {(float)(0.485 * 255.0), (float)(0.456 * 255.0), (float)(0.406 * 255.0)};
    const float stdv[3] = {(float)(0.229 * 255.0), (float)(0.224 * 255.0), (float)(0.225 * 255.0)};
    const uint32_t Wp = (uint32_t)W + 3;
    const uint32_t total = (uint32_t)B * H * Wp;
    for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += gridDim.x * blockDim.x) {
        const int xp = (int)(pix % Wp);
        const uint32_t rest = pix / Wp;
        const int y = (int)(rest % (uint32_t)H);
        const int b = (int)(rest / (uint32_t)H);
        const int x = xp - 1;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (x >= 0 && x < W) {
            const bool inside = (y < img_h) && (x < img_w);
            const uint8_t* p = in + (long)b * img_stride + (long)y * row_stride + x * 3;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = ((inside ? (float)p[c] : 0.0f) - mean[c]) / stdv[c];
            val = pack8(v, f16);
        }
        reinterpret_cast<uint4*>(out)[pix] = val;
    }
}

// ------------------------------------------------------------------------------------------- max pool
__device__ __forceinline__ uint32_t max2_packed(uint32_t a, uint32_t b, int f16) {
    if (f16) {
        const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
        return *reinterpret_cast<const uint32_t*>(&m);
    }
    const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&m);
}

__global__ void __launch_bounds__(256)
maxpool_kernel(const uint16_t* __restrict__ in, long in_pitch, int B, int H, int W, int C, uint16_t* __restrict__ out,
               long out_pitch, int OH, int OW, int kh, int kw, int sh, int sw_, int ph, int pw, int f16, int split) {
    // 32-bit index arithmetic (the launcher checks total < 2^31): 64-bit divisions would make this ALU-bound
    const uint32_t groups = (uint32_t)C >> 3;
    const uint32_t total = (uint32_t)B * OH * OW * groups;
    for (uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += gridDim.x * blockDim.x) {
        const uint32_t cg = gid % groups;
        const uint32_t pix = gid / groups;
        const int ox = (int)(pix % (uint32_t)OW);
        const uint32_t rest = pix / (uint32_t)OW;
        const int oy = (int)(rest % (uint32_t)OH);
        const int b = (int)(rest / (uint32_t)OH);
        if (!split) {
            // plain tensors: the maximum of 16-bit values is exact in their own format, four packed max per tap
            const uint32_t ninf = f16 ? 0xFC00FC00u : 0xFF80FF80u;
            uint4 mp = make_uint4(ninf, ninf, ninf, ninf);
            for (int ky = 0; ky < kh; ++ky) {
                const int iy = oy * sh + ky - ph;
                if (iy < 0 || iy >= H) continue;
                for (int kx = 0; kx < kw; ++kx) {
                    const int ix = ox * sw_ + kx - pw;
                    if (ix < 0 || ix >= W) continue;
                    const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + ((long)(b * H + iy) * W + ix) * in_pitch + cg * 8));
                    mp.x = max2_packed(mp.x, u.x, f16); mp.y = max2_packed(mp.y, u.y, f16);
                    mp.z = max2_packed(mp.z, u.z, f16); mp.w = max2_packed(mp.w, u.w, f16);
                }
            }
            *reinterpret_cast<uint4*>(out + (long)pix * out_pitch + cg * 8) = mp;
            continue;
        }
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
        for (int ky = 0; ky < kh; ++ky) {
            const int iy = oy * sh + ky - ph;
            if (iy < 0 || iy >= H) continue;
            for (int kx = 0; kx < kw; ++kx) {
                const int ix = ox * sw_ + kx - pw;
                if (ix < 0 || ix >= W) continue;
                const uint16_t* src = in + ((long)(b * H + iy) * W + ix) * in_pitch + cg * 8;
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(src));
                float f[8];
                unpack8(u, f, f16);
                if (split) {  // value = hi + lo (exact in fp32), lo lives C channels further
                    const uint4 ul = __ldg(reinterpret_cast<const uint4*>(src + C));
                    float l[8];
                    unpack8(ul, l, f16);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] += l[j];
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
            }
        }
        const uint4 hi = pack8(m, f16);
        uint16_t* dst = out + (long)pix * out_pitch + cg * 8;
        *reinterpret_cast<uint4*>(dst) = hi;
        if (split) {
            float h[8];
            unpack8(hi, h, f16);
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] -= h[j];
            *reinterpret_cast<uint4*>(dst + C) = pack8(m, f16);
        }
    }
}

// ------------------------------------------------------------------------------------------- bilinear 2x
// F.interpolate(scale 2, mode='bilinear', align_corners=False) (model.py:47,51,55).  One thread = 8 channels of one
// SOURCE pixel (k, l) and its 2 x 2 block of output pixels: output row 2k + dy mixes source rows k - 1 + dy and k + dy
// with weights 0.25 / 0.75 (dy = 0) or 0.75 / 0.25 (dy = 1), same for columns; rows / columns outside the tensor
// clamp to the edge, which reproduces PyTorch's clamping of negative source coordinates and of the last index.  Nine
// source loads feed four outputs (the plain per-output form needs sixteen) and the index arithmetic is shared.
__global__ void __launch_bounds__(256, 4)
upsample2x_kernel(const uint16_t* __restrict__ in, long in_pitch, int B, int H, int W, int C,
                  uint16_t* __restrict__ out, long out_pitch, int f16) {
    const uint32_t groups = (uint32_t)C >> 3;
    const uint32_t total = (uint32_t)B * H * W * groups;   // < 2^31, checked by the launcher
    const int OW = 2 * W;
    for (uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += gridDim.x * blockDim.x) {
        const uint32_t cg = gid % groups;
        const uint32_t pix = gid / groups;
        const int l = (int)(pix % (uint32_t)W);
        const uint32_t rest = pix / (uint32_t)W;
        const int k = (int)(rest % (uint32_t)H);
        const int b = (int)(rest / (uint32_t)H);
        const int ks[3] = {k > 0 ? k - 1 : 0, k, k < H - 1 ? k + 1 : H - 1};
        const int ls[3] = {l > 0 ? l - 1 : 0, l, l < W - 1 ? l + 1 : W - 1};
        const uint16_t* base = in + (long)b * H * W * in_pitch + cg * 8;
        // all nine loads first (memory-level parallelism), then row by row: horizontal lerps of a source row (output
        // columns 2l and 2l + 1), then the vertical lerp with the previous row - PyTorch's own order of operations
        // (value = hy * row(y0) + ly * row(y1) with row(y) = hx * p(x0) + lx * p(x1)); only two lerped rows stay live
        uint4 raw[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                raw[i][j] = __ldg(reinterpret_cast<const uint4*>(base + (long)(ks[i] * W + ls[j]) * in_pitch));
        uint16_t* obase = out + ((long)b * 2 * H * OW) * out_pitch + cg * 8;
        float hprev[2][8];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float a[8], m[8], c[8], hcur[2][8];
            unpack8(raw[i][0], a, f16);
            unpack8(raw[i][1], m, f16);
            unpack8(raw[i][2], c, f16);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                hcur[0][q] = 0.25f * a[q] + 0.75f * m[q];
                hcur[1][q] = 0.75f * m[q] + 0.25f * c[q];
            }
            if (i > 0) {
                const int dy = i - 1;
                const float ly = dy == 0 ? 0.75f : 0.25f, hy = 1.f - ly;
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    float r[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) r[q] = hy * hprev[dx][q] + ly * hcur[dx][q];
                    *reinterpret_cast<uint4*>(obase + ((long)(2 * k + dy) * OW + 2 * l + dx) * out_pitch) = pack8(r, f16);
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                hprev[0][q] = hcur[0][q];
                hprev[1][q] = hcur[1][q];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------- TPS localisation head
// AdaptiveAvgPool2d(1) + Linear(512, 256) + ReLU + Linear(256, 40) in fp32 (TPS_STN.py:48-51, :73-76).  Eight crops per
// CTA share every weight load (one CTA per crop re-read the 512 KB of fc1 once per crop: L2-bound at 635 crops, and a
// single 512-long dependent chain per thread).  Pooling: thread = (8 channels, one of 8 position groups), 16-byte
// loads, all of a crop's loads in flight at once, the position groups added in a fixed order.  fc1: two K halves x 256
// outputs, the eight crops' inputs of one k as two 16-byte shared-memory reads; fc2: 16 K slices x 40 outputs; partial
// sums added in a fixed order.  A crop's result does not depend on which other crops share its CTA.
constexpr int kLocG = 8;

__global__ void __launch_bounds__(512)
loc_head_kernel(const uint16_t* __restrict__ feat, int B, int hw, const float* __restrict__ w1t,
                const float* __restrict__ b1, const float* __restrict__ w2t, const float* __restrict__ b2,
                float* __restrict__ fid, int f16, int split) {
    __shared__ __align__(16) float pooled_t[512][kLocG];          // [k][crop]
    __shared__ __align__(16) float scratch[16 * kLocG * 40];      // pooling partials [8][512] / fc1 partials [2][256][G] / fc2 partials [16][G][40]
    __shared__ __align__(16) float hid_t[256][kLocG];             // [k][crop]
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * kLocG;
    const int ng = B - b0 < kLocG ? B - b0 : kLocG;
    const int pitch = split ? 1024 : 512;
    {
        const int cg = tid & 63, pg = tid >> 6;
        for (int g = 0; g < kLocG; ++g) {
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            if (g < ng) {
                const uint16_t* f = feat + (long)(b0 + g) * hw * pitch + cg * 8;
                for (int p0 = pg; p0 < hw; p0 += 64) {
                    uint4 hv[8], lv[8];
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int pp = p0 + 8 * it;
                        hv[it] = make_uint4(0u, 0u, 0u, 0u);
                        lv[it] = make_uint4(0u, 0u, 0u, 0u);
                        if (pp < hw) {
                            hv[it] = __ldg(reinterpret_cast<const uint4*>(f + (long)pp * pitch));
                            if (split) lv[it] = __ldg(reinterpret_cast<const uint4*>(f + (long)pp * pitch + 512));
                        }
                    }
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        if (p0 + 8 * it < hw) {
                            float a[8], l[8];
                            unpack8(hv[it], a, f16);
                            unpack8(lv[it], l, f16);     // zeros when the tensor is not split
#pragma unroll
                            for (int i = 0; i < 8; ++i) acc[i] += a[i] + l[i];
                        }
                    }
                }
            }
            float4* dst = reinterpret_cast<float4*>(&scratch[pg * 512 + cg * 8]);
            dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            __syncthreads();
            {
                float sum = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) sum += scratch[q * 512 + tid];
                pooled_t[tid][g] = sum / (float)hw;
            }
            __syncthreads();
        }
    }
    {
        const int j = tid & 255, kh = tid >> 8;
        const float* w = w1t + (long)(kh * 256) * 256 + j;
        float sacc[kLocG];
#pragma unroll
        for (int g = 0; g < kLocG; ++g) sacc[g] = 0.f;
#pragma unroll 16
        for (int k = 0; k < 256; ++k) {
            const float wv = __ldg(w + (long)k * 256);
            const float4 pa = *reinterpret_cast<const float4*>(&pooled_t[kh * 256 + k][0]);
            const float4 pb = *reinterpret_cast<const float4*>(&pooled_t[kh * 256 + k][4]);
            sacc[0] = fmaf(pa.x, wv, sacc[0]); sacc[1] = fmaf(pa.y, wv, sacc[1]);
            sacc[2] = fmaf(pa.z, wv, sacc[2]); sacc[3] = fmaf(pa.w, wv, sacc[3]);
            sacc[4] = fmaf(pb.x, wv, sacc[4]); sacc[5] = fmaf(pb.y, wv, sacc[5]);
            sacc[6] = fmaf(pb.z, wv, sacc[6]); sacc[7] = fmaf(pb.w, wv, sacc[7]);
        }
        float4* dst = reinterpret_cast<float4*>(&scratch[(kh * 256 + j) * kLocG]);
        dst[0] = make_float4(sacc[0], sacc[1], sacc[2], sacc[3]);
        dst[1] = make_float4(sacc[4], sacc[5], sacc[6], sacc[7]);
    }
    __syncthreads();
    if (tid < 256) {
        const float bj = b1[tid];
#pragma unroll
        for (int g = 0; g < kLocG; ++g)
            hid_t[tid][g] = fmaxf((bj + scratch[tid * kLocG + g]) + scratch[(256 + tid) * kLocG + g], 0.f);
    }
    __syncthreads();
    {
        const int wp = tid >> 5, lane = tid & 31;     // 16 warps = 16 K slices of 16
        float s0[kLocG], s1[kLocG];
#pragma unroll
        for (int g = 0; g < kLocG; ++g) { s0[g] = 0.f; s1[g] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const int k = wp * 16 + kk;
            const float wa = __ldg(&w2t[k * 40 + lane]);
            const float wb = lane < 8 ? __ldg(&w2t[k * 40 + 32 + lane]) : 0.f;
            const float4 ha = *reinterpret_cast<const float4*>(&hid_t[k][0]);
            const float4 hb = *reinterpret_cast<const float4*>(&hid_t[k][4]);
            const float hv8[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
            for (int g = 0; g < kLocG; ++g) {
                s0[g] = fmaf(hv8[g], wa, s0[g]);
                s1[g] = fmaf(hv8[g], wb, s1[g]);
            }
        }
        // (the fc1 partials in `scratch` were consumed before the barrier above)
#pragma unroll
        for (int g = 0; g < kLocG; ++g) {
            scratch[(wp * kLocG + g) * 40 + lane] = s0[g];
            if (lane < 8) scratch[(wp * kLocG + g) * 40 + 32 + lane] = s1[g];
        }
    }
    __syncthreads();
    if (tid < kLocG * 40) {
        const int g = tid / 40, j = tid - g * 40;
        if (g < ng) {
            float sum = b2[j];
#pragma unroll
            for (int wq = 0; wq < 16; ++wq) sum += scratch[(wq * kLocG + g) * 40 + j];
            fid[(long)(b0 + g) * 40 + j] = sum;
        }
    }
}

// ------------------------------------------------------------------------------------------- TPS grid + sampling
__global__ void __launch_bounds__(256)
tps_sample_kernel(const float* __restrict__ fid, const float* __restrict__ inv_delta_c,
                  const float* __restrict__ p_hat_t, const float* __restrict__ x, float* __restrict__ out,
                  float* __restrict__ grid, int B) {
    constexpr int F = 20, F3 = 23, IH = 32, IW = 100, NP = IH * IW;
    __shared__ float T[F3][2];
    __shared__ float img[NP];
    const int b = blockIdx.x;
    if (threadIdx.x < F3 * 2) {
        const int i = threadIdx.x >> 1, d = threadIdx.x & 1;
        float s = 0.f;
        for (int j = 0; j < F; ++j) s = fmaf(inv_delta_c[i * F3 + j], fid[b * 40 + j * 2 + d], s);
        T[i][d] = s;  // the three appended zero rows of C' contribute nothing
    }
    for (int p = threadIdx.x; p < NP; p += blockDim.x) img[p] = x[(long)b * NP + p];
    __syncthreads();
    for (int p = threadIdx.x; p < NP; p += blockDim.x) {
        float gx = 0.f, gy = 0.f;
#pragma unroll
        for (int i = 0; i < F3; ++i) {
            const float ph = __ldg(&p_hat_t[i * NP + p]);
            gx = fmaf(ph, T[i][0], gx);
            gy = fmaf(ph, T[i][1], gy);
        }
        if (grid != nullptr) {
            grid[((long)b * NP + p) * 2] = gx;
            grid[((long)b * NP + p) * 2 + 1] = gy;
        }
        // grid_sampler_2d: align_corners=True un-normalisation, border padding = clip coordinates
        float ix = ((gx + 1.f) * 0.5f) * (IW - 1);
        float iy = ((gy + 1.f) * 0.5f) * (IH - 1);
        ix = fminf(fmaxf(ix, 0.f), (float)(IW - 1));
        iy = fminf(fmaxf(iy, 0.f), (float)(IH - 1));
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
        const float nw = (x1 - ix) * (y1 - iy), ne = (ix - x0) * (y1 - iy);
        const float sw_ = (x1 - ix) * (iy - y0), se = (ix - x0) * (iy - y0);
        float v = 0.f;
        if (x0 < IW && y0 < IH) v += img[y0 * IW + x0] * nw;
        if (x1 < IW && y0 < IH) v += img[y0 * IW + x1] * ne;
        if (x0 < IW && y1 < IH) v += img[y1 * IW + x0] * sw_;
        if (x1 < IW && y1 < IH) v += img[y1 * IW + x1] * se;
        out[(long)b * NP + p] = v;
    }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ------------------------------------------------------------------------------------------- attention decoder
constexpr int kAttT = 26;

// 512 threads: thread (j = tid & 255, kh = tid >> 8) owns hidden unit j and one half of every reduction over k (the
// kernel is issue-bound on those loops: twice the warps per crop group, half the trip count each), the 26 attention
// scores are split between the halves by time step and the context vectors by crop.
// kAttG = crops per CTA, chosen by the launcher so that all CTAs are resident at once (one CTA of 512 threads per SM):
// a second, nearly empty wave would double the kernel time.
template <int kAttG, bool W32>
__global__ void __launch_bounds__(512)
attention_kernel(const uint16_t* __restrict__ feats, const float* __restrict__ fproj, AttnWeights w,
                 float* __restrict__ preds, int B, int C, int f16, long feat_pitch, long feat_lo_off) {
    __shared__ float hs[kAttG][256];
    __shared__ float ctx[kAttG][256];
    __shared__ float e[kAttG][32];
    __shared__ float red[kAttG][kAttT][8];
    __shared__ float logit[kAttG][64];
    __shared__ float part[kAttG][256];            // upper-half partial sums of h2h
    __shared__ float gpart[kAttG][4][256];        // upper-half partial sums of the LSTMCell gates
    __shared__ int prev[kAttG];
    const int tid = threadIdx.x;
    const int j = tid & 255, kh = tid >> 8;
    const int lane = tid & 31, wp = (tid >> 5) & 7;
    const int b0 = blockIdx.x * kAttG;
    const int k0 = kh * 128;
    float c[kAttG];
#pragma unroll
    for (int g = 0; g < kAttG; ++g) {
        c[g] = 0.f;
        if (kh == 0) hs[g][j] = 0.f;
    }
    if (tid < kAttG) prev[tid] = 0;  // [GO]
    __syncthreads();
    const float sc = w.score_w[j];
    for (int step = 0; step < kAttT; ++step) {
        // (a) hp = h2h(h): each half sums its 128 k, the upper half hands its part over
        float hp[kAttG];
#pragma unroll
        for (int g = 0; g < kAttG; ++g) hp[g] = kh == 0 ? w.h2h_b[j] : 0.f;
#pragma unroll 8
        for (int k = k0; k < k0 + 128; ++k) {
            const float wv = W32 ? __ldg(&w.h2h_wt32[k * 256 + j]) : act2f(__ldg(&w.h2h_wt[k * 256 + j]), f16);
#pragma unroll
            for (int g = 0; g < kAttG; ++g) hp[g] = fmaf(wv, hs[g][k], hp[g]);
        }
        if (kh == 1) {
#pragma unroll
            for (int g = 0; g < kAttG; ++g) part[g][j] = hp[g];
        }
        __syncthreads();
        if (kh == 0) {
#pragma unroll
            for (int g = 0; g < kAttG; ++g) {
                hp[g] += part[g][j];
                part[g][j] = hp[g];      // the total, for the upper half
            }
        }
        __syncthreads();
        if (kh == 1) {
#pragma unroll
            for (int g = 0; g < kAttG; ++g) hp[g] = part[g][j];
        }
        // (b) e[t] = score . tanh(i2h(H)[t] + hp): time steps 0..12 on the lower half, 13..25 on the upper
#pragma unroll
        for (int g = 0; g < kAttG; ++g) {
            const int b = b0 + g;
            for (int t = kh * 13; t < kh * 13 + 13; ++t) {
                float v = 0.f;
                if (b < B) v = sc * tanhf(fproj[((long)b * kAttT + t) * 256 + j] + hp[g]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) red[g][t][wp] = v;
            }
        }
        __syncthreads();
        if (tid < kAttG * 32) {
            const int g = tid >> 5, t = tid & 31;
            float v = -INFINITY;
            if (t < kAttT) {
                v = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) v += red[g][t][q];
            }
            // (c) softmax over the 26 time steps (one warp per crop)
            float m = v;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float ex = (t < kAttT) ? expf(v - m) : 0.f;
            float s = ex;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            e[g][t] = ex / s;
        }
        __syncthreads();
        // (d) context = alpha^T H: the first half of the crops on the lower half of the threads, the rest on the upper
        constexpr int kSplit = (kAttG + 1) / 2;
#pragma unroll
        for (int gg = 0; gg < kSplit; ++gg) {
            const int g = kh * kSplit + gg;
            if (g >= kAttG) break;
            const int b = b0 + g;
            float v = 0.f;
            if (b < B) {
                for (int t = 0; t < kAttT; ++t) {
                    const uint16_t* fp = feats + ((long)b * kAttT + t) * feat_pitch + j;
                    float fv = act2f(fp[0], f16);
                    if (feat_lo_off > 0) fv += act2f(fp[feat_lo_off], f16);     // split-precision features
                    v = fmaf(e[g][t], fv, v);
                }
            }
            ctx[g][j] = v;
        }
        __syncthreads();
        // (e) LSTMCell gates: each half sums its 128 k of [context | h]
        float acc[kAttG][4];
#pragma unroll
        for (int g = 0; g < kAttG; ++g) {
            if (kh == 0) {
                const float4 oh = __ldg(reinterpret_cast<const float4*>(w.woh + ((size_t)prev[g] * 256 + j) * 4));
                acc[g][0] = w.gate_b[j] + oh.x;
                acc[g][1] = w.gate_b[256 + j] + oh.y;
                acc[g][2] = w.gate_b[512 + j] + oh.z;
                acc[g][3] = w.gate_b[768 + j] + oh.w;
            } else {
                acc[g][0] = acc[g][1] = acc[g][2] = acc[g][3] = 0.f;
            }
        }
#pragma unroll 4
        for (int k = k0; k < k0 + 128; ++k) {
            float wf[8];
            if (W32) {
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(w.wg32 + ((size_t)k * 256 + j) * 8));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(w.wg32 + ((size_t)k * 256 + j) * 8 + 4));
                wf[0] = w0.x; wf[1] = w0.y; wf[2] = w0.z; wf[3] = w0.w;
                wf[4] = w1.x; wf[5] = w1.y; wf[6] = w1.z; wf[7] = w1.w;
            } else {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(w.wg + ((size_t)k * 256 + j) * 8));
                unpack8(u, wf, f16);
            }
#pragma unroll
            for (int g = 0; g < kAttG; ++g) {
                const float cv = ctx[g][k], hv = hs[g][k];
                acc[g][0] = fmaf(wf[0], cv, fmaf(wf[4], hv, acc[g][0]));
                acc[g][1] = fmaf(wf[1], cv, fmaf(wf[5], hv, acc[g][1]));
                acc[g][2] = fmaf(wf[2], cv, fmaf(wf[6], hv, acc[g][2]));
                acc[g][3] = fmaf(wf[3], cv, fmaf(wf[7], hv, acc[g][3]));
            }
        }
        if (kh == 1) {
#pragma unroll
            for (int g = 0; g < kAttG; ++g)
#pragma unroll
                for (int q = 0; q < 4; ++q) gpart[g][q][j] = acc[g][q];
        }
        __syncthreads();  // all reads of hs done, upper-half partial sums visible
        if (kh == 0) {
#pragma unroll
            for (int g = 0; g < kAttG; ++g) {
                const float ig = sigmoidf_(acc[g][0] + gpart[g][0][j]), fg = sigmoidf_(acc[g][1] + gpart[g][1][j]);
                const float gt = tanhf(acc[g][2] + gpart[g][2][j]), og = sigmoidf_(acc[g][3] + gpart[g][3][j]);
                c[g] = fg * c[g] + ig * gt;
                hs[g][j] = og * tanhf(c[g]);
            }
        }
        __syncthreads();
        // (g) generator: one warp per class
        for (int idx = tid >> 5; idx < kAttG * C; idx += 16) {
            const int g = idx / C, v = idx - g * C;
            float s = 0.f;
            for (int k = lane; k < 256; k += 32) s = fmaf(__ldg(&w.gen_w[v * 256 + k]), hs[g][k], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) logit[g][v] = s + w.gen_b[v];
        }
        __syncthreads();
        // (h) store + greedy argmax (first maximum wins, like torch.max)
        for (int idx = tid; idx < kAttG * C; idx += 512) {
            const int g = idx / C, v = idx - g * C;
            if (b0 + g < B) preds[((long)(b0 + g) * kAttT + step) * C + v] = logit[g][v];
        }
        if (tid < kAttG) {
            float best = logit[tid][0];
            int bi = 0;
            for (int v = 1; v < C; ++v)
                if (logit[tid][v] > best) {
                    best = logit[tid][v];
                    bi = v;
                }
            prev[tid] = bi;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------- token decode
// decode_kernel below walks the 26 steps one after the other (two dependent passes over global memory and two shuffle
// reductions per step: ~33 us whatever the batch).  decode_smem_kernel (C <= 64, i.e. both heads of the reference) loads
// a crop's 26 x C logits into shared memory with independent coalesced loads and lets lane t handle step t on its own:
// arg-max (first maximum, like torch) and soft-max probability of all 26 steps in one pass.
constexpr int kDecodeMaxC = 64;

__device__ __forceinline__ void decode_tail(const int* idv, const float* pv, int head_attn, char* out, int text_stride,
                                            int32_t* has_eos, float* conf);

__global__ void __launch_bounds__(128)
decode_smem_kernel(const float* __restrict__ logits, int B, int C, int head_attn, int32_t* __restrict__ ids,
                   char* __restrict__ text, int text_stride, int32_t* __restrict__ has_eos, float* __restrict__ conf) {
    constexpr int T = 26;
    __shared__ float s_lg[4][T * kDecodeMaxC];
    __shared__ int s_id[4][32];
    __shared__ float s_p[4][32];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crop = blockIdx.x * 4 + wib;
    if (crop >= B) return;
    const float* lg = logits + (long)crop * T * C;
    float* sl = s_lg[wib];
    const int n_el = T * C;
    for (int i = lane; i < n_el; i += 32) sl[i] = lg[i];
    __syncwarp();
    if (lane < T) {
        const float* row = sl + lane * C;
        float best = row[0];
        int bi = 0;
        for (int v = 1; v < C; ++v) {
            const float x = row[v];
            if (x > best) {
                best = x;
                bi = v;
            }
        }
        float sum = 0.f;
        for (int v = 0; v < C; ++v) sum += expf(row[v] - best);
        s_id[wib][lane] = bi;
        s_p[wib][lane] = 1.f / sum;
        ids[crop * T + lane] = bi;
    }
    __syncwarp();
    if (lane != 0) return;
    decode_tail(s_id[wib], s_p[wib], head_attn, text + (long)crop * text_stride, text_stride, has_eos + crop, conf + crop);
}

__global__ void __launch_bounds__(128)
decode_kernel(const float* __restrict__ logits, int B, int C, int head_attn, int32_t* __restrict__ ids,
              char* __restrict__ text, int text_stride, int32_t* __restrict__ has_eos, float* __restrict__ conf) {
    constexpr int T = 26;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B) return;
    const float* lg = logits + (long)warp * T * C;
    int my_id = 0;        // lane t keeps the argmax of step t
    float my_p = 1.f;     // and its softmax probability
    for (int t = 0; t < T; ++t) {
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int v = lane; v < C; v += 32) {
            const float x = lg[t * C + v];
            if (x > best) {
                best = x;
                bi = v;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) {
                best = ob;
                bi = oi;
            }
        }
        float s = 0.f;
        for (int v = lane; v < C; v += 32) s += expf(lg[t * C + v] - best);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == t) {
            my_id = bi;
            my_p = 1.f / s;
        }
    }
    if (lane < T) ids[warp * T + lane] = my_id;
    // The sequential tail needs values held by other lanes: gather them through shuffles executed by the whole warp.
    int idv[T];
    float pv[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        idv[t] = __shfl_sync(0xffffffffu, my_id, t);
        pv[t] = __shfl_sync(0xffffffffu, my_p, t);
    }
    if (lane != 0) return;
    decode_tail(idv, pv, head_attn, text + (long)warp * text_stride, text_stride, has_eos + warp, conf + warp);
}

// The sequential part (one lane per crop): collapse / cut the token sequence into the string and multiply the step
// probabilities in step order (the reference's cumprod, net.py:177-190).
__device__ __forceinline__ void decode_tail(const int* idv, const float* pv, int head_attn, char* out, int text_stride,
                                            int32_t* has_eos, float* conf) {
    constexpr int T = 26;
    const char* alphabet = "0123456789abcdefghijklmnopqrstuvwxyz";
    int n = 0;
    if (!head_attn) {
        float p = 1.f;
        for (int t = 0; t < T; ++t) {
            const int id = idv[t];
            if (id != 0 && !(t > 0 && idv[t - 1] == id) && n < text_stride - 1) out[n++] = alphabet[id - 1];
            p *= pv[t];
        }
        out[n] = 0;
        *has_eos = 1;
        *conf = p;
    } else {
        // tokens: 0 = "[GO]", 1 = "[s]", 2.. = alphabet.  The reference cuts string and probabilities at the CHARACTER
        // index of the first "[s]" (net.py:184-186).
        int eos_tok = -1;
        for (int t = 0; t < T; ++t)
            if (idv[t] == 1) {
                eos_tok = t;
                break;
            }
        const int upto = eos_tok < 0 ? T : eos_tok;
        for (int t = 0; t < upto; ++t) {
            const int id = idv[t];
            if (id == 0) {
                const char* go = "[GO]";
                for (int q = 0; q < 4 && n < text_stride - 1; ++q) out[n++] = go[q];
            } else if (id >= 2 && n < text_stride - 1) {
                out[n++] = alphabet[id - 2];
            }
        }
        out[n] = 0;
        if (eos_tok < 0) {
            *has_eos = 0;
            *conf = 0.f;
        } else if (n == 0) {
            *has_eos = -1;  // reference: cumprod of an empty tensor, [-1] raises IndexError
            *conf = 0.f;
        } else {
            const int m = n < T ? n : T;  // n = character index of "[s]"
            float p = 1.f;
            for (int t = 0; t < m; ++t) p *= pv[t];
            *has_eos = 1;
            *conf = p;
        }
    }
}

// ------------------------------------------------------------------------------------------- evaluation losses
// evaluation() of the reference's training script (ocr/train/crnn.py:142-240) for one batch, given the logits the
// recogniser left in HBM: the loss of each crop (torch.nn.CTCLoss(zero_infinity=True) over preds.log_softmax(2), or
// CrossEntropyLoss(ignore_index=0) over the attention decoder's steps), and whether the greedy prediction equals the
// label.  One warp per crop; T = 26 steps.
constexpr int kEvalT = 26;
constexpr int kEvalMaxS = 2 * 64 + 1;   // CTC states of a target of up to 64 symbols

__device__ __forceinline__ float warp_lse(const float* __restrict__ row, int C, int lane) {
    float m = -INFINITY;
    for (int v = lane; v < C; v += 32) m = fmaxf(m, row[v]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int v = lane; v < C; v += 32) s += expf(row[v] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return m + logf(s);
}

// targets: concatenated class indices (CTCLabelConverter.encode, recog_utils.py:24-30), tgt_off[b] = start of crop b's,
// tgt_len[b] its length.  loss[b] = -log p(target | logits) (0 when infeasible: zero_infinity), the alpha recursion of
// the connectionist temporal classification forward pass in log space, fp32 like torch's CPU kernel.
// correct[b] = 1 iff the collapsed arg-max path (ids, from decode_kernel) spells the target.
__global__ void __launch_bounds__(128)
ctc_loss_kernel(const float* __restrict__ logits, int B, int C, const int32_t* __restrict__ targets,
                const int32_t* __restrict__ tgt_off, const int32_t* __restrict__ tgt_len,
                const int32_t* __restrict__ ids, float* __restrict__ loss, int32_t* __restrict__ correct) {
    __shared__ float alpha_s[4][2][kEvalMaxS + 3];
    __shared__ float lse_s[4][kEvalT];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + wib;
    if (b >= B) return;
    const float* lg = logits + (long)b * kEvalT * C;
    const int32_t* tg = targets + tgt_off[b];
    const int L = tgt_len[b];
    const int S = 2 * L + 1;
    if (L > 64) {                 // longer than any 26-step path anyway: infeasible
        if (lane == 0) {
            loss[b] = 0.f;
            correct[b] = 0;
        }
        return;
    }
    for (int t = 0; t < kEvalT; ++t) {
        const float l = warp_lse(lg + t * C, C, lane);
        if (lane == 0) lse_s[wib][t] = l;
    }
    __syncwarp();
    float* a0 = alpha_s[wib][0];
    float* a1 = alpha_s[wib][1];
    for (int s = lane; s < S; s += 32) {
        float v = -INFINITY;
        if (s == 0) v = lg[0] - lse_s[wib][0];                         // blank = class 0
        else if (s == 1) v = lg[tg[0]] - lse_s[wib][0];
        a0[s] = v;
    }
    __syncwarp();
    for (int t = 1; t < kEvalT; ++t) {
        const float* row = lg + t * C;
        const float l = lse_s[wib][t];
        for (int s = lane; s < S; s += 32) {
            const int cls = (s & 1) ? tg[s >> 1] : 0;
            const float la1 = a0[s];
            const float la2 = s > 0 ? a0[s - 1] : -INFINITY;
            const float la3 = ((s & 1) && s > 1 && tg[s >> 1] != tg[(s >> 1) - 1]) ? a0[s - 2] : -INFINITY;
            float mx = fmaxf(la1, fmaxf(la2, la3));
            if (mx == -INFINITY) mx = 0.f;
            a1[s] = logf(expf(la1 - mx) + expf(la2 - mx) + expf(la3 - mx)) + mx + (row[cls] - l);
        }
        __syncwarp();
        float* tmp = a0; a0 = a1; a1 = tmp;
    }
    if (lane == 0) {
        const float l1 = a0[S - 1];
        const float l2 = L > 0 ? a0[S - 2] : -INFINITY;
        float mx = fmaxf(l1, l2);
        if (mx == -INFINITY) mx = 0.f;
        float v = -(logf(expf(l1 - mx) + expf(l2 - mx)) + mx);
        if (isinf(v) || isnan(v)) v = 0.f;                             // zero_infinity=True
        loss[b] = v;
        // greedy path collapsed like CTCLabelConverter.decode (recog_utils.py:32-47) against the label
        const int32_t* id = ids + b * kEvalT;
        int n = 0, ok = 1;
        for (int t = 0; t < kEvalT; ++t) {
            const int c = id[t];
            if (c != 0 && !(t > 0 && id[t - 1] == c)) {
                if (n >= L || tg[n] != c) ok = 0;
                ++n;
            }
        }
        correct[b] = (ok && n == L) ? 1 : 0;
    }
}

// targets [B][tw]: AttnLabelConverter.encode rows (recog_utils.py:84-96): [GO], the label's tokens, [s], then [GO]
// padding.  The loss compares step t of the decoder with targets[b][t + 1] and ignores padding (class 0):
// loss[b] = sum over the counted steps of -log_softmax(preds[b][t])[target], count[b] = their number.
// correct[b] follows evaluation()'s string comparison: label and prediction are cut at their first "[s]"; a prediction
// without "[s]" loses its last CHARACTER instead (str.find returns -1 there, crnn.py:226-228).
__global__ void __launch_bounds__(128)
attn_ce_kernel(const float* __restrict__ logits, int B, int C, const int32_t* __restrict__ targets, int tw,
               const int32_t* __restrict__ ids, float* __restrict__ loss, int32_t* __restrict__ count,
               int32_t* __restrict__ correct) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    const float* lg = logits + (long)b * kEvalT * C;
    const int32_t* tg = targets + (long)b * tw;
    const int steps = tw - 1 < kEvalT ? tw - 1 : kEvalT;
    float sum = 0.f;
    int cnt = 0;
    for (int t = 0; t < steps; ++t) {
        const int c = tg[t + 1];
        if (c == 0) continue;                       // warp-uniform
        const float l = warp_lse(lg + t * C, C, lane);
        sum += l - lg[t * C + c];
        ++cnt;
    }
    if (lane != 0) return;
    loss[b] = sum;
    count[b] = cnt;
    const int32_t* id = ids + b * kEvalT;
    int gl = 0;                                     // label tokens before its [s]
    while (gl < steps && tg[gl + 1] != 1) ++gl;
    int pe = -1;
    for (int t = 0; t < kEvalT; ++t)
        if (id[t] == 1) {
            pe = t;
            break;
        }
    int ok = 1;
    if (pe >= 0) {
        if (pe != gl) ok = 0;
        for (int t = 0; ok && t < pe; ++t)
            if (id[t] != tg[t + 1]) ok = 0;
    } else {
        // 26 tokens minus the last character: equal to a label only if every token is a single character
        if (gl != kEvalT - 1 || id[kEvalT - 1] < 2) ok = 0;
        for (int t = 0; ok && t < kEvalT - 1; ++t)
            if (id[t] != tg[t + 1]) ok = 0;
    }
    correct[b] = ok;
}

// ------------------------------------------------------------------------------------------- range audit
__global__ void __launch_bounds__(256)
absmax_kernel(const uint16_t* __restrict__ t, long rows, int C, long pitch, int f16, float* __restrict__ slot) {
    const int groups = C >> 3;
    const long total = rows * groups;
    float m = 0.f;
    for (long gid = (long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long)gridDim.x * blockDim.x) {
        const long r = gid / groups;
        const int g = (int)(gid - r * groups);
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(t + r * pitch + g * 8));
        float f[8];
        unpack8(u, f, f16);
#pragma unroll
        for (int j = 0; j < 8; ++j) m = fmaxf(m, fabsf(f[j]));     // fmaxf drops NaN: an infinity shows up as inf
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(slot), __float_as_uint(m));
}

inline int grid_for(long total, int block) {
    long g = (total + block - 1) / block;
    const long cap = 148L * 16;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

void launch_direct_conv3x3(const void* in, int u8_mode, int B, int H, int W, int img_h, int img_w, long row_stride,
                           long img_stride, const float* w, const float* bias, int Cin, int Cout, void* out,
                           long out_pitch, int relu, int is_f16, cudaStream_t s, int split_out, int pool, int out_row_px) {
    const long total = (long)B * (pool ? H / 2 : H) * (pool ? W / 2 : W) * (Cout / 32);
    const int grid = grid_for(total, 256);
    const size_t smem = (size_t)(9 * Cin * Cout + Cout) * sizeof(float);
    if (u8_mode)
        direct_conv3x3_kernel<3, true, false><<<grid, 256, smem, s>>>(in, B, H, W, img_h, img_w, row_stride, img_stride,
                                                                      w, bias, Cout, (uint16_t*)out, out_pitch, relu,
                                                                      is_f16, split_out);
    else if (pool && Cin == 1 && Cout % 16 == 0 && H % 2 == 0 && W % 2 == 0 && !getenv("LOCR_DIRECT_POOL_OLD")) {
        const long total4 = (long)B * (H / 2) * (W / 2) * (Cout / 16);
        direct_conv3x3_pool4_kernel<true><<<grid_for(total4, 256), 256, smem, s>>>((const float*)in, B, H, W, w, bias,
                                                                                   Cout, (uint16_t*)out, out_pitch, relu,
                                                                                   is_f16, split_out, 0);
    } else if (!pool && !u8_mode && Cin == 1 && Cout % 16 == 0 && H % 2 == 0 && W % 2 == 0 && !getenv("LOCR_DIRECT_POOL_OLD")) {
        const long total4 = (long)B * (H / 2) * (W / 2) * (Cout / 16);
        direct_conv3x3_pool4_kernel<false><<<grid_for(total4, 256), 256, smem, s>>>((const float*)in, B, H, W, w, bias,
                                                                                    Cout, (uint16_t*)out, out_pitch, relu,
                                                                                    is_f16, split_out, out_row_px);
    } else if (pool)
        direct_conv3x3_kernel<1, false, true><<<grid, 256, smem, s>>>(in, B, H, W, H, W, 0, 0, w, bias, Cout,
                                                                      (uint16_t*)out, out_pitch, relu, is_f16, split_out);
    else
        direct_conv3x3_kernel<1, false, false><<<grid, 256, smem, s>>>(in, B, H, W, H, W, 0, 0, w, bias, Cout,
                                                                       (uint16_t*)out, out_pitch, relu, is_f16, split_out);
}

void launch_absmax(const void* t, long rows, int C, long pitch, int is_f16, float* slot, cudaStream_t s) {
    if (rows <= 0 || C < 8) return;
    absmax_kernel<<<grid_for(rows * (C >> 3), 256), 256, 0, s>>>((const uint16_t*)t, rows, C, pitch, is_f16, slot);
}

void launch_preproc_nhwc16(const uint8_t* in, int B, int H, int W, int img_h, int img_w, long row_stride,
                           long img_stride, void* out, int is_f16, cudaStream_t s, int channels) {
    if (channels == 8) {
        const long total8 = (long)B * H * (W + 3);
        preproc_nhwc8_kernel<<<grid_for(total8, 256), 256, 0, s>>>(in, B, H, W, img_h, img_w, row_stride, img_stride,
                                                                   (uint16_t*)out, is_f16);
        return;
    }
    const long total = (long)B * H * (W + 3) * 2;
    preproc_nhwc16_kernel<<<grid_for(total, 256), 256, 0, s>>>(in, B, H, W, img_h, img_w, row_stride, img_stride,
                                                               (uint16_t*)out, is_f16);
}

void launch_maxpool(const void* in, long in_pitch, int B, int H, int W, int C, void* out, long out_pitch, int kh,
                    int kw, int sh, int sw, int ph, int pw, int is_f16, cudaStream_t s, int split) {
    const int OH = (H + 2 * ph - kh) / sh + 1, OW = (W + 2 * pw - kw) / sw + 1;
    const long total = (long)B * OH * OW * (C / 8);
    maxpool_kernel<<<grid_for(total, 256), 256, 0, s>>>((const uint16_t*)in, in_pitch, B, H, W, C, (uint16_t*)out,
                                                        out_pitch, OH, OW, kh, kw, sh, sw, ph, pw, is_f16, split);
}

void launch_upsample2x(const void* in, long in_pitch, int B, int H, int W, int C, void* out, long out_pitch,
                       int is_f16, cudaStream_t s) {
    const long total = (long)B * H * W * (C / 8);
    upsample2x_kernel<<<grid_for(total, 256), 256, 0, s>>>((const uint16_t*)in, in_pitch, B, H, W, C, (uint16_t*)out,
                                                           out_pitch, is_f16);
}

void launch_loc_head(const void* feat, int B, int hw, const float* w1t, const float* b1, const float* w2t,
                     const float* b2, float* fid, int is_f16, cudaStream_t s, int split) {
    loc_head_kernel<<<(B + kLocG - 1) / kLocG, 512, 0, s>>>((const uint16_t*)feat, B, hw, w1t, b1, w2t, b2, fid, is_f16, split);
}

void launch_tps_sample(const float* fid, const float* inv_delta_c, const float* p_hat_t, const float* x, float* out,
                       float* grid, int B, cudaStream_t s) {
    tps_sample_kernel<<<B, 256, 0, s>>>(fid, inv_delta_c, p_hat_t, x, out, grid, B);
}

void launch_attention(const void* feats, const float* fproj, AttnWeights w, float* preds, int B, int C, int is_f16,
                      cudaStream_t s, long feat_pitch, long feat_lo_off) {
    if (feat_pitch <= 0) feat_pitch = 256;
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const uint16_t* f = (const uint16_t*)feats;
    // 5 crops per CTA when that brings the grid down to one resident wave (148 < B / 4, B / 5 <= 148), else 4
    const bool w32 = w.h2h_wt32 != nullptr && w.wg32 != nullptr;
    if ((B + 3) / 4 > sms && (B + 4) / 5 <= sms) {
        if (w32) attention_kernel<5, true><<<(B + 4) / 5, 512, 0, s>>>(f, fproj, w, preds, B, C, is_f16, feat_pitch, feat_lo_off);
        else attention_kernel<5, false><<<(B + 4) / 5, 512, 0, s>>>(f, fproj, w, preds, B, C, is_f16, feat_pitch, feat_lo_off);
    } else {
        if (w32) attention_kernel<4, true><<<(B + 3) / 4, 512, 0, s>>>(f, fproj, w, preds, B, C, is_f16, feat_pitch, feat_lo_off);
        else attention_kernel<4, false><<<(B + 3) / 4, 512, 0, s>>>(f, fproj, w, preds, B, C, is_f16, feat_pitch, feat_lo_off);
    }
}

void launch_ctc_loss(const float* logits, int B, int C, const int32_t* targets, const int32_t* tgt_off,
                     const int32_t* tgt_len, const int32_t* ids, float* loss, int32_t* correct, cudaStream_t s) {
    ctc_loss_kernel<<<(B + 3) / 4, 128, 0, s>>>(logits, B, C, targets, tgt_off, tgt_len, ids, loss, correct);
}

void launch_attn_ce(const float* logits, int B, int C, const int32_t* targets, int tw, const int32_t* ids, float* loss,
                    int32_t* count, int32_t* correct, cudaStream_t s) {
    attn_ce_kernel<<<(B + 3) / 4, 128, 0, s>>>(logits, B, C, targets, tw, ids, loss, count, correct);
}

void launch_decode(const float* logits, int B, int C, int head_attn, int32_t* ids, char* text, int text_stride,
                   int32_t* has_eos, float* conf, cudaStream_t s) {
    const int warps_per_block = 4;
    if (C <= kDecodeMaxC)
        decode_smem_kernel<<<(B + warps_per_block - 1) / warps_per_block, 128, 0, s>>>(logits, B, C, head_attn, ids, text,
                                                                                       text_stride, has_eos, conf);
    else
        decode_kernel<<<(B + warps_per_block - 1) / warps_per_block, 128, 0, s>>>(logits, B, C, head_attn, ids, text,
                                                                                  text_stride, has_eos, conf);
}

}  // namespace locr
