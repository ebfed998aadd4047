// BiLSTM recurrence (reference ocr/modules/biLSTM.py:18,24 = nn.LSTM(bidirectional, batch_first)) as a persistent
// thread-block-cluster kernel with the recurrent weights resident in shared memory.
//
// One cluster of 8 CTAs = 128 crops x one direction, persistent over the T time steps.  W_hh of one direction is
// 1024 x 256 16-bit = 512 KB, more than one SM holds, so the cluster splits it by hidden unit: CTA r keeps the 128
// gate rows (i, f, g, o of units 32r .. 32r+31) x 256 k = 64 KB in its shared memory for the whole kernel.  Per step
//     gates[128 crops, 128] = xproj_t (precomputed GEMM, fp32) + h_{t-1}[128, 256] * W_r^T
// is one 128 x 128 x 256 tcgen05 GEMM (16 MMAs) with the accumulator in TMEM:
//   * A operand = h_{t-1} of the cluster's 128 crops, [4 k-chunks][128 crops][64] in the UMMA K-major / 128B-swizzle
//     layout, double buffered.  Every CTA emits its 32 units of h_t straight into the layer output (global memory,
//     which the next layer needs anyway); the 8 CTAs then pull the full h_t row block back with four TMA loads each
//     (L2 hits; the TMA unit applies the swizzle and zero-fills crops past the batch).  The hand-over is a
//     cluster-scope mbarrier per A buffer: every epilogue warp of every CTA arrives (release.cluster) on all 8 CTAs'
//     barriers after its stores, the TMA-issuing thread waits (acquire.cluster) for the 64 arrivals;
//   * B operand = the resident W_hh slice;
//   * 8 epilogue warps: thread = (crop, 16 units) keeps its fp32 cell state in registers for all T steps, applies the
//     gate non-linearities and writes h_t (16-bit, like every other activation of the path).
// Step 0 has h = 0: no MMA, the gates are xproj alone.
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-11 = gate math.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "nn_kernels.cuh"
#include "ptx.cuh"

#ifndef LOCR_LSTM_TRACE
#define LOCR_LSTM_TRACE 0   // tools only: per-step clock64 stamps of CTA (0, 0) into g_lstm_trace (locr_debug_lstm_trace)
#endif

namespace locr {

#if LOCR_LSTM_TRACE
__device__ long long g_lstm_trace[64 * 8];
#define LSTM_STAMP(step, slot) \
    do { if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) g_lstm_trace[(step) * 8 + (slot)] = clock64(); } while (0)
#else
#define LSTM_STAMP(step, slot) do {} while (0)
#endif

namespace {

constexpr int kCluster = 8;
constexpr int kChunkBytes = 128 * 128;          // [128 rows][64 k] 16-bit
constexpr int kWBytes = 4 * kChunkBytes;        // resident W_hh slice
constexpr int kABytes = 4 * kChunkBytes;        // h operand, one buffer
constexpr int kThreadsLstm = 384;
constexpr int kEpiWarps = 8;

struct LstmParams {
    const float* xproj;   // [B][T][2048], column = dir*1024 + unit*4 + gate
    uint16_t* out;        // [B][T][pitch]: 512 hidden states (+ their 512 lo halves in split-precision mode)
    int B, T, is_f16;
    int pitch;            // 512, or 1024 when SPLIT
    uint32_t idesc;
};

// Accurate to a few ulp (ex2.approx + rcp.approx); arguments are clamped so that no intermediate overflows.
__device__ __forceinline__ float sigmoid_f(float x) {
    x = fminf(fmaxf(x, -30.f), 30.f);
    return __fdividef(1.f, 1.f + __expf(-x));
}
__device__ __forceinline__ float tanh_f(float x) {
    x = fminf(fmaxf(x, -15.f), 15.f);
    return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x));
}
__device__ __forceinline__ uint32_t pack2h(float a, float b, int f16) {
    if (f16) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// SPLIT (LOCR_PREC_EXACT): h_t is carried as a hi + lo pair of 16-bit numbers - both halves go to the layer output
// ([hi 512 | lo 512] per time step) and both are pulled back as A operands (buffer 0 = hi, buffer 1 = lo; a single
// buffer per half is enough: the MMAs of step s have retired before any epilogue warp can publish h_s), so the step's
// GEMM is h_hi W^T + h_lo W^T and the 26-step feedback no longer rounds h to 11 bits.
template <bool SPLIT>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreadsLstm, 1)
lstm_cluster_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
                    const __grid_constant__ CUtensorMap tmap_hlo, const LstmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* w_buf = smem;                              // kWBytes
    uint8_t* a_buf = smem + kWBytes;                    // 2 x kABytes
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_buf + 2 * kABytes);
    uint64_t* wfull_bar = bars;                         // [1]  W slice landed
    uint64_t* afull_bar = bars + 1;                     // [2][4]  h chunk landed
    uint64_t* hready_bar = bars + 9;                    // [2]  all 64 epilogue warps of the cluster stored h
    uint64_t* tfull_bar = bars + 11;                    // [1]  accumulator complete
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)ptx::cluster_ctarank();
    const int dir = blockIdx.y;
    const int crop0 = (blockIdx.x / kCluster) * 128;
    const int T = p.T;

    if (warp == 0 && lane == 0) {
        ptx::tma_prefetch_desc(&tmap_w);
        ptx::tma_prefetch_desc(&tmap_h);
        if (SPLIT) ptx::tma_prefetch_desc(&tmap_hlo);
    }
    if (warp == 1 && lane == 0) {
        ptx::mbar_init(wfull_bar, 1);
        for (int i = 0; i < 8; ++i) ptx::mbar_init(&afull_bar[i], 1);
        ptx::mbar_init(&hready_bar[0], kCluster * kEpiWarps);
        ptx::mbar_init(&hready_bar[1], kCluster * kEpiWarps);
        ptx::mbar_init(tfull_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_ptr_smem, 128);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    // every CTA's barriers exist before anyone arrives on them remotely
    ptx::cluster_arrive_release();
    ptx::cluster_wait_acquire();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // Producer and MMA-issuer warps run their loops on warp-uniform values; one elected lane issues the TMA / MMA
    // instructions (a loop inside `if (lane == 0)` makes ptxas wrap every UTMALDG / UTCHMMA in an election loop).
    if (warp == 0) {
        if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(wfull_bar, kWBytes);
            for (int kc = 0; kc < 4; ++kc)
                ptx::tma_load_2d(w_buf + kc * kChunkBytes, &tmap_w, wfull_bar, kc * 64, dir * 1024 + rank * 128);
        }
        for (int step = 1; step < T; ++step) {
            const int b = step & 1;
            const uint32_t par = (uint32_t)(((step - 1) >> 1) & 1);
            const int t_prev = dir == 0 ? step - 1 : T - step;
            ptx::mbar_wait_cluster(&hready_bar[b], par, 500);
            LSTM_STAMP(step, 0);                      // all 64 arrivals seen
            if (ptx::elect_one()) {
                ptx::fence_proxy_async_all();
                if (SPLIT) {
#pragma unroll
                    for (int hl = 0; hl < 2; ++hl)
#pragma unroll
                        for (int kc = 0; kc < 4; ++kc) {
                            ptx::mbar_arrive_expect_tx(&afull_bar[hl * 4 + kc], kChunkBytes);
                            ptx::tma_load_3d(a_buf + hl * kABytes + kc * kChunkBytes, hl ? &tmap_hlo : &tmap_h,
                                             &afull_bar[hl * 4 + kc], dir * 256 + kc * 64, t_prev, crop0);
                        }
                } else {
#pragma unroll
                    for (int kc = 0; kc < 4; ++kc) {
                        ptx::mbar_arrive_expect_tx(&afull_bar[b * 4 + kc], kChunkBytes);
                        ptx::tma_load_3d(a_buf + b * kABytes + kc * kChunkBytes, &tmap_h, &afull_bar[b * 4 + kc],
                                         dir * 256 + kc * 64, t_prev, crop0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t desc_hi = (uint32_t)(ptx::make_kmajor_desc(0, 128) >> 32);
        const uint32_t a_lo0 = ((ptx::smem_u32(a_buf) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t w_lo0 = ((ptx::smem_u32(w_buf) & 0x3FFFFu) >> 4) | (1u << 16);
        ptx::mbar_wait(wfull_bar, 0, 600);
        for (int step = 1; step < T; ++step) {
            const int b = step & 1;
            const uint32_t par = (uint32_t)(((step - 1) >> 1) & 1);
            if (SPLIT) {
                // both buffers every step: the barriers' phase flips once per step
                const uint32_t par_s = (uint32_t)((step - 1) & 1);
#pragma unroll
                for (int hl = 0; hl < 2; ++hl)
#pragma unroll
                    for (int kc = 0; kc < 4; ++kc) {
                        ptx::mbar_wait(&afull_bar[hl * 4 + kc], par_s, 610);
                        ptx::tc_fence_after();
                        const uint32_t a_lo = a_lo0 + (uint32_t)((hl * kABytes + kc * kChunkBytes) >> 4);
                        const uint32_t w_lo = w_lo0 + (uint32_t)((kc * kChunkBytes) >> 4);
                        if (ptx::elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                ptx::umma_f16_lohi(tmem_base, a_lo + k * 2, desc_hi, w_lo + k * 2, desc_hi, p.idesc,
                                                   (hl | kc | k) ? 1u : 0u);
                            if (hl == 1 && kc == 3) ptx::umma_commit(tfull_bar);
                        }
                    }
            } else {
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
                ptx::mbar_wait(&afull_bar[b * 4 + kc], par, 610);
                ptx::tc_fence_after();
                const uint32_t a_lo = a_lo0 + (uint32_t)((b * kABytes + kc * kChunkBytes) >> 4);
                const uint32_t w_lo = w_lo0 + (uint32_t)((kc * kChunkBytes) >> 4);
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_f16_lohi(tmem_base, a_lo + k * 2, desc_hi, w_lo + k * 2, desc_hi, p.idesc,
                                           (kc | k) ? 1u : 0u);
                    if (kc == 3) ptx::umma_commit(tfull_bar);
                }
                if (kc == 0) LSTM_STAMP(step, 1);     // first h chunk landed
                if (kc == 3) LSTM_STAMP(step, 2);     // last chunk landed, MMAs issued
            }
            }
        }
    } else if (warp >= 4) {
        const int quarter = warp & 3;                 // TMEM lane quarter
        const int chalf = (warp - 4) >> 2;            // which 64 of the CTA's 128 gate columns (16 units)
        const int row = quarter * 32 + lane;          // crop within the cluster's block
        const int crop = crop0 + row;
        const bool valid = crop < p.B;
        const int ucol = rank * 32 + chalf * 16;      // first hidden unit of this thread
        const float* xp_base = p.xproj + (size_t)(valid ? crop : 0) * T * 2048 + dir * 1024 + ucol * 4;
        uint16_t* out_base = p.out + (size_t)(valid ? crop : 0) * T * p.pitch + dir * 256 + ucol;
        uint32_t hready_remote[2] = {0, 0};
        if (lane < kCluster) {
            hready_remote[0] = ptx::mapa(ptx::smem_u32(&hready_bar[0]), (uint32_t)lane);
            hready_remote[1] = ptx::mapa(ptx::smem_u32(&hready_bar[1]), (uint32_t)lane);
        }
        float cc[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) cc[u] = 0.f;
        float4 x[16];
        {
            const int t0 = dir == 0 ? 0 : T - 1;
            const float4* xp4 = reinterpret_cast<const float4*>(xp_base + (size_t)t0 * 2048);
#pragma unroll
            for (int u = 0; u < 16; ++u) x[u] = valid ? __ldg(xp4 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int step = 0; step < T; ++step) {
            const int t = dir == 0 ? step : T - 1 - step;
            if (step > 0) {
                ptx::mbar_wait(tfull_bar, (uint32_t)((step - 1) & 1), 700);
                ptx::tc_fence_after();
            }
            if (warp == 4) LSTM_STAMP(step, 3);       // accumulator complete
            uint32_t hp[8], hl_[8];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t r[32];
                if (step > 0) {
                    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(chalf * 64 + hf * 32), r);
                    ptx::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                float hv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 xv = x[hf * 8 + u];
                    const float gi = sigmoid_f(__uint_as_float(r[u * 4 + 0]) + xv.x);
                    const float gf = sigmoid_f(__uint_as_float(r[u * 4 + 1]) + xv.y);
                    const float gg = tanh_f(__uint_as_float(r[u * 4 + 2]) + xv.z);
                    const float go = sigmoid_f(__uint_as_float(r[u * 4 + 3]) + xv.w);
                    const float c = gf * cc[hf * 8 + u] + gi * gg;
                    cc[hf * 8 + u] = c;
                    hv[u] = go * tanh_f(c);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t hi = pack2h(hv[q * 2], hv[q * 2 + 1], p.is_f16);
                    hp[hf * 4 + q] = hi;
                    if (SPLIT) {
                        const float2 back = p.is_f16 ? __half22float2(*reinterpret_cast<const __half2*>(&hi))
                                                     : __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi));
                        hl_[hf * 4 + q] = pack2h(hv[q * 2] - back.x, hv[q * 2 + 1] - back.y, p.is_f16);
                    }
                }
            }
            if (warp == 4) LSTM_STAMP(step, 4);       // gate math done
            if (valid) {
                uint4* o = reinterpret_cast<uint4*>(out_base + (size_t)t * p.pitch);
                o[0] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                o[1] = make_uint4(hp[4], hp[5], hp[6], hp[7]);
                if (SPLIT) {
                    o[64] = make_uint4(hl_[0], hl_[1], hl_[2], hl_[3]);      // + 512 elements
                    o[65] = make_uint4(hl_[4], hl_[5], hl_[6], hl_[7]);
                }
            }
            if (step + 1 < T) {
                // h_t is in global memory (generic proxy); the peers read it through TMA (async proxy)
                ptx::fence_proxy_async_all();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane < kCluster) ptx::mbar_arrive_remote_release(hready_remote[(step + 1) & 1]);
                if (warp == 4) LSTM_STAMP(step, 5);   // h stored, peers signalled
                // gate inputs of the next step: in flight while the cluster exchanges h_t and the MMAs run
                const int tn = dir == 0 ? step + 1 : T - 2 - step;
                const float4* xp4 = reinterpret_cast<const float4*>(xp_base + (size_t)tn * 2048);
#pragma unroll
                for (int u = 0; u < 16; ++u) x[u] = valid ? __ldg(xp4 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 128);
    }
    // no CTA leaves while a peer could still signal it
    ptx::cluster_arrive_release();
    ptx::cluster_wait_acquire();
}

// ---------------------------------------------------------------------------------------------------------------------
// Push form of the hand-over (fast arithmetic; LOCR_LSTM_MCAST=0 switches back to the pull form above).  In the pull
// form every step pays: 8 remote arrivals per epilogue warp, a wait for all 64 warps of the cluster, and a 64 KB
// pull-back of h per CTA - 2/3 of the step.  Here the K range is cut into EIGHT 32-unit chunks ([128 rows][64 B],
// 64-byte swizzle), i.e. exactly one chunk per CTA of the cluster: as soon as the 8 epilogue warps of CTA r have stored
// their 32 units of h_t (one CTA-local mbarrier), its producer issues ONE multicast TMA load of that 8 KB slice, which
// lands as chunk r of the A buffer in all 8 CTAs and completes chunk r's barrier in each of them.  No cluster-wide
// arrival round, every CTA reads 8 KB instead of 64 KB per step, and the MMAs of a chunk start when that chunk lands.
// The A buffer is double buffered; a slice for step s + 1 can only be sent after the sender's epilogue of step s, which
// needed every peer's slice of step s - 1, which each peer published after its own MMAs of step s - 1 had retired: the
// buffer a multicast writes (last read at step s - 1) is free in every CTA, and barrier phases cannot alias.
constexpr int kSliceBytes = 128 * 64;           // [128 crops][32 units] 16-bit = one CTA's share of h_t

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreadsLstm, 1)
lstm_mcast_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
                  const LstmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* w_buf = smem;                              // 8 chunks x kSliceBytes
    uint8_t* a_buf = smem + kWBytes;                    // 2 x (8 chunks x kSliceBytes)
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_buf + 2 * kABytes);
    uint64_t* wfull_bar = bars;                         // [1]  W slice landed
    uint64_t* afull_bar = bars + 1;                     // [2][8]  chunk kc of h landed (sent by CTA kc)
    uint64_t* hlocal_bar = bars + 17;                   // [1]  this CTA's 8 epilogue warps stored their h slice
    uint64_t* tfull_bar = bars + 18;                    // [1]  accumulator complete
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 19);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)ptx::cluster_ctarank();
    const int dir = blockIdx.y;
    const int crop0 = (blockIdx.x / kCluster) * 128;
    const int T = p.T;

    if (warp == 0 && lane == 0) {
        ptx::tma_prefetch_desc(&tmap_w);
        ptx::tma_prefetch_desc(&tmap_h);
    }
    if (warp == 1 && lane == 0) {
        ptx::mbar_init(wfull_bar, 1);
        for (int i = 0; i < 16; ++i) ptx::mbar_init(&afull_bar[i], 1);
        ptx::mbar_init(hlocal_bar, kEpiWarps);
        ptx::mbar_init(tfull_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_ptr_smem, 128);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    // every CTA's barriers exist before a peer's multicast can signal them
    ptx::cluster_arrive_release();
    ptx::cluster_wait_acquire();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(wfull_bar, kWBytes);
            for (int kc = 0; kc < 8; ++kc)
                ptx::tma_load_2d(w_buf + kc * kSliceBytes, &tmap_w, wfull_bar, kc * 32, dir * 1024 + rank * 128);
        }
        for (int step = 1; step < T; ++step) {
            const int b = step & 1;
            const int t_prev = dir == 0 ? step - 1 : T - step;
            // arm this step's eight chunk barriers (their previous phase, step - 2, was consumed before this CTA's
            // epilogue of step - 2 could publish, which the previous iteration waited for)
            if (ptx::elect_one()) {
#pragma unroll
                for (int kc = 0; kc < 8; ++kc) ptx::mbar_arrive_expect_tx(&afull_bar[b * 8 + kc], kSliceBytes);
            }
            __syncwarp();
            ptx::mbar_wait(hlocal_bar, (uint32_t)((step - 1) & 1), 500);     // h_{step-1}: this CTA's 32 units are in global memory
            LSTM_STAMP(step, 0);
            if (ptx::elect_one()) {
                ptx::fence_proxy_async_all();
                ptx::tma_load_3d_mcast(a_buf + b * kABytes + rank * kSliceBytes, &tmap_h, &afull_bar[b * 8 + rank],
                                       dir * 256 + rank * 32, t_prev, crop0, (uint16_t)0xFF);
            }
        }
    } else if (warp == 1) {
        const uint32_t desc_hi = (uint32_t)(ptx::make_kmajor_desc(0, 64) >> 32);
        const uint32_t a_lo0 = ((ptx::smem_u32(a_buf) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t w_lo0 = ((ptx::smem_u32(w_buf) & 0x3FFFFu) >> 4) | (1u << 16);
        ptx::mbar_wait(wfull_bar, 0, 600);
        for (int step = 1; step < T; ++step) {
            const int b = step & 1;
            const uint32_t par = (uint32_t)(((step - 1) >> 1) & 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                // own chunk last: it is the one this CTA sends itself, i.e. the one that leaves latest
                const int kc = (rank + 1 + i) & 7;
                ptx::mbar_wait(&afull_bar[b * 8 + kc], par, 610);
                ptx::tc_fence_after();
                const uint32_t a_lo = a_lo0 + (uint32_t)((b * kABytes + kc * kSliceBytes) >> 4);
                const uint32_t w_lo = w_lo0 + (uint32_t)((kc * kSliceBytes) >> 4);
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        ptx::umma_f16_lohi(tmem_base, a_lo + k * 2, desc_hi, w_lo + k * 2, desc_hi, p.idesc, (i | k) ? 1u : 0u);
                    if (i == 7) ptx::umma_commit(tfull_bar);
                }
                if (i == 0) LSTM_STAMP(step, 1);
                if (i == 7) LSTM_STAMP(step, 2);
            }
        }
    } else if (warp >= 4) {
        const int quarter = warp & 3;                 // TMEM lane quarter
        const int chalf = (warp - 4) >> 2;            // which 64 of the CTA's 128 gate columns (16 units)
        const int row = quarter * 32 + lane;          // crop within the cluster's block
        const int crop = crop0 + row;
        const bool valid = crop < p.B;
        const int ucol = rank * 32 + chalf * 16;      // first hidden unit of this thread
        const float* xp_base = p.xproj + (size_t)(valid ? crop : 0) * T * 2048 + dir * 1024 + ucol * 4;
        uint16_t* out_base = p.out + (size_t)(valid ? crop : 0) * T * p.pitch + dir * 256 + ucol;
        float cc[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) cc[u] = 0.f;
        float4 x[16];
        {
            const int t0 = dir == 0 ? 0 : T - 1;
            const float4* xp4 = reinterpret_cast<const float4*>(xp_base + (size_t)t0 * 2048);
#pragma unroll
            for (int u = 0; u < 16; ++u) x[u] = valid ? __ldg(xp4 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int step = 0; step < T; ++step) {
            const int t = dir == 0 ? step : T - 1 - step;
            if (step > 0) {
                ptx::mbar_wait(tfull_bar, (uint32_t)((step - 1) & 1), 700);
                ptx::tc_fence_after();
            }
            if (warp == 4) LSTM_STAMP(step, 3);       // accumulator complete
            uint32_t hp[8];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t r[32];
                if (step > 0) {
                    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(chalf * 64 + hf * 32), r);
                    ptx::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                float hv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 xv = x[hf * 8 + u];
                    const float gi = sigmoid_f(__uint_as_float(r[u * 4 + 0]) + xv.x);
                    const float gf = sigmoid_f(__uint_as_float(r[u * 4 + 1]) + xv.y);
                    const float gg = tanh_f(__uint_as_float(r[u * 4 + 2]) + xv.z);
                    const float go = sigmoid_f(__uint_as_float(r[u * 4 + 3]) + xv.w);
                    const float c = gf * cc[hf * 8 + u] + gi * gg;
                    cc[hf * 8 + u] = c;
                    hv[u] = go * tanh_f(c);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) hp[hf * 4 + q] = pack2h(hv[q * 2], hv[q * 2 + 1], p.is_f16);
            }
            if (warp == 4) LSTM_STAMP(step, 4);       // gate math done
            if (valid) {
                uint4* o = reinterpret_cast<uint4*>(out_base + (size_t)t * p.pitch);
                o[0] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                o[1] = make_uint4(hp[4], hp[5], hp[6], hp[7]);
            }
            if (step + 1 < T) {
                // h_t is in global memory (generic proxy); this CTA's producer reads it back through TMA (async proxy)
                ptx::fence_proxy_async_all();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(hlocal_bar);
                if (warp == 4) LSTM_STAMP(step, 5);   // h stored, producer signalled
                const int tn = dir == 0 ? step + 1 : T - 2 - step;
                const float4* xp4 = reinterpret_cast<const float4*>(xp_base + (size_t)tn * 2048);
#pragma unroll
                for (int u = 0; u < 16; ++u) x[u] = valid ? __ldg(xp4 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 128);
    }
    // no CTA leaves while a peer's multicast could still be writing into it
    ptx::cluster_arrive_release();
    ptx::cluster_wait_acquire();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

cudaError_t launch_lstm_tc(const float* xproj, const void* whh_perm, void* out, int B, int T, int is_f16,
                           cudaStream_t s, int split) {
    static EncodeTiledFn encode = nullptr;
    if (encode == nullptr) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return cudaErrorNotSupported;
        encode = reinterpret_cast<EncodeTiledFn>(fp);
    }
    if (B <= 0 || T <= 0) return cudaErrorInvalidValue;
    const CUtensorMapDataType dt = is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUtensorMap mw, mh, mhlo;
    const int pitch = split ? 1024 : 512;
    static int allow_mcast = -1;
    if (allow_mcast < 0) { const char* e = getenv("LOCR_LSTM_MCAST"); allow_mcast = e ? atoi(e) : 1; }
    if (allow_mcast && !split) {
        // push form: 32-unit chunks ([128 rows][64 B], 64-byte swizzle), one per CTA of the cluster
        {
            cuuint64_t dims[2] = {256, 2048};
            cuuint64_t strides[1] = {512};
            cuuint32_t box[2] = {32, 128};
            cuuint32_t estr[2] = {1, 1};
            if (encode(&mw, dt, 2, const_cast<void*>(whh_perm), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return cudaErrorInvalidValue;
        }
        {
            cuuint64_t dims[3] = {512, (cuuint64_t)T, (cuuint64_t)B};
            cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)T * pitch * 2};
            cuuint32_t box[3] = {32, 1, 128};
            cuuint32_t estr[3] = {1, 1, 1};
            if (encode(&mh, dt, 3, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return cudaErrorInvalidValue;
        }
        static bool attr_m = false;
        const size_t smem_m = 1024 + kWBytes + 2 * kABytes + 24 * 8;
        if (!attr_m) {
            cudaError_t e = cudaFuncSetAttribute(lstm_mcast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m);
            if (e != cudaSuccess) return e;
            attr_m = true;
        }
        LstmParams pm;
        pm.xproj = xproj; pm.out = (uint16_t*)out; pm.B = B; pm.T = T; pm.is_f16 = is_f16; pm.pitch = pitch;
        pm.idesc = ptx::make_idesc_f16(is_f16 ? 0 : 1, 128, 128);
        dim3 grid_m(kCluster * ((B + 127) / 128), 2);
        lstm_mcast_kernel<<<grid_m, kThreadsLstm, smem_m, s>>>(mw, mh, pm);
        return cudaGetLastError();
    }
    {
        cuuint64_t dims[2] = {256, 2048};
        cuuint64_t strides[1] = {512};
        cuuint32_t box[2] = {64, 128};
        cuuint32_t estr[2] = {1, 1};
        if (encode(&mw, dt, 2, const_cast<void*>(whh_perm), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    {
        // layer output [B][T][512] viewed as (column, t, crop); one box = 64 units of one time step for 128 crops
        cuuint64_t dims[3] = {512, (cuuint64_t)T, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)T * pitch * 2};
        cuuint32_t box[3] = {64, 1, 128};
        cuuint32_t estr[3] = {1, 1, 1};
        if (encode(&mh, dt, 3, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
        mhlo = mh;
        if (split && encode(&mhlo, dt, 3, (uint16_t*)out + 512, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    static bool attr = false;
    const size_t smem = 1024 + kWBytes + 2 * kABytes + 16 * 8;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(lstm_cluster_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(lstm_cluster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    LstmParams p;
    p.xproj = xproj; p.out = (uint16_t*)out; p.B = B; p.T = T; p.is_f16 = is_f16; p.pitch = pitch;
    p.idesc = ptx::make_idesc_f16(is_f16 ? 0 : 1, 128, 128);
    dim3 grid(kCluster * ((B + 127) / 128), 2);
    if (split) lstm_cluster_kernel<true><<<grid, kThreadsLstm, smem, s>>>(mw, mh, mhlo, p);
    else lstm_cluster_kernel<false><<<grid, kThreadsLstm, smem, s>>>(mw, mh, mhlo, p);
    return cudaGetLastError();
}

#if LOCR_LSTM_TRACE
extern "C" __attribute__((visibility("default"))) int locr_debug_lstm_trace(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_lstm_trace, sizeof(long long) * 64 * 8);
}
#endif

}  // namespace locr
