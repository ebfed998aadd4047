// Implicit-GEMM convolution for sm_100a.
//
//   GEMM view:  D[M = output pixels, N = Cout] = sum over (tap, channel chunk) of A[M, 64ch] * W[N, 64ch]^T
//   A operand : one TMA box per (tap, channel chunk): a bw x bh x bb patch of NHWC pixels (128 rows of the M
//               tile) shifted by the tap offset; out-of-bounds pixels are zero-filled by the TMA unit, which IS the
//               convolution's zero padding - no im2col buffer, no halo copies, no boundary branches.
//   B operand : K-major weight slab [n_tile][64] from a 2-D tensor map.
//   MMA       : tcgen05.mma cta_group::1 kind::f16, M=128, N=n_tile (16..256), K=16; fp32 accumulators in TMEM,
//               double-buffered so the epilogue of tile i overlaps the mainloop of tile i+1.
//   Epilogue  : tcgen05.ld -> +bias (+residual) -> ReLU -> 16-bit NHWC (or fp32) stores.
//   Roles     : warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue.
//   Grid      : persistent, one CTA per SM, static round-robin over (m_tile, n_tile) with n fastest so the
//               activation patch is reused out of L2 by consecutive CTAs.
//
// Replaces the cuDNN conv + separate BN/ReLU kernels behind nn.Conv2d/nn.BatchNorm2d/nn.ReLU in
// reference ocr/modules/vgg_bn.py:23-55, ocr/model.py:21-31, ocr/modules/resnet50v1.py:55-82,
// ocr/modules/TPS_STN.py:38-58 and the nn.Linear GEMMs of ocr/modules/biLSTM.py:18-20.
#include "conv_tc.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "ptx.cuh"

#include <atomic>

namespace locr {

namespace {

// -DLOCR_CONV_EXPERIMENTS=1 (tools only): LOCR_CONV_DBG bit 1 skips the MMAs, 2 the A loads, 4 the B loads, 8 the stores
#ifndef LOCR_CONV_EXPERIMENTS
#define LOCR_CONV_EXPERIMENTS 0
#endif
// LOCR_CONV_DBG bit 32 (experiments build): block 0 stamps clock64 at the hand-over points of its producer (role 0),
// MMA-issuer (1) and first epilogue warp (2) into g_conv_trace; read with conv_tc_trace_read (tools/conv_trace.py).
#if LOCR_CONV_EXPERIMENTS
__device__ unsigned long long g_conv_trace[3][8192];
#define LOCR_TRACE(role, ev)                                                                          \
    do {                                                                                              \
        if ((p.dbg & 32) && blockIdx.x == 0 && lane == 0 && trace_n < 8192) {                         \
            g_conv_trace[role][trace_n] = ((unsigned long long)clock64() << 4) | (unsigned)(ev);      \
            ++trace_n;                                                                                \
        }                                                                                             \
    } while (0)
#else
#define LOCR_TRACE(role, ev) do { } while (0)
#endif
constexpr int kMaxStages = 8;
constexpr int kThreads = 384;
constexpr int kTileM = 128;
constexpr int kEpiRes = 64, kEpiPool = 128, kEpiSkip = 256;

struct ConvParams {
    int B, OH, OW, Cout;
    int bw, bh, bb;      // 128-pixel half box
    int halves;          // 1 or 2 half boxes per tile (M = 128 or 256), stacked along h (split_b = 0), b (1) or w (2)
    int split_b;
    int tiles_w, tiles_h, tiles_n, num_tiles;
    int num_pair_tiles;   // CTA pairs: ceil(m-tiles / 2) * tiles_n
    int n_tile, n_tile_alloc, tmem_cols;
    int cin_chunks, KW, dil_h, dil_w, pad_h, pad_w, stride2;
    int num_kblocks, stages;
    uint32_t a_stage_bytes, b_stage_bytes;
    uint32_t idesc;
    void* y;
    long y_pitch;
    int out_fp32;
    const void* res;
    long res_pitch;
    long res_lo_off;    // split-precision residual: offset (elements) of its lo halves, 0 = plain residual
    const float* bias;
    int relu, is_f16;
    // TMA-store epilogue: accumulators -> registers -> swizzled staging tile in smem -> one bulk tensor store
    int tma_store;      // 0 = direct per-thread global stores (unaligned fp32 outputs)
    int stage_cols;     // output columns per staging chunk (row bytes stage_rb = stage_cols * elem size: 32/64/128)
    int stage_rb;
    int n_chunks;       // n_tile / stage_cols
    int cin_wrap;       // channel coordinate wraps at this value (split-precision inputs), huge when unused
    int split_out;      // write hi / lo halves (split-precision outputs)
    int halo;           // 3x3 / 64-channel layers: one haloed input patch per tile, the nine taps are descriptor offsets
    int hstream;        // CTA pairs only: haloed patches (one per 64-channel chunk, ring of two) + a streamed weight ring
    int a_slots;        // A slots in shared memory in front of the B ring (= stages; 2 patches with hstream)
    int halo_pool;      // halo tiles whose only output is the 2x2 max-pooled tensor: pooling by warp shuffles
    int pool;           // fused 2x2/2 max-pool of the activated output (TMA-store epilogue only)
    int skip_full;      // pooled output only
    const float* tail_w;   // fused 1x1 tail (see ConvCall)
    float* tail_out;
    // first layer from the uint8 image (ConvCall::first_u8)
    const uint8_t* first_u8;
    int img_h, img_w;
    long img_row_stride, img_stride;
    int dbg;            // experiments only (LOCR_CONV_DBG): 1 = skip MMAs, 2 = skip A loads, 4 = skip B loads, 8 = skip epilogue math
    // Split-K (ConvCall::ksplit, generic kernel only): the input-channel chunks of every tap are cut into ksplit slices and
    // the tile walk covers ksplit * base_tiles tiles; tile q works on slice q / base_tiles of tile q % base_tiles
    // (cin_chunks and num_kblocks are those of ONE slice) and stores its fp32 partial sums at output channel
    // slice * ks_out_stride + n of the workspace.  cin_total = elements per tap of a weight row.
    int ksplit, base_tiles, cin_total, ks_out_stride;
};

struct TileCoord {
    int n0, ow0, oh0, b0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile) {
    TileCoord t;
    int n_idx = tile % p.tiles_n;
    int m_idx = tile / p.tiles_n;
    int tw = m_idx % p.tiles_w;
    int rest = m_idx / p.tiles_w;
    int th = rest % p.tiles_h;
    int tb = rest / p.tiles_h;
    t.n0 = n_idx * p.n_tile;
    t.ow0 = tw * p.bw * (p.split_b == 2 ? p.halves : 1);
    t.oh0 = th * p.bh * (p.split_b == 0 ? p.halves : 1);
    t.b0 = tb * p.bb * (p.split_b == 1 ? p.halves : 1);
    return t;
}

// Round-to-nearest conversion of two fp32 values to a packed 16-bit pair (a in the low half).  Saturating
// (F2FP.SATFINITE, same cost as the plain conversion): an activation beyond the fp16 range becomes +-65504 instead of
// an infinity that would turn into NaN in the next layer.
__device__ __forceinline__ uint32_t pack2(float a, float b, int is_f16) {
    uint32_t r;
    if (is_f16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    else asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// element-wise max of two packed 16-bit pairs
__device__ __forceinline__ uint32_t max2_16(uint32_t a, uint32_t b, int is_f16) {
    if (is_f16) {
        __half2 m = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
        return *reinterpret_cast<uint32_t*>(&m);
    } else {
        __nv_bfloat162 m = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
        return *reinterpret_cast<uint32_t*>(&m);
    }
}
__device__ __forceinline__ float2 unpack2(uint32_t u, int is_f16) {
    if (is_f16) {
        return __half22float2(*reinterpret_cast<__half2*>(&u));
    } else {
        return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
    }
}

// EPI = 0: generic epilogue (every option is a run-time switch).  Otherwise the 16-bit TMA-store epilogue with its options
// fixed at compile time, which matters for the layers whose tiles are epilogue-bound: EPI & 63 = columns per warp (32 / 16:
// 64- / 32-column staging chunks), kEpiRes = residual add, kEpiPool = fused 2x2 max-pool, kEpiSkip = pooled output only;
// never split-precision, fp32, fused-tail or halo-pool (those stay on the generic path).
constexpr int kFirstThreads = kThreads + 128;   // + four producer warps that build the first layer's A operand

// CTA2: two CTAs of a cluster (one TPC) share every MMA (tcgen05 cta_group::2, M = 256 = 128 rows per CTA): each CTA
// stages the A boxes of its own m-tile and HALF of the weight slab's N rows, the leader CTA issues the instruction for
// both, so per CTA the tensor core reads A + B/2 from shared memory instead of A + B and the weight fills halve.  Used
// for the N <= 128 layers, which are bound by shared-memory operand traffic.  A pair walks "pair tiles": the same n-tile
// over two consecutive m-tiles (rank 0 / rank 1).
template <int SWZ, int HALVES, int EPI, bool FIRST = false, bool CTA2 = false>
__global__ void __launch_bounds__(FIRST ? kFirstThreads : kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_p,
               const ConvParams p) {
    constexpr int BLOCK_K = SWZ / 2;      // 16-bit elements per swizzled row
    constexpr int MMAS_PER_STAGE = BLOCK_K / 16;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)p.a_slots * p.a_stage_bytes;
    // all stage sizes are multiples of 1024 bytes, so the epilogue staging tiles that follow stay 1024-aligned
    uint8_t* staging = smem_b + (size_t)p.stages * p.b_stage_bytes;                            // 2 x [128][stage_rb]
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * 128 * 128 + (p.pool ? 2 * 32 * 128 : 0));
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kMaxStages;
    uint64_t* tfull_bar = bars + 2 * kMaxStages;
    uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
    float* bias_s = reinterpret_cast<float*>(bars + 2 * kMaxStages + 6);                       // [256]
    uint8_t* pool_staging = staging + 2 * 128 * 128;                                           // 2 x [32][stage_rb]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
#if LOCR_CONV_EXPERIMENTS
    int trace_n = 0;
#endif
    // tile walk: CTAs (or CTA pairs) take tiles (pair tiles) round-robin
    const uint32_t rank = CTA2 ? ptx::cluster_ctarank() : 0u;
    const int t_begin = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int t_step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int t_end = CTA2 ? p.num_pair_tiles : p.num_tiles;
    // pair tile q -> this CTA's tile: same n-tile, m-tile 2 * (q / tiles_n) + rank (may lie past the end: its loads
    // are zero-filled and its stores clipped by the TMA unit)
    auto tile_of = [&](int q) {
        if (p.ksplit > 1) q %= p.base_tiles;       // split-K: the same tiles once per K slice
        return CTA2 ? ((2 * (q / p.tiles_n) + (int)rank) * p.tiles_n + q % p.tiles_n) : q;
    };

    if (warp == 0 && lane == 0) {
        if (!FIRST) ptx::tma_prefetch_desc(&tmap_x);
        ptx::tma_prefetch_desc(&tmap_w);
        if (p.tma_store && !p.skip_full && p.tail_out == nullptr) ptx::tma_prefetch_desc(&tmap_y);
        if (p.pool) ptx::tma_prefetch_desc(&tmap_p);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            ptx::mbar_init(&full_bar[s], FIRST ? 4 : 1);     // FIRST: one arrival per producer warp
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tfull_bar[a], 1);
            ptx::mbar_init(&tempty_bar[a], CTA2 ? 16 : 8);   // CTA2: the leader also waits for the peer's epilogue warps
        }
        if (p.halo || FIRST) ptx::mbar_init(&full_bar[kMaxStages - 1], 1);   // resident weights landed
        if (CTA2 && p.hstream)                                                // patch ring: slots kMaxStages - 2, - 1
            for (int s = kMaxStages - 2; s < kMaxStages; ++s) {
                ptx::mbar_init(&full_bar[s], 1);
                ptx::mbar_init(&empty_bar[s], 1);
            }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        if (CTA2) {
            ptx::tmem_alloc_2sm(tmem_ptr_smem, (uint32_t)p.tmem_cols);
            ptx::tmem_relinquish_2sm();
        } else {
            ptx::tmem_alloc(tmem_ptr_smem, (uint32_t)p.tmem_cols);
            ptx::tmem_relinquish();
        }
    }
    uint16_t* lut = reinterpret_cast<uint16_t*>(bias_s + 256);                                   // FIRST: [3][256]
    if (FIRST) {
        // normalizeMeanVariance as a table: the 16-bit value of (v - mean[c]) / std[c] for every byte v (BGR pixels meet
        // the RGB constants exactly as in the reference, imgproc.py:19-25)
        const float mean[3] = {(float)(0.485 * 255.0), (float)(0.456 * 255.0), (float)(0.406 * 255.0)};
        const float stdv[3] = {(float)(0.229 * 255.0), (float)(0.224 * 255.0), (float)(0.225 * 255.0)};
        for (int i = threadIdx.x; i < 768; i += blockDim.x) {
            const int c = i >> 8;
            const float f = ((float)(i & 255) - mean[c]) / stdv[c];
            lut[i] = (uint16_t)(pack2(f, 0.f, p.is_f16) & 0xffffu);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (CTA2) {      // the peer's barriers exist before anything is signalled across the pair
        ptx::cluster_arrive_release();
        ptx::cluster_wait_acquire();
    }
    const uint32_t tmem_base = *tmem_ptr_smem;
    // tempty hand-back of an epilogue warp: in a pair the accumulators of BOTH CTAs are rewritten by the leader's MMAs
    const uint32_t tempty_remote0 = CTA2 ? ptx::mapa(ptx::smem_u32(&tempty_bar[0]), 0u) : 0u;
    auto tempty_arrive = [&](int a) {
        if (CTA2) ptx::mbar_arrive_remote_release(tempty_remote0 + 8u * (uint32_t)a);
        else ptx::mbar_arrive(&tempty_bar[a]);
    };

    if (FIRST && warp >= kThreads / 32) {
        // ------------------------------------------------------------ first-layer A producer (four warps)
        // Thread pr owns row pr of both 128-pixel halves of every tile: it gathers the 3x3x3 neighbourhood of its pixel
        // and writes one 64-byte K-major row (k = (ky * 3 + kx) * 3 + c, 27 used) in the SWIZZLE_64B layout the MMA's
        // shared-memory descriptor expects: 16-byte piece q of row r lives at r * 64 + ((q ^ ((r >> 1) & 3)) << 4).
        const int pr = threadIdx.x - kThreads;
        const int rw = pr % p.bw, rh = (pr / p.bw) % p.bh, rb = pr / (p.bw * p.bh);
        const uint32_t full0 = ptx::smem_u32(full_bar), empty0 = ptx::smem_u32(empty_bar);
        const uint32_t row_off = (uint32_t)pr * 64u;
        const uint32_t sw = ((uint32_t)pr >> 1) & 3u;
        int stage = 0;
        uint32_t phase = 1;
        for (int tile = t_begin; tile < t_end; tile += t_step) {
            const TileCoord t = decode_tile(p, tile_of(tile));
            ptx::mbar_wait_a(empty0 + 8u * stage, phase, 100);
            uint8_t* a_stage = smem_a + (size_t)stage * p.a_stage_bytes;
#pragma unroll
            for (int hf = 0; hf < HALVES; ++hf) {
                const int ow = t.ow0 + (p.split_b == 2 ? hf * p.bw : 0) + rw;
                const int oh = t.oh0 + (p.split_b == 0 ? hf * p.bh : 0) + rh;
                const int b = t.b0 + (p.split_b == 1 ? hf * p.bb : 0) + rb;
                uint32_t h[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) h[i] = 0u;
                if (ow < p.OW && oh < p.OH && b < p.B) {
                    const uint8_t* img = p.first_u8 + (long)b * p.img_stride;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int y = oh + ky - 1;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const int x = ow + kx - 1;
                            uint32_t v0 = 0u, v1 = 0u, v2 = 0u;                      // conv padding: zero
                            if (y >= 0 && y < p.OH && x >= 0 && x < p.OW) {
                                uint32_t c0 = 0u, c1 = 0u, c2 = 0u;                  // canvas padding: raw 0
                                if (y < p.img_h && x < p.img_w) {
                                    const uint8_t* q = img + (long)y * p.img_row_stride + x * 3;
                                    c0 = __ldg(q); c1 = __ldg(q + 1); c2 = __ldg(q + 2);
                                }
                                v0 = lut[c0]; v1 = lut[256 + c1]; v2 = lut[512 + c2];
                            }
                            const int k = (ky * 3 + kx) * 3;                          // compile-time after unrolling
                            h[(k + 0) >> 1] |= v0 << (((k + 0) & 1) * 16);
                            h[(k + 1) >> 1] |= v1 << (((k + 1) & 1) * 16);
                            h[(k + 2) >> 1] |= v2 << (((k + 2) & 1) * 16);
                        }
                    }
                }
                uint8_t* row = a_stage + (size_t)hf * (kTileM * 64) + row_off;
#pragma unroll
                for (uint32_t q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(row + ((q ^ sw) << 4)) = make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
            }
            ptx::fence_proxy_async();          // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&full_bar[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
    } else if (FIRST && warp == 0) {
        // resident weights: one [64 x 32] K-major slab, once
        if (ptx::elect_one()) {
            const uint32_t full0 = ptx::smem_u32(full_bar);
            ptx::mbar_arrive_expect_tx_a(full0 + 8u * (kMaxStages - 1), (uint32_t)(p.n_tile * SWZ));
            ptx::tma_load_2d_a(ptx::smem_u32(smem_b), &tmap_w, full0 + 8u * (kMaxStages - 1), 0, 0);
        }
    } else if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        // The WHOLE warp runs the loop on warp-uniform values and one elected lane issues the TMA instructions.  With
        // the loop inside `if (lane == 0)` ptxas cannot prove uniformity: every UTMALDG / UTCHMMA then gets wrapped in
        // a per-thread election loop with R2UR moves (~10 extra instructions per issue), which bounded the k-block
        // rate of every layer with N <= 128.
        {
            // CTA2: the leader's barrier counts the boxes of both CTAs (each: its A boxes + half of the weight slab)
            uint32_t tx_bytes = CTA2 ? 2u * ((uint32_t)(p.halves * kTileM * SWZ) + (uint32_t)(p.n_tile / 2 * SWZ))
                                     : (uint32_t)(p.halves * kTileM * SWZ) + (uint32_t)(p.n_tile * SWZ);
            if (LOCR_CONV_EXPERIMENTS && (p.dbg & 2)) tx_bytes -= (uint32_t)(p.halves * kTileM * SWZ);
            if (LOCR_CONV_EXPERIMENTS && (p.dbg & 4)) tx_bytes -= (uint32_t)(p.n_tile * SWZ);
            const uint32_t a0 = ptx::smem_u32(smem_a), b0s = ptx::smem_u32(smem_b);
            const uint32_t full0 = ptx::smem_u32(full_bar), empty0 = ptx::smem_u32(empty_bar);
            const uint32_t a_end = a0 + (uint32_t)p.stages * p.a_stage_bytes;
            uint32_t a_s = a0, b_s = b0s, full_s = full0, empty_s = empty0, phase = 1;
            const int KH = p.num_kblocks / (p.KW * p.cin_chunks);
            const uint32_t full_lead0 = CTA2 ? ptx::mapa(full0, 0u) : 0u;     // the leader's full barriers (shared::cluster)
            const int n_half = CTA2 ? (int)rank * (p.n_tile / 2) : 0;
            if (CTA2 && p.hstream) {
                // weight ring only (the patches come from warp 3): per tile, chunk-major, one [n_tile / 2 x 64] slab per
                // (chunk, tap) and CTA; the leader's barrier counts the halves of both CTAs
                const uint32_t b_bytes = 2u * (uint32_t)(p.n_tile / 2 * SWZ);
                const uint32_t b_end = b0s + (uint32_t)p.stages * p.b_stage_bytes;
                const int cin = p.cin_chunks * BLOCK_K;
                for (int tile = t_begin; tile < t_end; tile += t_step) {
                    const TileCoord t = decode_tile(p, tile_of(tile));
                    for (int cc = 0; cc < p.cin_chunks; ++cc) {
                        for (int tp = 0; tp < 9; ++tp) {
                            ptx::mbar_wait_a(empty_s, phase, 120);
                            if (ptx::elect_one()) {
                                const uint32_t lead = full_lead0 + (full_s - full0);
                                if (rank == 0) ptx::mbar_arrive_expect_tx_a(full_s, b_bytes);
                                ptx::tma_load_2d_2sm(b_s, &tmap_w, lead, tp * cin + cc * BLOCK_K, t.n0 + n_half);
                            }
                            b_s += p.b_stage_bytes; full_s += 8; empty_s += 8;
                            if (b_s == b_end) {
                                b_s = b0s; full_s = full0; empty_s = empty0;
                                phase ^= 1u;
                            }
                        }
                    }
                }
            } else if (p.halo) {
                // resident weights: nine [n_tile x Cin] tap slabs (Cin = one swizzled row), once; then ONE haloed patch
                // (18 rows x 24 pixels) per tile
                if (ptx::elect_one()) {
                    const uint32_t slab = (uint32_t)(p.n_tile * SWZ);
                    ptx::mbar_arrive_expect_tx_a(full0 + 8u * (kMaxStages - 1), 9u * slab);
                    for (int tp = 0; tp < 9; ++tp)
                        ptx::tma_load_2d_a(b0s + (uint32_t)tp * slab, &tmap_w, full0 + 8u * (kMaxStages - 1), tp * BLOCK_K, 0);
                }
                for (int tile = t_begin; tile < t_end; tile += t_step) {
                    const TileCoord t = decode_tile(p, tile_of(tile));
                    LOCR_TRACE(0, 0);
                    ptx::mbar_wait_a(empty_s, phase, 110);
                    LOCR_TRACE(0, 1);
                    if (ptx::elect_one()) {
                        ptx::mbar_arrive_expect_tx_a(full_s, p.a_stage_bytes);
                        ptx::tma_load_5d_a(a_s, &tmap_x, full_s, 0, t.ow0 - 1, 0, t.oh0 - 1, t.b0);
                    }
                    LOCR_TRACE(0, 2);
                    a_s += p.a_stage_bytes; full_s += 8; empty_s += 8;
                    if (a_s == a_end) {
                        a_s = a0; full_s = full0; empty_s = empty0;
                        phase ^= 1u;
                    }
                }
            } else
            for (int tile = t_begin; tile < t_end; tile += t_step) {
                const TileCoord t = decode_tile(p, tile_of(tile));
                LOCR_TRACE(0, 0);
                int kcoord = 0;
                const int ks_c0 = p.ksplit > 1 ? (tile / p.base_tiles) * p.cin_chunks * BLOCK_K : 0;   // split-K: first channel of the slice
                for (int kh = 0; kh < KH; ++kh) {
                    const int c2 = p.stride2 ? kh : 0;
                    const int c3 = p.stride2 ? t.oh0 : t.oh0 + kh * p.dil_h - p.pad_h;
                    int iw0 = t.ow0 - p.pad_w;
                    for (int kw = 0; kw < p.KW; ++kw, iw0 += p.dil_w) {
                        for (int cc = 0; cc < p.cin_chunks; ++cc, kcoord += BLOCK_K) {
                            ptx::mbar_wait_a(empty_s, phase, 100);
                            LOCR_TRACE(0, 1);
                            int cch = cc * BLOCK_K;
                            if (cch >= p.cin_wrap) cch -= p.cin_wrap;   // [hi | lo | hi] of a split-precision input
                            int kc = kcoord;
                            if (p.ksplit > 1) {
                                cch += ks_c0;
                                kc = (kh * p.KW + kw) * p.cin_total + cch;
                            }
                            if (ptx::elect_one()) {
                                if (CTA2) {
                                    const uint32_t lead = full_lead0 + (full_s - full0);
                                    if (rank == 0) ptx::mbar_arrive_expect_tx_a(full_s, tx_bytes);
                                    ptx::tma_load_5d_2sm(a_s, &tmap_x, lead, cch, iw0, c2, c3, t.b0);
                                    ptx::tma_load_2d_2sm(b_s, &tmap_w, lead, kc, t.n0 + n_half);
                                } else {
                                if (LOCR_CONV_EXPERIMENTS && tx_bytes == 0) ptx::mbar_arrive(&full_bar[(full_s - full0) >> 3]);
                                else ptx::mbar_arrive_expect_tx_a(full_s, tx_bytes);
                                if (!(LOCR_CONV_EXPERIMENTS && (p.dbg & 2)))
                                    ptx::tma_load_5d_a(a_s, &tmap_x, full_s, cch, iw0, c2, c3, t.b0);
                                if (!(LOCR_CONV_EXPERIMENTS && (p.dbg & 4)))
                                    ptx::tma_load_2d_a(b_s, &tmap_w, full_s, kc, t.n0);
                                }
                            }
                            LOCR_TRACE(0, 2);
                            a_s += p.a_stage_bytes; b_s += p.b_stage_bytes; full_s += 8; empty_s += 8;
                            if (a_s == a_end) {
                                a_s = a0; b_s = b0s; full_s = full0; empty_s = empty0;
                                phase ^= 1u;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && !(CTA2 && rank != 0)) {
        // ------------------------------------------------------------ MMA issuer (whole warp waits, one elected lane issues;
        // in a CTA pair only the leader's)
        {
            // descriptors as 32-bit running values: hi word constant, lo word = (addr >> 4) | LBO field
            const uint32_t desc_hi = (uint32_t)(ptx::make_kmajor_desc(0, SWZ) >> 32);
            const uint32_t lo_flag = 1u << 16;
            const uint32_t a_lo0 = ((ptx::smem_u32(smem_a) & 0x3FFFFu) >> 4) | lo_flag;
            const uint32_t b_lo0 = ((ptx::smem_u32(smem_b) & 0x3FFFFu) >> 4) | lo_flag;
            const uint32_t a_step = p.a_stage_bytes >> 4, b_step = p.b_stage_bytes >> 4;
            const uint32_t full0 = ptx::smem_u32(full_bar), empty0 = ptx::smem_u32(empty_bar);
            const uint32_t full_end = full0 + 8u * (uint32_t)p.stages;
            uint32_t a_lo = a_lo0, b_lo = b_lo0, full_s = full0, empty_s = empty0, phase = 0;
            const uint32_t tfull0 = ptx::smem_u32(tfull_bar), tempty0 = ptx::smem_u32(tempty_bar);
            uint32_t acc = 0, acc_phase = 1;
            const uint32_t acc_cols = (uint32_t)(HALVES * p.n_tile_alloc);
            const uint32_t idesc = p.idesc;
            constexpr uint32_t kHalfStep = (uint32_t)(kTileM * SWZ) >> 4;
            if (FIRST) {
                // one k-block per tile (K = 32 = two K16 MMAs per half) against the resident weight slab
                ptx::mbar_wait_a(full0 + 8u * (kMaxStages - 1), 0, 310);
                for (int tile = t_begin; tile < t_end; tile += t_step) {
                    ptx::mbar_wait_a(tempty0 + acc * 8u, acc_phase, 200);
                    ptx::mbar_wait_a(full_s, phase, 300);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * acc_cols;
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int k = 0; k < MMAS_PER_STAGE; ++k) {
#pragma unroll
                            for (int hf = 0; hf < HALVES; ++hf)
                                ptx::umma_f16_lohi(d_tmem + (uint32_t)(hf * p.n_tile_alloc), a_lo + hf * kHalfStep + k * 2,
                                                   desc_hi, b_lo0 + k * 2, desc_hi, idesc, k ? 1u : 0u);
                        }
                        ptx::umma_commit_a(empty_s);
                        ptx::umma_commit_a(tfull0 + acc * 8u);
                    }
                    a_lo += a_step; full_s += 8; empty_s += 8;
                    if (full_s == full_end) {
                        a_lo = a_lo0; full_s = full0; empty_s = empty0;
                        phase ^= 1u;
                    }
                    acc ^= 1u;
                    if (acc == 0) acc_phase ^= 1u;
                }
            } else if (CTA2 && p.hstream) {
                // Patches [18 rows][24 pixels][128 B] per 64-channel chunk (ring of two, barrier slots kMaxStages - 2 / - 1)
                // against the streamed weight ring: tap (kh, kw) of half hf reads the patch from pixel (kh, kw + 8 hf) on,
                // exactly like the resident-weight halo form below; M = 256 over the pair, N = 128.
                const uint32_t hi_patch = ((24u * (uint32_t)SWZ) >> 4) | (1u << 14) | (2u << 29);
                const uint32_t pfull0 = full0 + 8u * (kMaxStages - 2), pempty0 = empty0 + 8u * (kMaxStages - 2);
                uint32_t pslot = 0, pphase = 0;
                for (int tile = t_begin; tile < t_end; tile += t_step) {
                    ptx::mbar_wait_a(tempty0 + acc * 8u, acc_phase, 200);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * acc_cols;
                    for (int cc = 0; cc < p.cin_chunks; ++cc) {
                        ptx::mbar_wait_a(pfull0 + 8u * pslot, pphase, 320);
                        const uint32_t a_patch = a_lo0 + pslot * a_step;
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                            for (int kw = 0; kw < 3; ++kw) {
                                ptx::mbar_wait_a(full_s, phase, 300);
                                ptx::tc_fence_after();
                                if (ptx::elect_one()) {
#pragma unroll
                                    for (int k = 0; k < MMAS_PER_STAGE; ++k) {
#pragma unroll
                                        for (int hf = 0; hf < 2; ++hf) {
                                            const uint32_t a_tap = a_patch + (uint32_t)(((kh * 24 + kw + 8 * hf) * SWZ) >> 4) + k * 2;
                                            ptx::umma_f16_lohi_2sm(d_tmem + (uint32_t)(hf * p.n_tile_alloc), a_tap, hi_patch,
                                                                   b_lo + k * 2, desc_hi, idesc, (cc | kh | kw | k) ? 1u : 0u);
                                        }
                                    }
                                    ptx::umma_commit_2sm(empty_s);
                                }
                                b_lo += b_step; full_s += 8; empty_s += 8;
                                if (full_s == full_end) {
                                    b_lo = b_lo0; full_s = full0; empty_s = empty0;
                                    phase ^= 1u;
                                }
                            }
                        }
                        if (ptx::elect_one()) ptx::umma_commit_2sm(pempty0 + 8u * pslot);   // patch slot free in both CTAs
                        pslot ^= 1u;
                        if (pslot == 0) pphase ^= 1u;
                    }
                    if (ptx::elect_one()) ptx::umma_commit_2sm(tfull0 + acc * 8u);
                    acc ^= 1u;
                    if (acc == 0) acc_phase ^= 1u;
                }
            } else if (p.halo) {
                // A patch in smem: [18 rows][24 pixels][128 B], 128B-swizzled by the TMA unit.  Tap (kh, kw) of half hf is
                // the same patch read from pixel (kh, kw + 8 hf) on: start address + (kh * 24 + kw + 8 hf) * 128 B,
                // 8-pixel core groups one image row (3072 B) apart.  The tensor core applies the 128B swizzle to the
                // absolute shared-memory address (measured: the descriptor's base-offset field must stay 0), which is
                // exactly where the TMA unit put the data.
                ptx::mbar_wait_a(full0 + 8u * (kMaxStages - 1), 0, 310);
                // (SWZ = 64: 32-channel layers, 64-byte pixels - the same construction on 64-byte rows)
                const uint32_t hi_common = ((24u * (uint32_t)SWZ) >> 4) | (1u << 14) | ((SWZ == 128 ? 2u : 4u) << 29);
                const uint32_t slab16 = (uint32_t)(p.n_tile * SWZ) >> 4;
                for (int tile = t_begin; tile < t_end; tile += t_step) {
                    LOCR_TRACE(1, 0);
                    ptx::mbar_wait_a(tempty0 + acc * 8u, acc_phase, 200);
                    LOCR_TRACE(1, 1);
                    ptx::mbar_wait_a(full_s, phase, 300);
                    LOCR_TRACE(1, 2);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * acc_cols;
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                            for (int kw = 0; kw < 3; ++kw) {
                                const uint32_t a_hi = hi_common;
                                const uint32_t b_tap = b_lo0 + (uint32_t)(kh * 3 + kw) * slab16;
#pragma unroll
                                for (int k = 0; k < MMAS_PER_STAGE; ++k) {
#pragma unroll
                                    for (int hf = 0; hf < 2; ++hf) {
                                        const uint32_t a_tap = a_lo + (uint32_t)(((kh * 24 + kw + 8 * hf) * SWZ) >> 4) + k * 2;
                                        ptx::umma_f16_lohi(d_tmem + (uint32_t)(hf * p.n_tile_alloc), a_tap, a_hi, b_tap + k * 2,
                                                           desc_hi, idesc, (kh | kw | k) ? 1u : 0u);
                                    }
                                }
                            }
                        }
                        ptx::umma_commit_a(empty_s);
                        ptx::umma_commit_a(tfull0 + acc * 8u);
                    }
                    LOCR_TRACE(1, 4);
                    a_lo += a_step; full_s += 8; empty_s += 8;
                    if (full_s == full_end) {
                        a_lo = a_lo0; full_s = full0; empty_s = empty0;
                        phase ^= 1u;
                    }
                    acc ^= 1u;
                    if (acc == 0) acc_phase ^= 1u;
                }
            } else
            for (int tile = t_begin; tile < t_end; tile += t_step) {
                LOCR_TRACE(1, 0);
                ptx::mbar_wait_a(tempty0 + acc * 8u, acc_phase, 200);
                LOCR_TRACE(1, 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * acc_cols;
                uint32_t accum = 0;
                for (int kb = 0; kb < p.num_kblocks; ++kb) {
                    ptx::mbar_wait_a(full_s, phase, 300);
                    LOCR_TRACE(1, 2);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int k = 0; k < MMAS_PER_STAGE; ++k) {
#pragma unroll
                            for (int hf = 0; hf < HALVES; ++hf) {
                                // independent accumulators (M = 256 tiles) alternate, hiding the MMA pipeline latency
                                if (CTA2)
                                ptx::umma_f16_lohi_2sm(d_tmem + (uint32_t)(hf * p.n_tile_alloc), a_lo + hf * kHalfStep + k * 2,
                                                       desc_hi, b_lo + k * 2, desc_hi, idesc, (k == 0) ? accum : 1u);
                                else if (!(LOCR_CONV_EXPERIMENTS && (p.dbg & 1)))
                                ptx::umma_f16_lohi(d_tmem + (uint32_t)(hf * p.n_tile_alloc), a_lo + hf * kHalfStep + k * 2,
                                                   desc_hi, b_lo + k * 2, desc_hi, idesc, (k == 0) ? accum : 1u);
                            }
                        }
                        if (CTA2) ptx::umma_commit_2sm(empty_s);      // frees the slot in both CTAs
                        else if (LOCR_CONV_EXPERIMENTS && (p.dbg & 64))   // skeleton runs: plain arrive instead of the commit
                            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_s) : "memory");
                        else
                        ptx::umma_commit_a(empty_s);  // frees the smem slot once these MMAs retire
                    }
                    LOCR_TRACE(1, 3);
                    accum = 1;
                    a_lo += a_step; b_lo += b_step; full_s += 8; empty_s += 8;
                    if (full_s == full_end) {
                        a_lo = a_lo0; b_lo = b_lo0; full_s = full0; empty_s = empty0;
                        phase ^= 1u;
                    }
                }
                if (ptx::elect_one()) {                                       // accumulators complete -> epilogue(s)
                    if (CTA2) ptx::umma_commit_2sm(tfull0 + acc * 8u);
                    else ptx::umma_commit_a(tfull0 + acc * 8u);
                }
                LOCR_TRACE(1, 4);
                acc ^= 1u;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (CTA2 && warp == 3 && p.hstream) {
        // ------------------------------------------------------------ patch producer (hstream): one haloed patch per tile
        // and 64-channel chunk into a ring of two; the leader's barrier counts the patches of both CTAs
        const uint32_t a0 = ptx::smem_u32(smem_a);
        const uint32_t pfull0 = ptx::smem_u32(full_bar) + 8u * (kMaxStages - 2);
        const uint32_t pempty0 = ptx::smem_u32(empty_bar) + 8u * (kMaxStages - 2);
        const uint32_t pfull_lead0 = ptx::mapa(pfull0, 0u);
        uint32_t pslot = 0, pphase = 1;
        for (int tile = t_begin; tile < t_end; tile += t_step) {
            const TileCoord t = decode_tile(p, tile_of(tile));
            for (int cc = 0; cc < p.cin_chunks; ++cc) {
                ptx::mbar_wait_a(pempty0 + 8u * pslot, pphase, 130);
                if (ptx::elect_one()) {
                    if (rank == 0) ptx::mbar_arrive_expect_tx_a(pfull0 + 8u * pslot, 2u * p.a_stage_bytes);
                    ptx::tma_load_5d_2sm(a0 + pslot * p.a_stage_bytes, &tmap_x, pfull_lead0 + 8u * pslot, cc * BLOCK_K,
                                         t.ow0 - 1, 0, t.oh0 - 1, t.b0);
                }
                pslot ^= 1u;
                if (pslot == 0) pphase ^= 1u;
            }
        }
    } else if (warp >= 4 && warp < kThreads / 32) {
        // ------------------------------------------------------------ epilogue (8 warps: lane quarter x column half)
        const int ew = warp & 3;
        const int half = (warp - 4) >> 2;
        const int row = ew * 32 + lane;
        const int rw = row % p.bw;
        const int rh = (row / p.bw) % p.bh;
        const int rb = row / (p.bw * p.bh);
        int acc = 0;
        uint32_t acc_phase = 0;
        const int chunks = (p.n_tile + 31) / 32;
        if (EPI == 0 && p.tail_out != nullptr) {
            // ---- conv (Cout = 16) + ReLU + 1x1 (16->16) + ReLU + 1x1 (16->2), all in the registers of the pixel's thread.
            // With M = 256 tiles the two warp groups take one 128-pixel half each.
            float* tw = reinterpret_cast<float*>(staging);          // [256 W6 | 16 b6 | 32 W8 | 2 b8 | 16 conv bias]
            const int etid = threadIdx.x - 128;
            for (int i = etid; i < 306; i += 256) tw[i] = __ldg(&p.tail_w[i]);
            if (etid < 16) tw[306 + etid] = __ldg(&p.bias[etid]);
            ptx::named_bar_sync(1, 256);
            for (int tile = t_begin; tile < t_end; tile += t_step) {
                const TileCoord t = decode_tile(p, tile_of(tile));
                ptx::mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
                ptx::tc_fence_after();
                const int hf = half;
                const bool work = hf < p.halves;
                uint32_t r[16];
                if (work) {
                    ptx::tmem_ld_32x16(tmem_base + ((uint32_t)(ew * 32) << 16) +
                                           (uint32_t)((acc * p.halves + hf) * p.n_tile_alloc), r);
                    ptx::tmem_ld_wait();
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) tempty_arrive(acc);
                if (work) {
                    const int oh0 = t.oh0 + (p.split_b == 0 ? hf * p.bh : 0), b0 = t.b0 + (p.split_b == 1 ? hf * p.bb : 0);
                    const int ow = t.ow0 + (p.split_b == 2 ? hf * p.bw : 0) + rw, oh = oh0 + rh, b = b0 + rb;
                    float h1[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) h1[k] = fmaxf(__uint_as_float(r[k]) + tw[306 + k], 0.f);
                    float o0 = tw[304], o1 = tw[305];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float a = tw[256 + j];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 w4 = *reinterpret_cast<const float4*>(&tw[j * 16 + q * 4]);
                            a = fmaf(w4.x, h1[q * 4 + 0], a);
                            a = fmaf(w4.y, h1[q * 4 + 1], a);
                            a = fmaf(w4.z, h1[q * 4 + 2], a);
                            a = fmaf(w4.w, h1[q * 4 + 3], a);
                        }
                        a = fmaxf(a, 0.f);
                        o0 = fmaf(tw[272 + j], a, o0);
                        o1 = fmaf(tw[288 + j], a, o1);
                    }
                    if ((ow < p.OW) && (oh < p.OH) && (b < p.B))
                        reinterpret_cast<float2*>(p.tail_out)[((long)b * p.OH + oh) * p.OW + ow] = make_float2(o0, o1);
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        } else if (EPI == 32 && HALVES == 2 && (FIRST || (p.n_chunks == 1 && p.tiles_n == 1)) &&
                   !(LOCR_CONV_EXPERIMENTS && (p.dbg & 256))) {
            // ---- plain 16-bit TMA-store epilogue as TWO INDEPENDENT warpgroups.  A staging chunk costs a chain of
            // latencies (buffer free -> barrier -> TMEM load -> convert -> st.shared -> proxy fence -> barrier -> TMA
            // store) that eight warps marching in step cannot overlap; for the small-K layers that chain, not the MMAs,
            // set the tile rate.  Here each warpgroup (4 warps = all 128 TMEM lanes) takes every other (half, chunk) unit
            // of the tile with its own staging buffer, named barrier and store queue, so two chains are in flight.
            // Used for the 64-channel layers with M = 256 tiles (one 128-pixel half per group: `slice1.0` 0.55 -> 0.49 ms);
            // measured neutral or slightly worse for the 32-column and multi-chunk tiles, which keep the lock-step form.
            constexpr int kCols = 2 * (EPI & 63);      // columns per staging chunk: 64 or 32
            constexpr int kRb = 2 * kCols;             // bytes per staged row: 128 or 64
            const int g = half;
            const int gtid = threadIdx.x - 128 - g * 128;
            uint8_t* sbuf = staging + (size_t)g * 128 * kRb;
            const uint32_t xor_term = ((((uint32_t)row * (uint32_t)kRb) >> 7) & (uint32_t)(kRb / 16 - 1)) << 4;
            const uint32_t row_off = (uint32_t)row * (uint32_t)kRb;
            for (int i = threadIdx.x - 128; i < p.n_tile; i += 256) bias_s[i] = __ldg(&p.bias[i]);
            ptx::named_bar_sync(1, 256);
            const int units = p.halves * p.n_chunks;
            int last_u = -1;
            for (int u = g; u < units; u += 2) last_u = u;
            for (int tile = t_begin; tile < t_end; tile += t_step) {
                const TileCoord t = decode_tile(p, tile_of(tile));
                ptx::mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
                ptx::tc_fence_after();
                if (last_u < 0) {          // single-unit tiles: this group only takes part in the TMEM hand-back
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tempty_arrive(acc);
                }
                for (int u = g; u < units; u += 2) {
                    const int hf = u / p.n_chunks, c = u - hf * p.n_chunks;
                    const int oh0 = t.oh0 + (p.split_b == 0 ? hf * p.bh : 0), b0 = t.b0 + (p.split_b == 1 ? hf * p.bb : 0);
                    const int ow0 = t.ow0 + (p.split_b == 2 ? hf * p.bw : 0);
                    if (gtid == 0) ptx::tma_store_wait_read<0>();     // this group's previous store has left its buffer
                    ptx::named_bar_sync(2 + g, 128);
                    const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) +
                                           (uint32_t)((acc * p.halves + hf) * p.n_tile_alloc + c * kCols);
                    uint32_t r0[32], r1[32];
                    ptx::tmem_ld_32x32(taddr, r0);
                    if (kCols == 64) ptx::tmem_ld_32x32(taddr + 32u, r1);
                    ptx::tmem_ld_wait();
                    if (u == last_u) {     // accumulators fully read by this group
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tempty_arrive(acc);
                    }
#pragma unroll
                    for (int q = 0; q < kCols / 8; ++q) {
                        const float4 ba = *reinterpret_cast<const float4*>(&bias_s[c * kCols + q * 8]);
                        const float4 bb4 = *reinterpret_cast<const float4*>(&bias_s[c * kCols + q * 8 + 4]);
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            v[j] = __uint_as_float(q < 4 ? r0[(q & 3) * 8 + j] : r1[(q & 3) * 8 + j]);
                        v[0] += ba.x; v[1] += ba.y; v[2] += ba.z; v[3] += ba.w;
                        v[4] += bb4.x; v[5] += bb4.y; v[6] += bb4.z; v[7] += bb4.w;
                        if (p.relu) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
                        }
                        uint4 o;
                        o.x = pack2(v[0], v[1], p.is_f16); o.y = pack2(v[2], v[3], p.is_f16);
                        o.z = pack2(v[4], v[5], p.is_f16); o.w = pack2(v[6], v[7], p.is_f16);
                        *reinterpret_cast<uint4*>(sbuf + row_off + (((uint32_t)q << 4) ^ xor_term)) = o;
                    }
                    ptx::fence_proxy_async();
                    ptx::named_bar_sync(2 + g, 128);
                    if (gtid == 0) {
                        if (!(LOCR_CONV_EXPERIMENTS && (p.dbg & 8)))
                            ptx::tma_store_4d(&tmap_y, sbuf, t.n0 + c * kCols, ow0, oh0, b0);
                        ptx::tma_store_commit();
                    }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
            if (gtid == 0) ptx::tma_store_wait_all();
        } else if (EPI != 0 || p.tma_store) {
            constexpr int kCpw = EPI & 63;
            const bool e_res = EPI == 0 ? p.res != nullptr : (EPI & kEpiRes) != 0;
            const int e_fp32 = EPI == 0 ? p.out_fp32 : 0, e_split = EPI == 0 ? p.split_out : 0;
            const int e_pool = EPI == 0 ? p.pool : ((EPI & kEpiPool) ? 1 : 0), e_halo_pool = EPI == 0 ? p.halo_pool : 0;
            const int e_skip_full = EPI == 0 ? p.skip_full : ((EPI & kEpiSkip) ? 1 : 0);
            const int stage_cols = EPI ? 2 * kCpw : p.stage_cols, stage_rb = EPI ? 4 * kCpw : p.stage_rb;
            // 8 epilogue warps: warp (ew, half) owns TMEM lanes [32*ew, +32) and one half of every staging chunk's
            // columns, so each scheduler overlaps two warps' worth of TMEM loads / conversions / smem stores.
            const int etid = threadIdx.x - 128;
            const int cpw = EPI ? kCpw : (stage_cols >> 1);            // columns per warp per chunk: 32, 16 or 8
            const uint32_t xor_term = ((((uint32_t)row * (uint32_t)stage_rb) >> 7) &
                                       (uint32_t)(stage_rb / 16 - 1)) << 4;
            const uint32_t row_off = (uint32_t)row * (uint32_t)stage_rb;
            uint32_t chunk_ctr = 0;
            // fused 2x2 max-pool: thread = (pooled pixel, 16-byte piece); its four source rows of the staged tile
            const int pool_pr = etid >> 3;
            const uint32_t pool_piece = (uint32_t)(etid & 7) << 4;
            const bool pool_active = e_pool && (int)pool_piece < stage_rb;
            uint32_t pool_src[4] = {0, 0, 0, 0}, pool_dst = 0;
            if (pool_active) {
                const int pw2 = p.bw >> 1, ph2 = p.bh >> 1;
                const int qw = pool_pr % pw2, qh = (pool_pr / pw2) % ph2, qb = pool_pr / (pw2 * ph2);
                const int r00 = (qb * p.bh + 2 * qh) * p.bw + 2 * qw;
                const int rows[4] = {r00, r00 + 1, r00 + p.bw, r00 + p.bw + 1};
                const uint32_t msk = (uint32_t)(stage_rb / 16 - 1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t off = (uint32_t)rows[i] * (uint32_t)stage_rb;
                    pool_src[i] = off + (pool_piece ^ (((off >> 7) & msk) << 4));
                }
                const uint32_t offp = (uint32_t)pool_pr * (uint32_t)stage_rb;
                pool_dst = offp + (pool_piece ^ (((offp >> 7) & msk) << 4));
            }
            // one n-tile: the bias slice is the same for every tile of this CTA (the first chunk's barrier publishes it)
            const bool bias_once = p.tiles_n == 1;
            const bool no_bias = EPI == 0 && p.bias == nullptr;     // split-K partial sums: the reduction adds the bias
            if (bias_once)
                for (int i = etid; i < p.n_tile; i += 256) bias_s[i] = no_bias ? 0.f : __ldg(&p.bias[i]);
            for (int tile = t_begin; tile < t_end; tile += t_step) {
                const TileCoord t = decode_tile(p, tile_of(tile));
                // split-K: slice s of the K range writes its partial sums at channel s * ks_out_stride + n
                const int n_out = t.n0 + ((EPI == 0 && p.ksplit > 1) ? (tile / p.base_tiles) * p.ks_out_stride : 0);
                if (etid == 0) LOCR_TRACE(2, 0);
                if (!bias_once)
                    for (int i = etid; i < p.n_tile; i += 256) bias_s[i] = no_bias ? 0.f : __ldg(&p.bias[t.n0 + i]);
                ptx::mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
                if (etid == 0) LOCR_TRACE(2, 1);
                ptx::tc_fence_after();
                if (LOCR_CONV_EXPERIMENTS && (p.dbg & 16)) {   // epilogue reduced to the TMEM hand-shake
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tempty_arrive(acc);
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1u;
                    continue;
                }
                if (e_halo_pool) {
                    // 16 x 16 pixel tile = two 8-wide halves, lane = 8 * (row & 3) + column: the 2x2 window of a pixel
                    // is lanes {l, l ^ 1, l ^ 8} x their combination, so the pool is two shuffles on the packed 16-bit
                    // pairs (max commutes with the rounding) and the full-resolution tile never touches shared
                    // memory — the N = 64 MMAs of this layer already use most of its bandwidth.  The 8 x 8 pooled
                    // tile leaves through one TMA store; it has a whole tile time to drain before its buffer is reused.
                    if (etid == 0) ptx::tma_store_wait_read<0>();
                    ptx::named_bar_sync(1, 256);
                    if (etid == 0) LOCR_TRACE(2, 2);
                    uint32_t r0[32], r1[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) +
                                           (uint32_t)(acc * 2 * p.n_tile_alloc + half * 32);
                    ptx::tmem_ld_32x32(taddr, r0);
                    ptx::tmem_ld_32x32(taddr + (uint32_t)p.n_tile_alloc, r1);
                    ptx::tmem_ld_wait();
                    if (etid == 0) LOCR_TRACE(2, 3);
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tempty_arrive(acc);
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t hv[16];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 b4 = *reinterpret_cast<const float4*>(&bias_s[half * 32 + q * 4]);
                            float v0 = __uint_as_float(hf ? r1[q * 4 + 0] : r0[q * 4 + 0]) + b4.x;
                            float v1 = __uint_as_float(hf ? r1[q * 4 + 1] : r0[q * 4 + 1]) + b4.y;
                            float v2 = __uint_as_float(hf ? r1[q * 4 + 2] : r0[q * 4 + 2]) + b4.z;
                            float v3 = __uint_as_float(hf ? r1[q * 4 + 3] : r0[q * 4 + 3]) + b4.w;
                            if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
                            hv[q * 2 + 0] = pack2(v0, v1, p.is_f16);
                            hv[q * 2 + 1] = pack2(v2, v3, p.is_f16);
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            uint32_t m = max2_16(hv[j], __shfl_xor_sync(0xffffffffu, hv[j], 1), p.is_f16);
                            hv[j] = max2_16(m, __shfl_xor_sync(0xffffffffu, m, 8), p.is_f16);
                        }
                        if ((lane & 9) == 0) {
                            const uint32_t pr = (uint32_t)((2 * ew + (lane >> 4)) * 8 + hf * 4 + ((lane >> 1) & 3));
                            uint8_t* prow = pool_staging + pr * 128u;
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                *reinterpret_cast<uint4*>(prow + ((((uint32_t)(half * 4 + q)) ^ (pr & 7u)) << 4)) =
                                    make_uint4(hv[q * 4 + 0], hv[q * 4 + 1], hv[q * 4 + 2], hv[q * 4 + 3]);
                        }
                    }
                    if (etid == 0) LOCR_TRACE(2, 5);
                    ptx::fence_proxy_async();
                    ptx::named_bar_sync(1, 256);
                    if (etid == 0) LOCR_TRACE(2, 7);
                    if (etid == 0) {
                        ptx::tma_store_4d(&tmap_p, pool_staging, t.n0, t.ow0 >> 1, t.oh0 >> 1, t.b0);
                        ptx::tma_store_commit();
                    }
                    if (etid == 0) LOCR_TRACE(2, 8);
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1u;
                    continue;
                }
                for (int hf = 0; hf < p.halves; ++hf) {
                const int oh0 = t.oh0 + (p.split_b == 0 ? hf * p.bh : 0), b0 = t.b0 + (p.split_b == 1 ? hf * p.bb : 0);
                const int ow0 = t.ow0 + (p.split_b == 2 ? hf * p.bw : 0);
                const int ow = ow0 + rw, oh = oh0 + rh, b = b0 + rb;
                const bool valid = (ow < p.OW) && (oh < p.OH) && (b < p.B);
                const long pix = ((long)b * p.OH + oh) * p.OW + ow;
                for (int c = 0; c < p.n_chunks; ++c) {
                    uint8_t* sbuf = staging + (size_t)((e_split ? 0u : chunk_ctr) & 1u) * 128 * stage_rb;
                    uint8_t* sbuf_lo = staging + (size_t)128 * stage_rb;   // split-precision: second tile = lo parts
                    if (etid == 0) {
                        if (e_split) ptx::tma_store_wait_read<0>();
                        else ptx::tma_store_wait_read<1>();        // the store that used this buffer two chunks ago
                    }
                    ptx::named_bar_sync(1, 256);                    // ... is done; bias_s of this tile is visible
                    if (etid == 0) LOCR_TRACE(2, 2);
                    const int col0 = c * stage_cols + half * cpw;  // first column (within the n-tile) of this warp
                    const uint32_t taddr =
                        tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((acc * p.halves + hf) * p.n_tile_alloc + col0);
                    float v[32];
                    if (cpw == 32) {
                        uint32_t r[32];
                        ptx::tmem_ld_32x32(taddr, r);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    } else if (cpw == 16) {
                        uint32_t r[16];
                        ptx::tmem_ld_32x16(taddr, r);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                    } else {
                        uint32_t r[8];
                        ptx::tmem_ld_32x8(taddr, r);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
                    }
                    if (etid == 0) LOCR_TRACE(2, 3);
                    if (c == p.n_chunks - 1 && hf == p.halves - 1) {  // accumulators fully read: hand the TMEM stage back to the MMA warp
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tempty_arrive(acc);
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (q * 4 < cpw) {
                            const float4 b4 = *reinterpret_cast<const float4*>(&bias_s[col0 + q * 4]);
                            v[q * 4 + 0] += b4.x; v[q * 4 + 1] += b4.y; v[q * 4 + 2] += b4.z; v[q * 4 + 3] += b4.w;
                        }
                    }
                    if (e_res && valid) {
                        const int nb = t.n0 + col0;
                        const uint16_t* rp = reinterpret_cast<const uint16_t*>(p.res) + pix * p.res_pitch + nb;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (q * 8 < cpw && nb + q * 8 < p.Cout) {
                                const uint4 u = __ldg(reinterpret_cast<const uint4*>(rp + q * 8));
                                const float2 f0 = unpack2(u.x, p.is_f16), f1 = unpack2(u.y, p.is_f16);
                                const float2 f2 = unpack2(u.z, p.is_f16), f3 = unpack2(u.w, p.is_f16);
                                v[q * 8 + 0] += f0.x; v[q * 8 + 1] += f0.y;
                                v[q * 8 + 2] += f1.x; v[q * 8 + 3] += f1.y;
                                v[q * 8 + 4] += f2.x; v[q * 8 + 5] += f2.y;
                                v[q * 8 + 6] += f3.x; v[q * 8 + 7] += f3.y;
                                if (EPI == 0 && p.res_lo_off > 0) {   // split-precision residual: value = hi + lo
                                    const uint4 ul = __ldg(reinterpret_cast<const uint4*>(rp + p.res_lo_off + q * 8));
                                    const float2 l0 = unpack2(ul.x, p.is_f16), l1 = unpack2(ul.y, p.is_f16);
                                    const float2 l2 = unpack2(ul.z, p.is_f16), l3 = unpack2(ul.w, p.is_f16);
                                    v[q * 8 + 0] += l0.x; v[q * 8 + 1] += l0.y;
                                    v[q * 8 + 2] += l1.x; v[q * 8 + 3] += l1.y;
                                    v[q * 8 + 4] += l2.x; v[q * 8 + 5] += l2.y;
                                    v[q * 8 + 6] += l3.x; v[q * 8 + 7] += l3.y;
                                }
                            }
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                    }
                    // swizzled staging write: 16-byte piece k of this row lands at (k*16) ^ xor_term
                    if (e_fp32) {
                        const uint32_t piece0 = (uint32_t)(half * cpw) >> 2;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (q * 4 < cpw) {
                                uint4 u;
                                u.x = __float_as_uint(v[q * 4 + 0]); u.y = __float_as_uint(v[q * 4 + 1]);
                                u.z = __float_as_uint(v[q * 4 + 2]); u.w = __float_as_uint(v[q * 4 + 3]);
                                *reinterpret_cast<uint4*>(sbuf + row_off + (((piece0 + q) << 4) ^ xor_term)) = u;
                            }
                        }
                    } else {
                        const uint32_t piece0 = (uint32_t)(half * cpw) >> 3;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (q * 8 < cpw) {
                                uint4 u;
                                u.x = pack2(v[q * 8 + 0], v[q * 8 + 1], p.is_f16);
                                u.y = pack2(v[q * 8 + 2], v[q * 8 + 3], p.is_f16);
                                u.z = pack2(v[q * 8 + 4], v[q * 8 + 5], p.is_f16);
                                u.w = pack2(v[q * 8 + 6], v[q * 8 + 7], p.is_f16);
                                *reinterpret_cast<uint4*>(sbuf + row_off + (((piece0 + q) << 4) ^ xor_term)) = u;
                                if (e_split) {
                                    const float2 h0 = unpack2(u.x, p.is_f16), h1 = unpack2(u.y, p.is_f16);
                                    const float2 h2 = unpack2(u.z, p.is_f16), h3 = unpack2(u.w, p.is_f16);
                                    uint4 l;
                                    l.x = pack2(v[q * 8 + 0] - h0.x, v[q * 8 + 1] - h0.y, p.is_f16);
                                    l.y = pack2(v[q * 8 + 2] - h1.x, v[q * 8 + 3] - h1.y, p.is_f16);
                                    l.z = pack2(v[q * 8 + 4] - h2.x, v[q * 8 + 5] - h2.y, p.is_f16);
                                    l.w = pack2(v[q * 8 + 6] - h3.x, v[q * 8 + 7] - h3.y, p.is_f16);
                                    *reinterpret_cast<uint4*>(sbuf_lo + row_off + (((piece0 + q) << 4) ^ xor_term)) = l;
                                }
                            }
                        }
                    }
                    if (etid == 0) LOCR_TRACE(2, 4);
                    uint8_t* pbuf = pool_staging + (size_t)((e_split ? 0u : chunk_ctr) & 1u) * 32 * stage_rb;
                    uint8_t* pbuf_lo = pool_staging + (size_t)32 * stage_rb;
                    if (e_pool) {
                        ptx::named_bar_sync(1, 256);   // the whole 128-pixel chunk is staged
                        if (pool_active) {
                            float m[8];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint4 u = *reinterpret_cast<const uint4*>(sbuf + pool_src[i]);
                                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
                                float f[8];
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float2 v2 = unpack2(w4[q], p.is_f16);
                                    f[2 * q] = v2.x; f[2 * q + 1] = v2.y;
                                }
                                if (e_split) {   // value = hi + lo, exact in fp32
                                    const uint4 ul = *reinterpret_cast<const uint4*>(sbuf_lo + pool_src[i]);
                                    const uint32_t l4[4] = {ul.x, ul.y, ul.z, ul.w};
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const float2 v2 = unpack2(l4[q], p.is_f16);
                                        f[2 * q] += v2.x; f[2 * q + 1] += v2.y;
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < 8; ++j) m[j] = i == 0 ? f[j] : fmaxf(m[j], f[j]);
                            }
                            uint4 hi;
                            hi.x = pack2(m[0], m[1], p.is_f16); hi.y = pack2(m[2], m[3], p.is_f16);
                            hi.z = pack2(m[4], m[5], p.is_f16); hi.w = pack2(m[6], m[7], p.is_f16);
                            *reinterpret_cast<uint4*>(pbuf + pool_dst) = hi;
                            if (e_split) {
                                const float2 h0 = unpack2(hi.x, p.is_f16), h1 = unpack2(hi.y, p.is_f16);
                                const float2 h2 = unpack2(hi.z, p.is_f16), h3 = unpack2(hi.w, p.is_f16);
                                uint4 lo;
                                lo.x = pack2(m[0] - h0.x, m[1] - h0.y, p.is_f16);
                                lo.y = pack2(m[2] - h1.x, m[3] - h1.y, p.is_f16);
                                lo.z = pack2(m[4] - h2.x, m[5] - h2.y, p.is_f16);
                                lo.w = pack2(m[6] - h3.x, m[7] - h3.y, p.is_f16);
                                *reinterpret_cast<uint4*>(pbuf_lo + pool_dst) = lo;
                            }
                        }
                    }
                    if (etid == 0) LOCR_TRACE(2, 5);
                    ptx::fence_proxy_async();
                    if (etid == 0) LOCR_TRACE(2, 6);
                    ptx::named_bar_sync(1, 256);
                    if (etid == 0) LOCR_TRACE(2, 7);
                    if (etid == 0) {
                        if (!e_skip_full && !(LOCR_CONV_EXPERIMENTS && (p.dbg & 8))) {
                            ptx::tma_store_4d(&tmap_y, sbuf, n_out + c * stage_cols, ow0, oh0, b0);
                            if (e_split)
                                ptx::tma_store_4d(&tmap_y, sbuf_lo, p.Cout + t.n0 + c * stage_cols, ow0, oh0, b0);
                        }
                        if (e_pool) {
                            ptx::tma_store_4d(&tmap_p, pbuf, t.n0 + c * stage_cols, ow0 >> 1, oh0 >> 1, b0);
                            if (e_split)
                                ptx::tma_store_4d(&tmap_p, pbuf_lo, p.Cout + t.n0 + c * stage_cols, ow0 >> 1, oh0 >> 1, b0);
                        }
                        ptx::tma_store_commit();
                    }
                    if (etid == 0) LOCR_TRACE(2, 8);
                    ++chunk_ctr;
                }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
            if (etid == 0) ptx::tma_store_wait_all();
        } else if (half == 1) {
            // direct-store path uses four warps; the second warpgroup only keeps the TMEM hand-shake balanced
            for (int tile = t_begin; tile < t_end; tile += t_step) {
                ptx::mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
                __syncwarp();
                if (lane == 0) tempty_arrive(acc);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        } else
        for (int tile = t_begin; tile < t_end; tile += t_step) {
            const TileCoord t = decode_tile(p, tile_of(tile));
            const int ow = t.ow0 + rw, oh = t.oh0 + rh, b = t.b0 + rb;
            const bool valid = (ow < p.OW) && (oh < p.OH) && (b < p.B);
            const long pix = ((long)b * p.OH + oh) * p.OW + ow;
            ptx::mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
            ptx::tc_fence_after();
            for (int c = 0; c < chunks; ++c) {
                uint32_t r[32];
                const uint32_t taddr =
                    tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * p.n_tile_alloc + c * 32);
                ptx::tmem_ld_32x32(taddr, r);
                ptx::tmem_ld_wait();
                if (valid) {
                    const int nb = t.n0 + c * 32;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + __ldg(&p.bias[nb + j]);
                    if (p.res != nullptr) {
                        const uint16_t* rp = reinterpret_cast<const uint16_t*>(p.res) + pix * p.res_pitch + nb;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (nb + q * 8 < p.Cout) {
                                const uint4 u = __ldg(reinterpret_cast<const uint4*>(rp + q * 8));
                                const float2 f0 = unpack2(u.x, p.is_f16), f1 = unpack2(u.y, p.is_f16);
                                const float2 f2 = unpack2(u.z, p.is_f16), f3 = unpack2(u.w, p.is_f16);
                                v[q * 8 + 0] += f0.x; v[q * 8 + 1] += f0.y;
                                v[q * 8 + 2] += f1.x; v[q * 8 + 3] += f1.y;
                                v[q * 8 + 4] += f2.x; v[q * 8 + 5] += f2.y;
                                v[q * 8 + 6] += f3.x; v[q * 8 + 7] += f3.y;
                            }
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                    }
                    if (p.out_fp32) {
                        float* yp = reinterpret_cast<float*>(p.y) + pix * p.y_pitch + nb;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (nb + j < p.Cout) yp[j] = v[j];
                    } else {
                        uint16_t* yp = reinterpret_cast<uint16_t*>(p.y) + pix * p.y_pitch + nb;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (nb + q * 8 < p.Cout) {
                                uint4 u;
                                u.x = pack2(v[q * 8 + 0], v[q * 8 + 1], p.is_f16);
                                u.y = pack2(v[q * 8 + 2], v[q * 8 + 3], p.is_f16);
                                u.z = pack2(v[q * 8 + 4], v[q * 8 + 5], p.is_f16);
                                u.w = pack2(v[q * 8 + 6], v[q * 8 + 7], p.is_f16);
                                *reinterpret_cast<uint4*>(yp + q * 8) = u;
                            }
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) tempty_arrive(acc);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (CTA2) {      // neither CTA frees its TMEM (or exits) while the pair's MMAs / remote arrivals may still touch it
        ptx::cluster_arrive_release();
        ptx::cluster_wait_acquire();
    }
    if (warp == 2) {
        ptx::tc_fence_after();
        if (CTA2) ptx::tmem_dealloc_2sm(tmem_base, (uint32_t)p.tmem_cols);
        else ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<EncodeTiledFn>(p);
        }
    }
    return fn;
}

void set_err(char* err, int errlen, const char* msg) {
    if (err != nullptr && errlen > 0) {
        strncpy(err, msg, errlen - 1);
        err[errlen - 1] = 0;
    }
}

template <int SWZ, int HALVES, int EPI, bool FIRST = false, bool CTA2 = false>
cudaError_t launch_swz(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, const CUtensorMap& mp,
                       const ConvParams& p, int grid, size_t smem, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<SWZ, HALVES, EPI, FIRST, CTA2>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (CTA2) {      // clusters of two CTAs (one TPC): the pair shares every MMA
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, conv_tc_kernel<SWZ, HALVES, EPI, FIRST, CTA2>, mx, mw, my, mp, p);
    }
    conv_tc_kernel<SWZ, HALVES, EPI, FIRST, CTA2><<<grid, FIRST ? kFirstThreads : kThreads, smem, stream>>>(mx, mw, my, mp, p);
    return cudaGetLastError();
}

// Second half of a split-K call: y = act(sum over slices of the fp32 partial sums + bias [+ residual]) rounded to 16 bits.
// One thread per pixel and 8 output channels; ws is [pixels][slices * Cp] fp32.
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int slices, int Cp, long pixels,
                                     const float* __restrict__ bias, const uint16_t* __restrict__ res, long res_pitch,
                                     uint16_t* __restrict__ y, long y_pitch, int relu, int is_f16) {
    const int groups = Cp >> 3;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= pixels * groups) return;
    const long pix = idx / groups;
    const int n = (int)(idx - pix * groups) << 3;
    float v[8];
    {
        const float4 a = __ldg(reinterpret_cast<const float4*>(bias + n));
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    const float* q = ws + pix * ((long)slices * Cp) + n;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < slices; ++s, q += Cp) {
        const float4 a = *reinterpret_cast<const float4*>(q);
        const float4 b = *reinterpret_cast<const float4*>(q + 4);
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
        acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += acc[j];      // same order as the unsplit epilogue: accumulator + bias
    if (res != nullptr) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(res + pix * res_pitch + n));
        const float2 f0 = unpack2(u.x, is_f16), f1 = unpack2(u.y, is_f16);
        const float2 f2 = unpack2(u.z, is_f16), f3 = unpack2(u.w, is_f16);
        v[0] += f0.x; v[1] += f0.y; v[2] += f1.x; v[3] += f1.y;
        v[4] += f2.x; v[5] += f2.y; v[6] += f3.x; v[7] += f3.y;
    }
    if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
    }
    uint4 o;
    o.x = pack2(v[0], v[1], is_f16); o.y = pack2(v[2], v[3], is_f16);
    o.z = pack2(v[4], v[5], is_f16); o.w = pack2(v[6], v[7], is_f16);
    *reinterpret_cast<uint4*>(y + pix * y_pitch + n) = o;
}

}  // namespace

static std::atomic<long long> g_splitk_calls{0};
long long conv_tc_splitk_calls() { return g_splitk_calls.load(std::memory_order_relaxed); }

int device_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

cudaError_t conv_tc_launch(const ConvCall& c, cudaStream_t stream, char* err, int errlen) {
    EncodeTiledFn encode = get_encode_fn();
    if (encode == nullptr) {
        set_err(err, errlen, "cuTensorMapEncodeTiled entry point not available");
        return cudaErrorNotSupported;
    }
    const bool first = c.first_u8 != nullptr;
    if (first && (c.Cin != 32 || c.KH != 1 || c.KW != 1 || c.pad_h != 0 || c.pad_w != 0 || c.stride_h != 1 ||
                  c.Cout != 64 || c.Cout_pad != 64 || c.out_fp32 || c.split_out || c.cin_wrap != 0 || c.pool_y != nullptr ||
                  c.residual != nullptr || c.tail_out != nullptr || c.x_row_px != 0 || c.y_row_px != 0 ||
                  c.OH != c.H || c.OW != c.W || c.img_h > c.H || c.img_w > c.W)) {
        set_err(err, errlen, "conv_tc: the fused first layer is a 27(+5) -> 64 channel im2col GEMM at canvas resolution");
        return cudaErrorInvalidValue;
    }
    int swz = 0;
    if (c.Cin % 64 == 0) swz = 128;
    else if (c.Cin % 32 == 0) swz = 64;
    else if (c.Cin % 16 == 0) swz = 32;
    else {
        set_err(err, errlen, "conv_tc: Cin must be a multiple of 16");
        return cudaErrorInvalidValue;
    }
    const int block_k = swz / 2;
    if (c.stride_h != 1 && !(c.stride_h == 2 && c.KH == 2 && c.pad_h == 0 && c.dil_h == 1 && c.H % 2 == 0)) {
        set_err(err, errlen, "conv_tc: unsupported vertical stride configuration");
        return cudaErrorInvalidValue;
    }
    if (c.Cout_pad % 16 != 0 || c.x_pitch % 8 != 0 || (!c.out_fp32 && (c.y_pitch % 8 != 0 || c.Cout % 8 != 0))) {
        set_err(err, errlen, "conv_tc: channel counts / pitches must keep 16-byte alignment");
        return cudaErrorInvalidValue;
    }
    if (c.split_out && (c.out_fp32 || c.Cout != c.Cout_pad || c.Cout % 64 != 0)) {
        set_err(err, errlen, "conv_tc: split-precision output needs a 16-bit output with Cout a multiple of 64");
        return cudaErrorInvalidValue;
    }
    // ---- split-K (see ConvCall::splitk_ws): few output pixels, many k-blocks
    static int allow_splitk = -1;
    if (allow_splitk < 0) { const char* e = getenv("LOCR_CONV_SPLITK"); allow_splitk = e ? atoi(e) : 1; }
    if (allow_splitk && c.ksplit == 0 && c.splitk_ws != nullptr && !first && swz == 128 && c.n_tile == 0 &&
        c.cin_wrap == 0 && !c.split_out && !c.out_fp32 && c.pool_y == nullptr && c.tail_out == nullptr &&
        c.x_row_px == 0 && c.y_row_px == 0 && c.res_lo_off == 0 && c.Cout == c.Cout_pad && c.Cout_pad % 64 == 0 &&
        c.y_pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(c.y) % 16) == 0 &&
        (c.residual == nullptr || (c.res_pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(c.residual) % 16) == 0))) {
        const long pixels = (long)c.B * c.OH * c.OW;
        const int chunks = c.Cin / 64;
        // m-tiles of 128 pixels the generic kernel would walk (its box search below finds the same or fewer)
        long m_tiles = -1;
        for (int bw = 128; bw >= 1; bw >>= 1)
            for (int bh = 128 / bw; bh >= 1; bh >>= 1) {
                const int bb = 128 / (bw * bh);
                const long t = (long)((c.OW + bw - 1) / bw) * ((c.OH + bh - 1) / bh) * ((c.B + bb - 1) / bb);
                if (m_tiles < 0 || t < m_tiles) m_tiles = t;
            }
        int slices = 1;
        if (m_tiles <= 8 && c.KH * c.KW * chunks >= 8) {
            const long units = m_tiles * (c.Cout_pad / 64);
            for (int cand = 8; cand >= 2; cand >>= 1)
                if (chunks % cand == 0 && units * cand <= (long)device_sm_count() + units / 2) { slices = cand; break; }
        }
        if (slices > 1 && (size_t)pixels * slices * c.Cout_pad * 4 <= c.splitk_ws_bytes) {
            ConvCall part = c;
            part.ksplit = slices;
            part.n_tile = 64;
            part.y = c.splitk_ws;
            part.y_pitch = (long)slices * c.Cout_pad;
            part.out_fp32 = 1;
            part.bias = nullptr;
            part.residual = nullptr;
            part.relu = 0;
            part.splitk_ws = nullptr;
            cudaError_t e = conv_tc_launch(part, stream, err, errlen);
            if (e != cudaSuccess) return e;
            const long work = pixels * (c.Cout_pad / 8);
            splitk_reduce_kernel<<<(unsigned)((work + 255) / 256), 256, 0, stream>>>(
                reinterpret_cast<const float*>(c.splitk_ws), slices, c.Cout_pad, pixels, c.bias,
                reinterpret_cast<const uint16_t*>(c.residual), c.res_pitch, reinterpret_cast<uint16_t*>(c.y), c.y_pitch,
                c.relu, c.dtype == ACT_F16 ? 1 : 0);
            e = cudaGetLastError();
            if (e != cudaSuccess) set_err(err, errlen, cudaGetErrorString(e));
            else g_splitk_calls.fetch_add(1, std::memory_order_relaxed);
            return e;
        }
    }
    const int ksplit = c.ksplit > 1 ? c.ksplit : 1;
    int n_tile = c.n_tile;
    if (n_tile == 0) {
        if (c.Cout_pad % 256 == 0) n_tile = 256;
        else if (c.Cout_pad % 128 == 0) n_tile = 128;
        else if (c.Cout_pad <= 256) n_tile = c.Cout_pad;
        else if (c.Cout_pad % 64 == 0) n_tile = 64;
        else n_tile = 16;
    }
    if (n_tile < 16 || n_tile > 256 || n_tile % 16 != 0 || c.Cout_pad % n_tile != 0) {
        set_err(err, errlen, "conv_tc: invalid n_tile");
        return cudaErrorInvalidValue;
    }

    // Pick the 128-pixel patch shape (bw x bh x bb, all powers of two) that covers the output with fewest tiles.
    // An M = 256 tile is two such patches stacked along b (if the patch spans several images) or along h.  It is used
    // for n_tile <= 128: one TMA instruction then feeds two independent accumulators, which amortises the fixed
    // per-instruction cost of the producer and hides the latency of back-to-back MMAs into one accumulator.
    const int pool = c.pool_y != nullptr ? 1 : 0;
    if (pool && (c.out_fp32 || c.stride_h != 1)) {
        set_err(err, errlen, "conv_tc: fused max-pool needs a 16-bit, stride-1 output");
        return cudaErrorInvalidValue;
    }
    auto search = [&](int rows, int& obw, int& obh, int& obb) {
        long best = -1;
        for (int bw = rows > 256 ? 256 : rows; bw >= 1; bw >>= 1) {
            for (int bh = rows / bw; bh >= 1; bh >>= 1) {
                const int bb = rows / (bw * bh);
                if (bw > 256 || bh > 256 || bb > 256) continue;
                if (rows == 256 && !((bb > 1) || (bh > 1))) continue;   // must split into two 128-pixel halves
                // fused 2x2 pooling: every 128-pixel half must hold whole 2x2 windows
                if (pool && (bw < 2 || (rows == 256 && bb == 1 ? bh < 4 : bh < 2))) continue;
                const long tiles = (long)((c.OW + bw - 1) / bw) * ((c.OH + bh - 1) / bh) * ((c.B + bb - 1) / bb);
                if (best < 0 || tiles < best) {
                    best = tiles;
                    obw = bw; obh = bh; obb = bb;
                }
            }
        }
        return best;
    };
    int best_bw = 0, best_bh = 0, best_bb = 0;
    const long tiles128 = search(128, best_bw, best_bh, best_bb);
    int halves = 1, split_b = 0;
    {
        static int allow256 = -1;
        if (allow256 < 0) { const char* e = getenv("LOCR_CONV_M256"); allow256 = e ? atoi(e) : 1; }
        const int elem_ = c.out_fp32 ? 4 : 2;
        const int sc_ = n_tile < 128 / elem_ ? n_tile : 128 / elem_;
        const int rb_ = sc_ * elem_;
        const bool aligned_out = (c.y_pitch * elem_) % 16 == 0 && (reinterpret_cast<uintptr_t>(c.y) % 16) == 0 &&
                                 n_tile % sc_ == 0 && (rb_ == 32 || rb_ == 64 || rb_ == 128);
        int w2 = 0, h2 = 0, b2 = 0;
        if ((allow256 || first) && n_tile <= 128 && aligned_out && ksplit == 1) {
            const long tiles256 = search(256, w2, h2, b2);
            const long n_tiles_n = c.Cout_pad / n_tile;
            // worthwhile when it wastes no more pixels than the 128-row tiling and still fills the machine
            // (the fused first layer always takes M = 256 tiles: its producer and epilogue are written for them)
            if (tiles256 > 0 && (first || (tiles256 * 2 <= tiles128 + tiles128 / 16 && tiles256 * n_tiles_n >= 2L * device_sm_count()))) {
                halves = 2;
                if (b2 > 1) { split_b = 1; best_bw = w2; best_bh = h2; best_bb = b2 / 2; }
                else { split_b = 0; best_bw = w2; best_bh = h2 / 2; best_bb = b2; }
            }
        }
    }

    // Halo mode (3x3, stride 1, pad 1, 64 -> 64 channels): the A operand of all nine taps comes from ONE haloed patch
    // per 16 x 16 output tile (18 rows x 24 pixels x 128 B, one TMA load instead of nine boxes of twice the tile:
    // 55 KB instead of 360 KB of L2 -> SM traffic per 256 pixels) and the weights stay resident in shared memory.
    static int allow_halo = -1;
    if (allow_halo < 0) { const char* e = getenv("LOCR_CONV_HALO"); allow_halo = e ? atoi(e) : 1; }
    const int elem_h = 2;
    // The same form serves the narrow decoder-tail layers (64 -> 32, and 32 -> 64 / 32 / 16 channels on 64-byte pixels,
    // incl. the fused 1x1 tail): with one box per tap those layers pull 9 - 12 x their input through L2 -> SM and are
    // bound by it.  LOCR_CONV_HALO=2 restricts the mode to the 64 -> 64 layers again (A/B runs).
    const bool halo_wide = c.Cin == 64 && c.Cout_pad == 64 && n_tile == 64 && c.tail_out == nullptr;
    const bool halo_narrow = allow_halo == 1 && !pool &&
                             ((c.Cin == 64 && c.Cout_pad == 32 && n_tile == 32 && c.tail_out == nullptr) ||
                              (c.Cin == 32 && (c.Cout_pad == 64 || c.Cout_pad == 32 || c.Cout_pad == 16) &&
                               n_tile == c.Cout_pad));
    const bool halo_base = allow_halo && c.KH == 3 && c.KW == 3 && c.dil_h == 1 && c.dil_w == 1 && c.pad_h == 1 &&
                      c.pad_w == 1 && c.stride_h == 1 && (halo_wide || halo_narrow) && c.cin_wrap == 0 &&
                      !c.out_fp32 && !c.split_out && c.x_row_px == 0 && (c.y_row_px == 0 || halo_narrow) &&
                      c.residual == nullptr &&
                      (c.tail_out != nullptr ||
                       ((c.y_pitch * elem_h) % 16 == 0 && (reinterpret_cast<uintptr_t>(c.y) % 16) == 0));
    // hstream: the 128 -> 128 3x3 layers at high resolution (slice1.10) as CTA pairs over haloed patches (one per
    // 64-channel chunk, ring of two) against a streamed weight ring: per 256-pixel tile and CTA 55 KB of A per chunk
    // instead of nine 32 KB boxes - that layer is bound by shared-memory traffic (fills + operand reads), and the fills
    // drop from 40 KB to ~14 KB per k-block: 0.541 -> 0.430 ms per 8 canvases (1340 -> 1684 TFLOP/s).  Measured on the
    // other candidates: slower with ONE chunk (64 input channels, slice1.7: 0.314 -> 0.333 ms - a two-deep patch ring is
    // then only one tile of look-ahead), neutral with four (upconv2.conv.3, 256 channels at 160x120: 0.081 -> 0.084).
    // Only where 16 x 16 tiles waste <= 5 % of the pixels.
    static int allow_hstream = -1;
    if (allow_hstream < 0) { const char* e = getenv("LOCR_CONV_HSTREAM"); allow_hstream = e ? atoi(e) : 1; }
    bool hstream = false;
    if (allow_hstream && !halo_base && !first && c.KH == 3 && c.KW == 3 && c.dil_h == 1 && c.dil_w == 1 && c.pad_h == 1 &&
        c.pad_w == 1 && c.stride_h == 1 && c.Cin == 128 && c.cin_wrap == 0 && c.Cout_pad == 128 &&
        n_tile == 128 && !c.out_fp32 && !c.split_out && c.x_row_px == 0 && c.y_row_px == 0 && c.tail_out == nullptr &&
        c.residual == nullptr && (c.y_pitch * elem_h) % 16 == 0 && (reinterpret_cast<uintptr_t>(c.y) % 16) == 0) {
        const long t16 = (long)((c.OW + 15) / 16) * ((c.OH + 15) / 16) * c.B;
        if (t16 >= device_sm_count() && t16 * 256 * 100 <= (long)c.OW * c.OH * c.B * 105) hstream = true;
    }
    const bool halo = halo_base;
    if (halo || hstream) {
        halves = 2; split_b = 2;
        best_bw = 8; best_bh = 16; best_bb = 1;
    }

    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.B = c.B; p.OH = c.OH; p.OW = c.OW; p.Cout = c.Cout;
    p.bw = best_bw; p.bh = best_bh; p.bb = best_bb;
    p.halves = halves; p.split_b = split_b;
    p.halo = halo ? 1 : 0;
    p.hstream = hstream ? 1 : 0;
    const int full_bh = p.bh * (split_b == 0 ? halves : 1), full_bb = p.bb * (split_b == 1 ? halves : 1);
    p.tiles_w = (c.OW + p.bw * (split_b == 2 ? halves : 1) - 1) / (p.bw * (split_b == 2 ? halves : 1));
    p.tiles_h = (c.OH + full_bh - 1) / full_bh;
    const int tiles_b = (c.B + full_bb - 1) / full_bb;
    p.tiles_n = c.Cout_pad / n_tile;
    const long num_tiles = (long)p.tiles_w * p.tiles_h * tiles_b * p.tiles_n;
    if (num_tiles <= 0 || num_tiles > 0x7fffffffL) {
        set_err(err, errlen, "conv_tc: empty or oversized problem");
        return cudaErrorInvalidValue;
    }
    p.num_tiles = (int)num_tiles;
    p.n_tile = n_tile;
    p.n_tile_alloc = (n_tile + 31) / 32 * 32;
    int cols = 32;
    while (cols < 2 * halves * p.n_tile_alloc) cols <<= 1;
    p.tmem_cols = cols;
    p.cin_chunks = c.Cin / block_k / ksplit;     // per K slice
    p.ksplit = ksplit;
    p.base_tiles = p.num_tiles;
    p.cin_total = c.Cin;
    p.ks_out_stride = c.Cout_pad;
    if (ksplit > 1) p.num_tiles *= ksplit;
    p.KW = c.KW; p.dil_h = c.dil_h; p.dil_w = c.dil_w; p.pad_h = c.pad_h; p.pad_w = c.pad_w;
    p.stride2 = (c.stride_h == 2) ? 1 : 0;
    p.num_kblocks = c.KH * c.KW * p.cin_chunks;
    // CTA pairs (conv_tc_kernel CTA2) for the N = 128 layers with M = 256 tiles and K >= 1152 (18 k-blocks): measured
    // -11 % on slice1.10 (128 -> 128, 3x3) and -7 % on the CRNN's conv1, neutral at 9 k-blocks, and +30 % on 1x1 layers
    // with one or two k-blocks per tile, where the cross-CTA barrier round trips are not amortised.
    // LOCR_CONV_CTA2=0 switches them off, LOCR_CONV_CTA2=2 takes every eligible layer (tests, A/B runs).
    static int allow_cta2 = -1;
    if (allow_cta2 < 0) {
        const char* e = getenv("LOCR_CONV_CTA2");
        allow_cta2 = e ? atoi(e) : 1;
    }
    const int elem_c = c.out_fp32 ? 4 : 2;
    // N = 256 tiles (M = 128 per CTA, the 256- / 512- / 1024-channel layers) run as pairs too: M = 256 x N = 256 MMAs,
    // each CTA staging 16 KB of A and 16 KB (half) of B per k-block instead of 16 + 32 KB - measured 1497 -> 1617 TFLOP/s
    // on the 512-channel VGG layers (98 % of the burst peak) and, because the halved operand traffic also lowers the
    // power per FLOP, +6.5 % on the power-capped end-to-end bench.  LOCR_CONV_CTA2_N256=0 switches it off, =2 takes
    // every eligible layer; LOCR_CONV_CTA2_MINKB overrides the k-block threshold.
    static int allow_cta2_n256 = -1, cta2_min_kb = 18;
    if (allow_cta2_n256 < 0) {
        const char* e = getenv("LOCR_CONV_CTA2_N256");
        allow_cta2_n256 = e ? atoi(e) : 1;
        const char* m = getenv("LOCR_CONV_CTA2_MINKB");
        if (m) cta2_min_kb = atoi(m);
    }
    const int kblocks_all = c.KH * c.KW * p.cin_chunks;
    const bool cta2_shape = (halves == 2 && n_tile == 128 && (allow_cta2 >= 2 || kblocks_all >= cta2_min_kb)) ||
                            (halves == 1 && n_tile == 256 && allow_cta2_n256 &&
                             (allow_cta2_n256 >= 2 || kblocks_all >= (cta2_min_kb < 16 ? cta2_min_kb : 16)));
    // split-precision inputs / outputs and fp32 outputs take the pair form of the generic epilogue (EPI = 0)
    const bool cta2 = hstream || (allow_cta2 && ksplit == 1 && !first && !halo && swz == 128 && cta2_shape && c.tail_out == nullptr &&
                                  c.x_row_px == 0 && c.y_row_px == 0 && (long)p.tiles_w * p.tiles_h * tiles_b >= 2);
    (void)elem_c;
    const int n_load = cta2 ? n_tile / 2 : n_tile;      // weight rows each CTA stages per k-block
    p.num_pair_tiles = (int)((((long)p.tiles_w * p.tiles_h * tiles_b + 1) / 2) * p.tiles_n);
    p.a_stage_bytes = (uint32_t)(halves * kTileM * swz);
    p.b_stage_bytes = (uint32_t)((n_load * swz + 1023) / 1024 * 1024);
    if (halo) {
        p.a_stage_bytes = 18u * 24u * (uint32_t)swz;               // haloed patch
        p.b_stage_bytes = 9u * (uint32_t)(n_load * swz) / 2u;      // x 2 "stages" = the nine resident tap slabs
    }
    if (hstream) p.a_stage_bytes = 18u * 24u * 128u;               // one patch per 64-channel chunk; b = half slab (8 KB)
    size_t stage_bytes = (size_t)p.a_stage_bytes + p.b_stage_bytes;
    // epilogue staging: only when the output rows keep 16-byte alignment and the n-tile splits into whole chunks
    const int elem = c.out_fp32 ? 4 : 2;
    p.stage_cols = n_tile < 128 / elem ? n_tile : 128 / elem;
    p.stage_rb = p.stage_cols * elem;
    p.n_chunks = n_tile / p.stage_cols;
    p.tma_store = ((c.y_pitch * elem) % 16 == 0 && (reinterpret_cast<uintptr_t>(c.y) % 16) == 0 &&
                   n_tile % p.stage_cols == 0 && (p.stage_rb == 32 || p.stage_rb == 64 || p.stage_rb == 128))
                      ? 1 : 0;
    const size_t tail_bytes = (2 * kMaxStages + 6) * 8 + 256 * 4 + 2 * 128 * 128 + (pool ? 2 * 32 * 128 : 0) +
                              (first ? 3 * 256 * 2 : 0);
    int stages = (int)((227 * 1024 - 1024 - tail_bytes) / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages > p.num_kblocks && p.num_kblocks >= 2) stages = p.num_kblocks;
    if (stages < 2 || halo) stages = 2;
    if (halo) {
        // patch ring as deep as the resident slabs leave room for (one barrier round per tile: the prefetch distance is
        // what hides the load latency); the slabs are spread over `stages` equal, 1024-byte-aligned shares
        const size_t slabs = 9u * (size_t)(n_load * swz);
        static int halo_stages_max = -1;
        if (halo_stages_max < 0) { const char* e = getenv("LOCR_CONV_HALO_STAGES"); halo_stages_max = e ? atoi(e) : 6; }
        int st = (int)((227 * 1024 - 1024 - tail_bytes - slabs - 6 * 1024) / p.a_stage_bytes);
        if (st > halo_stages_max) st = halo_stages_max;
        if (st > kMaxStages - 2) st = kMaxStages - 2;      // barrier slot kMaxStages - 1 belongs to the resident weights
        if (st > 2) {
            stages = st;
            p.b_stage_bytes = (uint32_t)(((slabs + st - 1) / st + 1023) / 1024 * 1024);
            stage_bytes = (size_t)p.a_stage_bytes + p.b_stage_bytes;
        }
    }
    if (first && stages > 4) stages = 4;   // barrier slot kMaxStages - 1 belongs to the resident weights (like halo)
    p.a_slots = stages;
    if (hstream) {
        // two patch slots in front of a weight ring of up to kMaxStages - 2 half slabs (the last two barrier slots
        // belong to the patch ring)
        p.a_slots = 2;
        stages = (int)((227 * 1024 - 1024 - tail_bytes - 2 * (size_t)p.a_stage_bytes) / p.b_stage_bytes);
        if (stages > kMaxStages - 2) stages = kMaxStages - 2;
    }
    p.stages = stages;
    p.idesc = ptx::make_idesc_f16(c.dtype == ACT_BF16 ? 1 : 0, cta2 ? 2 * kTileM : kTileM, n_tile);
    p.pool = pool;
    p.skip_full = (pool && c.skip_full) ? 1 : 0;
    p.halo_pool = (halo && p.skip_full) ? 1 : 0;
    p.tail_w = c.tail_w;
    p.tail_out = c.tail_out;
    p.first_u8 = c.first_u8; p.img_h = c.img_h; p.img_w = c.img_w;
    p.img_row_stride = c.img_row_stride; p.img_stride = c.img_stride;
    if (first && (halves != 2 || !p.tma_store || p.stage_cols != 64)) {
        set_err(err, errlen, "conv_tc: the fused first layer needs M = 256 tiles and an aligned 64-channel output");
        return cudaErrorInvalidValue;
    }
    if (c.y_row_px > 0 && (!p.tma_store || c.residual != nullptr || pool)) {
        set_err(err, errlen, "conv_tc: row-padded outputs need the plain TMA-store epilogue");
        return cudaErrorInvalidValue;
    }
    if (c.tail_out != nullptr && (c.tail_w == nullptr || n_tile != 16 || c.Cout != 16 || pool || c.residual != nullptr)) {
        set_err(err, errlen, "conv_tc: the fused 1x1 tail needs Cout = 16 and no pooling / residual");
        return cudaErrorInvalidValue;
    }
    if (pool && !p.tma_store) {
        set_err(err, errlen, "conv_tc: fused max-pool needs the TMA-store epilogue (aligned 16-bit output)");
        return cudaErrorInvalidValue;
    }
    p.y = c.y; p.y_pitch = c.y_pitch; p.out_fp32 = c.out_fp32;
    p.res = c.residual; p.res_pitch = c.res_pitch; p.res_lo_off = c.res_lo_off;
    p.bias = c.bias; p.relu = c.relu; p.is_f16 = (c.dtype == ACT_F16) ? 1 : 0;
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("LOCR_CONV_DBG"); dbg = e ? atoi(e) : 0; }
        p.dbg = dbg;
        p.cin_wrap = c.cin_wrap > 0 ? c.cin_wrap : 0x7fffffff;
        p.split_out = c.split_out;
    }

    const CUtensorMapDataType dt =
        (c.dtype == ACT_BF16) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const CUtensorMapSwizzle sw =
        swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);

    CUtensorMap mx, mw;
    memset(&mx, 0, sizeof(mx));
    if (!first) {
        const int S = p.stride2 ? 2 : 1;
        cuuint64_t dims[5] = {(cuuint64_t)(c.cin_wrap > 0 ? c.cin_wrap : c.Cin), (cuuint64_t)c.W, (cuuint64_t)S, (cuuint64_t)(c.H / S),
                              (cuuint64_t)c.B};
        const cuuint64_t pb = (cuuint64_t)c.x_pitch * 2;
        const cuuint64_t xw = (cuuint64_t)(c.x_row_px > 0 ? c.x_row_px : c.W);
        cuuint64_t strides[4] = {pb, pb * xw, pb * xw * S, pb * xw * c.H};
        cuuint32_t box[5] = {(cuuint32_t)block_k, (cuuint32_t)p.bw, 1u, (cuuint32_t)full_bh, (cuuint32_t)full_bb};
        if (halo || hstream) { box[1] = 24; box[3] = 18; box[4] = 1; }
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = encode(&mx, dt, 5, const_cast<void*>(c.x), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            char msg[160];
            snprintf(msg, sizeof(msg), "conv_tc: activation tensor map encode failed (CUresult %d)", (int)r);
            set_err(err, errlen, msg);
            return cudaErrorInvalidValue;
        }
    }
    {
        const cuuint64_t ktot = (cuuint64_t)c.KH * c.KW * c.Cin;
        cuuint64_t dims[2] = {ktot, (cuuint64_t)c.Cout_pad};
        cuuint64_t strides[1] = {ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)block_k, (cuuint32_t)n_load};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&mw, dt, 2, const_cast<void*>(c.w), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            char msg[160];
            snprintf(msg, sizeof(msg), "conv_tc: weight tensor map encode failed (CUresult %d)", (int)r);
            set_err(err, errlen, msg);
            return cudaErrorInvalidValue;
        }
    }

    CUtensorMap my;
    memset(&my, 0, sizeof(my));
    if (p.tma_store && !p.skip_full && c.tail_out == nullptr) {
        const cuuint64_t eb = (cuuint64_t)elem;
        cuuint64_t dims[4] = {(cuuint64_t)(c.split_out ? 2 * c.Cout : (ksplit > 1 ? ksplit * c.Cout_pad : c.Cout)),
                              (cuuint64_t)c.OW, (cuuint64_t)c.OH, (cuuint64_t)c.B};
        const cuuint64_t pb = (cuuint64_t)c.y_pitch * eb;
        const cuuint64_t yw = (cuuint64_t)(c.y_row_px > 0 ? c.y_row_px : c.OW);
        cuuint64_t strides[3] = {pb, pb * yw, pb * yw * c.OH};
        cuuint32_t box[4] = {(cuuint32_t)p.stage_cols, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bb};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUtensorMapDataType ydt = c.out_fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dt;
        const CUtensorMapSwizzle ysw = p.stage_rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                         : (p.stage_rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                             : CU_TENSOR_MAP_SWIZZLE_32B);
        CUresult r = encode(&my, ydt, 4, c.y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, ysw,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            char msg[160];
            snprintf(msg, sizeof(msg), "conv_tc: output tensor map encode failed (CUresult %d)", (int)r);
            set_err(err, errlen, msg);
            return cudaErrorInvalidValue;
        }
    }
    CUtensorMap mp;
    memset(&mp, 0, sizeof(mp));
    if (pool) {
        const int PH = c.OH / 2, PW = c.OW / 2;
        if (PH < 1 || PW < 1 || (c.pool_pitch * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(c.pool_y) % 16) != 0) {
            set_err(err, errlen, "conv_tc: bad pooled output");
            return cudaErrorInvalidValue;
        }
        cuuint64_t dims[4] = {(cuuint64_t)(c.split_out ? 2 * c.Cout : c.Cout), (cuuint64_t)PW, (cuuint64_t)PH,
                              (cuuint64_t)c.B};
        const cuuint64_t pb = (cuuint64_t)c.pool_pitch * 2;
        cuuint64_t strides[3] = {pb, pb * PW, pb * PW * PH};
        cuuint32_t box[4] = {(cuuint32_t)p.stage_cols, (cuuint32_t)(p.bw / 2), (cuuint32_t)(p.bh / 2), (cuuint32_t)p.bb};
        if (p.halo_pool) box[1] = (cuuint32_t)p.bw;   // both halves of the tile in one 8 x 8 pooled box
        cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUtensorMapSwizzle psw = p.stage_rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                         : (p.stage_rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                             : CU_TENSOR_MAP_SWIZZLE_32B);
        CUresult r = encode(&mp, dt, 4, c.pool_y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, psw,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            char msg[160];
            snprintf(msg, sizeof(msg), "conv_tc: pooled tensor map encode failed (CUresult %d)", (int)r);
            set_err(err, errlen, msg);
            return cudaErrorInvalidValue;
        }
    }
    const size_t smem = hstream ? 1024 + 2 * (size_t)p.a_stage_bytes + (size_t)p.stages * p.b_stage_bytes + tail_bytes
                                : 1024 + (size_t)p.stages * stage_bytes + tail_bytes;
    if (LOCR_CONV_EXPERIMENTS && getenv("LOCR_CONV_VERBOSE"))
        fprintf(stderr, "conv_tc: %dx%dx%d cin %d cout %d k%dx%d: swz %d halves %d split %d box %dx%dx%d n_tile %d stages %d "
                "kblocks %d tiles %d halo %d pool %d smem %zu\n", c.B, c.OH, c.OW, c.Cin, c.Cout, c.KH, c.KW, swz, halves,
                split_b, p.bw, p.bh, p.bb, n_tile, p.stages, p.num_kblocks, p.num_tiles, p.halo, p.pool, smem);
    int grid = p.num_tiles < device_sm_count() ? p.num_tiles : device_sm_count();
    if (cta2) {
        const int pairs_max = device_sm_count() / 2;
        grid = 2 * (p.num_pair_tiles < pairs_max ? p.num_pair_tiles : pairs_max);
    }
    cudaError_t e;
    // 16-bit TMA-store epilogues with compile-time options (see conv_tc_kernel); anything else takes the generic one
    static int allow_epi = -1;
    if (allow_epi < 0) { const char* ev = getenv("LOCR_CONV_EPI"); allow_epi = ev ? atoi(ev) : 1; }
    int epi = 0;
    // (64-byte operand rows: only the plain M = 256 / 64-column form, the 8-channel first layer)
    const bool epi_swz = swz == 128 || (swz == 64 && halves == 2 && c.residual == nullptr && !pool);
    if (allow_epi && epi_swz && p.tma_store && c.tail_out == nullptr && !c.split_out && !c.out_fp32 && !p.halo_pool &&
        c.res_lo_off == 0 &&
        (p.stage_cols == 64 || p.stage_cols == 32)) {
        epi = p.stage_cols / 2;
        if (c.residual != nullptr) epi |= kEpiRes;
        if (pool) epi |= kEpiPool | (p.skip_full ? kEpiSkip : 0);
    }
    bool done = false;
    e = cudaSuccess;
    if (first) {
        e = launch_swz<64, 2, 32, true>(mx, mw, my, mp, p, grid, smem, stream);
        done = true;
    }
    if (cta2 && !done) {
#define LOCR_PAIR_CASE(HV, E)                                                                      \
        if (!done && halves == HV && epi == (E)) {                                                 \
            e = launch_swz<128, HV, (E), false, true>(mx, mw, my, mp, p, grid, smem, stream);      \
            done = true;                                                                           \
        }
        LOCR_PAIR_CASE(2, 32) LOCR_PAIR_CASE(2, 32 | kEpiRes) LOCR_PAIR_CASE(2, 32 | kEpiPool)
        LOCR_PAIR_CASE(2, 32 | kEpiPool | kEpiSkip)
        LOCR_PAIR_CASE(1, 32) LOCR_PAIR_CASE(1, 32 | kEpiRes) LOCR_PAIR_CASE(1, 32 | kEpiPool)
        LOCR_PAIR_CASE(1, 32 | kEpiPool | kEpiSkip)
#undef LOCR_PAIR_CASE
        if (!done) {     // generic epilogue: fp32 / split-precision outputs, unaligned outputs
            if (halves == 2) e = launch_swz<128, 2, 0, false, true>(mx, mw, my, mp, p, grid, smem, stream);
            else e = launch_swz<128, 1, 0, false, true>(mx, mw, my, mp, p, grid, smem, stream);
            done = true;
        }
    }
    if (!done && swz == 64 && epi == 32) {
        e = launch_swz<64, 2, 32>(mx, mw, my, mp, p, grid, smem, stream);
        done = true;
    }
    if (!done && swz == 64 && epi == 16) {
        e = launch_swz<64, 2, 16>(mx, mw, my, mp, p, grid, smem, stream);
        done = true;
    }
#define LOCR_EPI_CASE(HV, E)                                                                       \
    if (!done && halves == HV && epi == (E)) {                                                     \
        e = launch_swz<128, HV, (E)>(mx, mw, my, mp, p, grid, smem, stream);                       \
        done = true;                                                                               \
    }
    LOCR_EPI_CASE(1, 32) LOCR_EPI_CASE(2, 32) LOCR_EPI_CASE(1, 16) LOCR_EPI_CASE(2, 16)
    LOCR_EPI_CASE(1, 32 | kEpiRes) LOCR_EPI_CASE(2, 32 | kEpiRes)
    LOCR_EPI_CASE(1, 32 | kEpiPool) LOCR_EPI_CASE(2, 32 | kEpiPool)
    LOCR_EPI_CASE(1, 32 | kEpiPool | kEpiSkip) LOCR_EPI_CASE(2, 32 | kEpiPool | kEpiSkip)
#undef LOCR_EPI_CASE
    if (!done) {
        if (halves == 2) {
            if (swz == 128) e = launch_swz<128, 2, 0>(mx, mw, my, mp, p, grid, smem, stream);
            else if (swz == 64) e = launch_swz<64, 2, 0>(mx, mw, my, mp, p, grid, smem, stream);
            else e = launch_swz<32, 2, 0>(mx, mw, my, mp, p, grid, smem, stream);
        } else {
            if (swz == 128) e = launch_swz<128, 1, 0>(mx, mw, my, mp, p, grid, smem, stream);
            else if (swz == 64) e = launch_swz<64, 1, 0>(mx, mw, my, mp, p, grid, smem, stream);
            else e = launch_swz<32, 1, 0>(mx, mw, my, mp, p, grid, smem, stream);
        }
    }
    if (e != cudaSuccess) set_err(err, errlen, cudaGetErrorString(e));
    return e;
}

int conv_tc_trace_read(unsigned long long* out, int* counts) {
#if LOCR_CONV_EXPERIMENTS
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out, g_conv_trace, sizeof(unsigned long long) * 3 * 8192) != cudaSuccess) return -1;
    for (int r = 0; r < 3; ++r) {
        int n = 0;
        while (n < 8192 && out[r * 8192 + n] != 0) ++n;
        counts[r] = n;
    }
    void* sym = nullptr;
    if (cudaGetSymbolAddress(&sym, g_conv_trace) != cudaSuccess) return -1;
    if (cudaMemset(sym, 0, sizeof(unsigned long long) * 3 * 8192) != cudaSuccess) return -1;
    return 0;
#else
    (void)out;
    counts[0] = counts[1] = counts[2] = 0;
    return 0;
#endif
}

}  // namespace locr
