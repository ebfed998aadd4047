// BiLSTM recurrence on the tensor cores (reference ocr/modules/biLSTM.py:18,24 = nn.LSTM(bidirectional, batch_first)).
//
// One CTA = 128 crops x one direction, persistent over the 26 time steps.  Per step
//     gates[128 crops, 1024] = xproj_t (precomputed GEMM, fp32) + h_{t-1}[128, 256] * W_hh^T
// runs as four passes of a 128 x 256 x 256 tcgen05 GEMM (64 hidden units x 4 gates per pass, columns ordered
// unit-major so that one 32-column TMEM load holds the i, f, g, o pre-activations of 8 units):
//   * A operand = h_{t-1} kept in shared memory in the UMMA K-major / 128B-swizzle layout, double buffered, written by
//     the epilogue warps of the previous step (16-bit storage, like every other activation of the path);
//   * B operand = W_hh slab [256 x 64] streamed by TMA from L2 (W_hh of one direction is 512 KB, more than one SM's
//     shared memory; it stays L2-resident, every CTA re-reads it each step);
//   * accumulators double-buffered in TMEM (2 x 256 columns), so the gate math of pass p overlaps the MMAs of p+1;
//   * 8 epilogue warps apply the gate non-linearities, update the fp32 cell state (global, L2-resident) and emit h_t
//     both into the next step's A operand and to the layer output.
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-11 = gate math.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <string.h>

#include "nn_kernels.cuh"
#include "ptx.cuh"

namespace locr {

namespace {

constexpr int kStagesW = 3;                 // W_hh slabs in flight
constexpr int kSlabBytes = 256 * 128;       // [256 gate columns][64 k] 16-bit
constexpr int kABytes = 4 * 128 * 128;      // h operand: 4 k-chunks x [128 crops][64 k]
constexpr int kThreadsLstm = 384;

struct LstmParams {
    const float* xproj;   // [B][T][2048], column = dir*1024 + pass*256 + unit_in_pass*4 + gate
    float* cstate;        // [gridDim.x*128][2][256] fp32 scratch
    uint16_t* out;        // [B][T][512]
    int B, T, is_f16;
    uint32_t idesc;
};

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
    // tanh(x) = 2 sigmoid(2x) - 1, evaluated with the same exp so that i/f/o and g/c share error behaviour
    const float e = __expf(-2.f * x);
    return __fdividef(1.f - e, 1.f + e);
}
__device__ __forceinline__ uint32_t pack2h(float a, float b, int f16) {
    if (f16) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kThreadsLstm, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const LstmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* a_buf = smem;                              // 2 x kABytes
    uint8_t* w_buf = smem + 2 * kABytes;                // kStagesW x kSlabBytes
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_buf + kStagesW * kSlabBytes);
    uint64_t* full_bar = bars;                          // [kStagesW]
    uint64_t* empty_bar = bars + kStagesW;              // [kStagesW]
    uint64_t* tfull_bar = bars + 2 * kStagesW;          // [2]
    uint64_t* tempty_bar = bars + 2 * kStagesW + 2;     // [2]
    uint64_t* hready_bar = bars + 2 * kStagesW + 4;     // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStagesW + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const int crop0 = blockIdx.x * 128;

    // h_0 = 0
    for (int i = threadIdx.x; i < kABytes / 16; i += kThreadsLstm) reinterpret_cast<uint4*>(a_buf)[i] = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async();
    if (warp == 0 && lane == 0) ptx::tma_prefetch_desc(&tmap_w);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStagesW; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tfull_bar[a], 1);
            ptx::mbar_init(&tempty_bar[a], 8);
            ptx::mbar_init(&hready_bar[a], 8);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_ptr_smem, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 1;
            for (int step = 0; step < p.T; ++step)
                for (int pass = 0; pass < 4; ++pass)
                    for (int kc = 0; kc < 4; ++kc) {
                        ptx::mbar_wait(&empty_bar[stage], phase, 500);
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], kSlabBytes);
                        ptx::tma_load_2d(w_buf + stage * kSlabBytes, &tmap_w, &full_bar[stage], kc * 64,
                                         dir * 1024 + pass * 256);
                        if (++stage == kStagesW) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t desc_hi = (uint32_t)(ptx::make_kmajor_desc(0, 128) >> 32);
            const uint32_t a_lo0 = ((ptx::smem_u32(a_buf) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t w_lo0 = ((ptx::smem_u32(w_buf) & 0x3FFFFu) >> 4) | (1u << 16);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 1;
            for (int step = 0; step < p.T; ++step) {
                if (step > 0) {  // wait for h_{step-1}: buffer step&1, its ((step-1)/2)-th completion
                    ptx::mbar_wait(&hready_bar[step & 1], (uint32_t)(((step - 1) >> 1) & 1), 600);
                }
                ptx::tc_fence_after();
                const uint32_t a_lo = a_lo0 + (uint32_t)(step & 1) * (kABytes >> 4);
                for (int pass = 0; pass < 4; ++pass) {
                    ptx::mbar_wait(&tempty_bar[acc], acc_phase, 610);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * 256u;
                    uint32_t accum = 0;
                    for (int kc = 0; kc < 4; ++kc) {
                        ptx::mbar_wait(&full_bar[stage], phase, 620);
                        ptx::tc_fence_after();
                        const uint32_t w_lo = w_lo0 + stage * (kSlabBytes >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            ptx::umma_f16_lohi(d_tmem, a_lo + kc * (16384 >> 4) + k * 2, desc_hi, w_lo + k * 2, desc_hi,
                                               p.idesc, accum);
                            accum = 1;
                        }
                        ptx::umma_commit(&empty_bar[stage]);
                        if (++stage == kStagesW) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    ptx::umma_commit(&tfull_bar[acc]);
                    acc ^= 1u;
                    if (acc == 0) acc_phase ^= 1u;
                }
            }
        }
    } else if (warp >= 4) {
        const int quarter = warp & 3;                 // TMEM lane quarter
        const int chalf = (warp - 4) >> 2;            // which 128 of the pass's 256 columns (32 units)
        const int row = quarter * 32 + lane;          // crop within the CTA
        const int crop = crop0 + row;
        const bool valid = crop < p.B;
        float* cst = p.cstate + ((size_t)(blockIdx.x * 128 + row) * 2 + dir) * 256;
        for (int i = 0; i < 128; i += 4)
            *reinterpret_cast<float4*>(cst + chalf * 32 + (i >> 5) * 64 + (i & 31)) = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t acc = 0, acc_phase = 0;
        for (int step = 0; step < p.T; ++step) {
            const int t = dir == 0 ? step : p.T - 1 - step;
            uint8_t* a_next = a_buf + ((step + 1) & 1) * kABytes;
            const float* xp_row = p.xproj + ((size_t)(valid ? crop : 0) * p.T + t) * 2048 + dir * 1024;
            uint16_t* out_row = p.out + ((size_t)(valid ? crop : 0) * p.T + t) * 512 + dir * 256;
            for (int pass = 0; pass < 4; ++pass) {
                ptx::mbar_wait(&tfull_bar[acc], acc_phase, 700);
                ptx::tc_fence_after();
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    const int col0 = chalf * 128 + ch * 32;            // column within the pass
                    const int unit0 = pass * 64 + (col0 >> 2);         // first of 8 hidden units
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 256u + (uint32_t)col0, r);
                    ptx::tmem_ld_wait();
                    if (ch == 3) {  // this warp's part of the accumulator is in registers
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
                    }
                    float hv[8];
                    const float4* xp4 = reinterpret_cast<const float4*>(xp_row + pass * 256 + col0);
                    float4 c0 = *reinterpret_cast<float4*>(cst + unit0), c1 = *reinterpret_cast<float4*>(cst + unit0 + 4);
                    float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        float4 x = valid ? __ldg(xp4 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float gi = fast_sigmoid(__uint_as_float(r[u * 4 + 0]) + x.x);
                        const float gf = fast_sigmoid(__uint_as_float(r[u * 4 + 1]) + x.y);
                        const float gg = fast_tanh(__uint_as_float(r[u * 4 + 2]) + x.z);
                        const float go = fast_sigmoid(__uint_as_float(r[u * 4 + 3]) + x.w);
                        cc[u] = gf * cc[u] + gi * gg;
                        hv[u] = go * fast_tanh(cc[u]);
                    }
                    *reinterpret_cast<float4*>(cst + unit0) = make_float4(cc[0], cc[1], cc[2], cc[3]);
                    *reinterpret_cast<float4*>(cst + unit0 + 4) = make_float4(cc[4], cc[5], cc[6], cc[7]);
                    uint4 hp;
                    hp.x = pack2h(hv[0], hv[1], p.is_f16);
                    hp.y = pack2h(hv[2], hv[3], p.is_f16);
                    hp.z = pack2h(hv[4], hv[5], p.is_f16);
                    hp.w = pack2h(hv[6], hv[7], p.is_f16);
                    // next step's A operand: k-chunk unit0/64, row `row`, 16-byte piece (unit0%64)/8, 128B swizzle
                    const uint32_t kc = (uint32_t)unit0 >> 6, piece = ((uint32_t)unit0 & 63u) >> 3;
                    *reinterpret_cast<uint4*>(a_next + kc * 16384u + (uint32_t)row * 128u +
                                              ((piece ^ ((uint32_t)row & 7u)) << 4)) = hp;
                    if (valid) *reinterpret_cast<uint4*>(out_row + unit0) = hp;
                }
                acc ^= 1u;
                if (acc == 0) acc_phase ^= 1u;
            }
            ptx::fence_proxy_async();   // h_t (generic-proxy stores) must be visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&hready_bar[(step + 1) & 1]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

size_t lstm_tc_cstate_bytes(int B) { return (size_t)((B + 127) / 128) * 128 * 2 * 256 * sizeof(float); }

cudaError_t launch_lstm_tc(const float* xproj, const void* whh_perm, float* cstate, void* out, int B, int T, int is_f16,
                           cudaStream_t s) {
    static EncodeTiledFn encode = nullptr;
    if (encode == nullptr) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return cudaErrorNotSupported;
        encode = reinterpret_cast<EncodeTiledFn>(fp);
    }
    CUtensorMap mw;
    cuuint64_t dims[2] = {256, 2048};
    cuuint64_t strides[1] = {512};
    cuuint32_t box[2] = {64, 256};
    cuuint32_t estr[2] = {1, 1};
    if (encode(&mw, is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
               const_cast<void*>(whh_perm), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    static bool attr = false;
    const size_t smem = 1024 + 2 * kABytes + kStagesW * kSlabBytes + (2 * kStagesW + 8) * 8;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    LstmParams p;
    p.xproj = xproj; p.cstate = cstate; p.out = (uint16_t*)out; p.B = B; p.T = T; p.is_f16 = is_f16;
    p.idesc = ptx::make_idesc_f16(is_f16 ? 0 : 1, 128, 256);
    dim3 grid((B + 127) / 128, 2);
    lstm_tc_kernel<<<grid, kThreadsLstm, smem, s>>>(mw, p);
    return cudaGetLastError();
}

}  // namespace locr
