// Host-side helpers shared by the translation units of liblocr: error plumbing, 16-bit conversions, device buffers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "locr.h"

namespace locr {

// Last error text of the calling thread (handles keep their own copy as well).
std::string& tls_error();
int fail(int code, const std::string& msg);

#define LOCR_CUDA_OK(expr)                                                                          \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            char _b[512];                                                                           \
            snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                     __FILE__, __LINE__);                                                           \
            return ::locr::fail(LOCR_ERR_CUDA, _b);                                                 \
        }                                                                                           \
    } while (0)

inline uint16_t f32_to_act(float f, int act_dtype) {
    if (act_dtype == LOCR_ACT_BF16) {
        __nv_bfloat16 b = __float2bfloat16_rn(f);
        uint16_t u;
        memcpy(&u, &b, 2);
        return u;
    }
    __half h = __float2half_rn(f);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
}
inline float act_to_f32(uint16_t u, int act_dtype) {
    if (act_dtype == LOCR_ACT_BF16) {
        __nv_bfloat16 b;
        memcpy(&b, &u, 2);
        return __bfloat162float(b);
    }
    __half h;
    memcpy(&h, &u, 2);
    return __half2float(h);
}

// RAII device buffer for the test hooks and one-off allocations.
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() {}
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    cudaError_t alloc(size_t n) {
        if (p) { cudaFree(p); p = nullptr; }
        bytes = n;
        return cudaMalloc(&p, n ? n : 16);
    }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace locr
