// Kernel-level test hooks behind the C ABI.  They drive exactly the kernels the product path uses, on host data.
#include <stdlib.h>

#include <vector>

#include "conv_tc.cuh"
#include "jpeg.cuh"
#include "png.cuh"
#include "nn_kernels.cuh"
#include "util.cuh"

namespace locr {

std::string& tls_error() {
    static thread_local std::string e;
    return e;
}
int fail(int code, const std::string& msg) {
    tls_error() = msg;
    return code;
}

}  // namespace locr

using namespace locr;

extern "C" {

LOCR_API const char* locr_version(void) { return "liblocr 0.1 (sm_100a)"; }

static int test_conv_impl(const locr_conv_desc* d, const float* x, const float* w, const float* bias,
                          const float* residual, float* y, float* y_pool) {
    if (d == nullptr || x == nullptr || w == nullptr || (y == nullptr && y_pool == nullptr))
        return fail(LOCR_ERR_INVALID, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(LOCR_ERR_CUDA, "no CUDA device: liblocr has no CPU fallback");
    const int S = d->stride_h;
    const int OH = (d->H + 2 * d->pad_h - d->dil_h * (d->KH - 1) - 1) / S + 1;
    const int OW = (d->W + 2 * d->pad_w - d->dil_w * (d->KW - 1) - 1) + 1;
    const int cout_pad = (d->Cout + 15) / 16 * 16;
    const size_t nx = (size_t)d->B * d->H * d->W * d->x_pitch;
    const size_t ktot = (size_t)d->KH * d->KW * d->Cin;
    const size_t nw = (size_t)cout_pad * ktot;
    const size_t npix = (size_t)d->B * OH * OW;
    const size_t ny = npix * d->y_pitch;

    std::vector<uint16_t> hx(nx), hw(nw, 0);
    for (size_t i = 0; i < nx; ++i) hx[i] = f32_to_act(x[i], d->act_dtype);
    for (size_t i = 0; i < (size_t)d->Cout * ktot; ++i) hw[i] = f32_to_act(w[i], d->act_dtype);
    std::vector<float> hb(cout_pad, 0.f);
    if (bias) for (int i = 0; i < d->Cout; ++i) hb[i] = bias[i];
    std::vector<uint16_t> hr;
    if (residual) {
        hr.resize(npix * d->Cout);
        for (size_t i = 0; i < hr.size(); ++i) hr[i] = f32_to_act(residual[i], d->act_dtype);
    }

    DevBuf dx, dw, db, dr, dy;
    LOCR_CUDA_OK(dx.alloc(nx * 2));
    LOCR_CUDA_OK(dw.alloc(nw * 2));
    LOCR_CUDA_OK(db.alloc(cout_pad * 4));
    LOCR_CUDA_OK(dy.alloc(ny * (d->out_fp32 ? 4 : 2)));
    LOCR_CUDA_OK(cudaMemcpy(dx.p, hx.data(), nx * 2, cudaMemcpyHostToDevice));
    LOCR_CUDA_OK(cudaMemcpy(dw.p, hw.data(), nw * 2, cudaMemcpyHostToDevice));
    LOCR_CUDA_OK(cudaMemcpy(db.p, hb.data(), cout_pad * 4, cudaMemcpyHostToDevice));
    LOCR_CUDA_OK(cudaMemset(dy.p, 0, ny * (d->out_fp32 ? 4 : 2)));
    if (residual) {
        LOCR_CUDA_OK(dr.alloc(hr.size() * 2));
        LOCR_CUDA_OK(cudaMemcpy(dr.p, hr.data(), hr.size() * 2, cudaMemcpyHostToDevice));
    }

    ConvCall c;
    c.x = dx.p; c.B = d->B; c.H = d->H; c.W = d->W; c.Cin = d->Cin; c.x_pitch = d->x_pitch;
    c.w = dw.p; c.Cout = d->Cout; c.Cout_pad = cout_pad;
    c.KH = d->KH; c.KW = d->KW; c.dil_h = d->dil_h; c.dil_w = d->dil_w; c.pad_h = d->pad_h; c.pad_w = d->pad_w;
    c.stride_h = S;
    c.y = dy.p; c.OH = OH; c.OW = OW; c.y_pitch = d->y_pitch; c.out_fp32 = d->out_fp32;
    c.bias = db.as<float>();
    c.residual = residual ? dr.p : nullptr; c.res_pitch = d->Cout;
    c.relu = d->relu; c.dtype = d->act_dtype; c.n_tile = d->n_tile;
    DevBuf dp;
    const size_t np_ = (size_t)d->B * (OH / 2) * (OW / 2) * d->Cout;
    if (y_pool != nullptr) {
        LOCR_CUDA_OK(dp.alloc(np_ * 2));
        LOCR_CUDA_OK(cudaMemset(dp.p, 0, np_ * 2));
        c.pool_y = dp.p;
        c.pool_pitch = d->Cout;
        c.skip_full = y == nullptr ? 1 : 0;
        if (y == nullptr) c.y = nullptr;
    }
    // like the engine: calls with at most 1024 output pixels get a split-K workspace (LOCR_TEST_SPLITK=0, read per call,
    // runs the same call unsplit)
    DevBuf dws;
    const char* sk = getenv("LOCR_TEST_SPLITK");
    if (npix <= 1024 && !(sk && atoi(sk) == 0)) {
        const size_t wsb = npix * 8 * (size_t)cout_pad * 4;
        LOCR_CUDA_OK(dws.alloc(wsb));
        LOCR_CUDA_OK(cudaMemset(dws.p, 0xff, wsb));      // NaN pattern: a partial sum that was never written shows
        c.splitk_ws = dws.p;
        c.splitk_ws_bytes = wsb;
    }
    char err[256] = {0};
    cudaError_t e = conv_tc_launch(c, 0, err, sizeof(err));
    if (e != cudaSuccess) return fail(e == cudaErrorInvalidValue ? LOCR_ERR_INVALID : LOCR_ERR_CUDA, err);
    LOCR_CUDA_OK(cudaDeviceSynchronize());
    if (y != nullptr) {
        if (d->out_fp32) {
            LOCR_CUDA_OK(cudaMemcpy(y, dy.p, ny * 4, cudaMemcpyDeviceToHost));
        } else {
            std::vector<uint16_t> hy(ny);
            LOCR_CUDA_OK(cudaMemcpy(hy.data(), dy.p, ny * 2, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < ny; ++i) y[i] = act_to_f32(hy[i], d->act_dtype);
        }
    }
    if (y_pool != nullptr) {
        std::vector<uint16_t> hy(np_);
        LOCR_CUDA_OK(cudaMemcpy(hy.data(), dp.p, np_ * 2, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < np_; ++i) y_pool[i] = act_to_f32(hy[i], d->act_dtype);
    }
    return LOCR_OK;
}

LOCR_API int locr_test_conv(const locr_conv_desc* d, const float* x, const float* w, const float* bias,
                            const float* residual, float* y) {
    return test_conv_impl(d, x, w, bias, residual, y, nullptr);
}

/* Same with the fused MaxPool2d(2, 2): y_pool [B, OH/2, OW/2, Cout] fp32; y may be NULL (pooled output only). */
LOCR_API int locr_test_conv_pool(const locr_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                                 float* y_pool) {
    if (y_pool == nullptr) return fail(LOCR_ERR_INVALID, "null argument");
    return test_conv_impl(d, x, w, bias, nullptr, y, y_pool);
}

/* Split-K convolutions (conv_tc.cuh splitk_ws) launched by this process so far. */
LOCR_API int64_t locr_test_splitk_calls(void) { return (int64_t)conv_tc_splitk_calls(); }

/* Times `iters` back-to-back launches of one conv layer on uninitialised (zeroed) device buffers. */
LOCR_API int locr_bench_conv(const locr_conv_desc* d, int iters, float* ms_per_iter) {
    const int S = d->stride_h;
    const int OH = (d->H + 2 * d->pad_h - d->dil_h * (d->KH - 1) - 1) / S + 1;
    const int OW = (d->W + 2 * d->pad_w - d->dil_w * (d->KW - 1) - 1) + 1;
    const int cout_pad = (d->Cout + 15) / 16 * 16;
    const size_t nx = (size_t)d->B * d->H * d->W * d->x_pitch;
    const size_t ktot = (size_t)d->KH * d->KW * d->Cin;
    const size_t nw = (size_t)cout_pad * ktot;
    const size_t ny = (size_t)d->B * OH * OW * d->y_pitch;
    DevBuf dx, dw, db, dy;
    LOCR_CUDA_OK(dx.alloc(nx * 2));
    LOCR_CUDA_OK(dw.alloc(nw * 2));
    LOCR_CUDA_OK(db.alloc(cout_pad * 4));
    LOCR_CUDA_OK(dy.alloc(ny * (d->out_fp32 ? 4 : 2)));
    LOCR_CUDA_OK(cudaMemset(dx.p, 0x11, nx * 2));
    LOCR_CUDA_OK(cudaMemset(dw.p, 0x11, nw * 2));
    LOCR_CUDA_OK(cudaMemset(db.p, 0, cout_pad * 4));
    ConvCall c;
    c.x = dx.p; c.B = d->B; c.H = d->H; c.W = d->W; c.Cin = d->Cin; c.x_pitch = d->x_pitch;
    c.w = dw.p; c.Cout = d->Cout; c.Cout_pad = cout_pad;
    c.KH = d->KH; c.KW = d->KW; c.dil_h = d->dil_h; c.dil_w = d->dil_w; c.pad_h = d->pad_h; c.pad_w = d->pad_w;
    c.stride_h = S;
    c.y = dy.p; c.OH = OH; c.OW = OW; c.y_pitch = d->y_pitch; c.out_fp32 = d->out_fp32;
    c.bias = db.as<float>();
    c.relu = d->relu; c.dtype = d->act_dtype; c.n_tile = d->n_tile;
    DevBuf dp;
    const char* penv = getenv("LOCR_BENCH_POOL");     // 1: fused 2x2 max-pool + full output, 2: pooled output only
    if (penv && atoi(penv) > 0 && !d->out_fp32) {
        LOCR_CUDA_OK(dp.alloc((size_t)d->B * (OH / 2) * (OW / 2) * d->Cout * 2));
        c.pool_y = dp.p;
        c.pool_pitch = d->Cout;
        c.skip_full = atoi(penv) == 2 ? 1 : 0;
    }
    DevBuf dws;
    const char* sk = getenv("LOCR_TEST_SPLITK");
    if ((size_t)d->B * OH * OW <= 1024 && !(sk && atoi(sk) == 0)) {
        const size_t wsb = (size_t)d->B * OH * OW * 8 * (size_t)cout_pad * 4;
        LOCR_CUDA_OK(dws.alloc(wsb));
        c.splitk_ws = dws.p;
        c.splitk_ws_bytes = wsb;
    }
    char err[256] = {0};
    const char* wenv = getenv("LOCR_BENCH_WARMUP");   // profiler runs: 0 warm-ups keep the capture to one launch per layer
    const int warm = wenv ? atoi(wenv) : 3;
    for (int i = 0; i < warm; ++i) {
        cudaError_t e = conv_tc_launch(c, 0, err, sizeof(err));
        if (e != cudaSuccess) return fail(LOCR_ERR_CUDA, err);
    }
    cudaEvent_t e0, e1;
    LOCR_CUDA_OK(cudaEventCreate(&e0));
    LOCR_CUDA_OK(cudaEventCreate(&e1));
    LOCR_CUDA_OK(cudaEventRecord(e0, 0));
    for (int i = 0; i < iters; ++i) conv_tc_launch(c, 0, err, sizeof(err));
    LOCR_CUDA_OK(cudaEventRecord(e1, 0));
    LOCR_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    LOCR_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_iter = ms / iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return LOCR_OK;
}

/* Host half of the JPEG reader alone (no GPU): see jpeg.cuh jpeg_host_coefficients. */
LOCR_API int locr_test_jpeg_coefficients(const uint8_t* data, int64_t nbytes, int16_t* out, int64_t capacity, int* info) {
    if (data == nullptr || nbytes <= 0 || info == nullptr) return fail(LOCR_ERR_INVALID, "null argument");
    std::string err;
    const int rc = jpeg_host_coefficients(data, (size_t)nbytes, out, (size_t)(capacity < 0 ? 0 : capacity), info, &err);
    return rc == LOCR_OK ? LOCR_OK : fail(rc, err);
}

/* Host half of the PNG reader alone (no GPU): see png.cuh png_host_scanlines. */
LOCR_API int locr_test_png_scanlines(const uint8_t* data, int64_t nbytes, uint8_t* out, int64_t capacity, int64_t* need) {
    if (data == nullptr || nbytes <= 0) return fail(LOCR_ERR_INVALID, "null argument");
    std::string err;
    size_t nd = 0;
    const int rc = png_host_scanlines(data, (size_t)nbytes, out, (size_t)(capacity < 0 ? 0 : capacity), &nd, &err);
    if (need) *need = (int64_t)nd;
    return rc == LOCR_OK ? LOCR_OK : fail(rc, err);
}

/* Experiments build only: the conv kernel's in-kernel timeline (see conv_tc.cuh); out [3*8192], counts [3]. */
LOCR_API int locr_conv_trace(unsigned long long* out, int* counts) {
    if (out == nullptr || counts == nullptr) return fail(LOCR_ERR_INVALID, "null argument");
    return conv_tc_trace_read(out, counts) == 0 ? LOCR_OK : fail(LOCR_ERR_CUDA, "trace read failed");
}

/* The BiLSTM recurrence kernel alone.  xproj [B][T][2048] fp32 and whh [2][1024][256] fp32 come in PyTorch's row order
 * (dir*1024 + gate*256 + unit); out [B][T][512] fp32.  iters > 0 additionally times `iters` launches (ms per launch). */
LOCR_API int locr_test_lstm(const float* xproj, const float* whh, int B, int T, int act_dtype, float* out, int iters,
                            float* ms_per_iter, int split) {
    if (xproj == nullptr || whh == nullptr || out == nullptr || B <= 0 || T <= 0) return fail(LOCR_ERR_INVALID, "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(LOCR_ERR_CUDA, "no CUDA device: liblocr has no CPU fallback");
    const size_t nx = (size_t)B * T * 2048;
    std::vector<float> xp(nx);
    std::vector<uint16_t> w16((size_t)2048 * 256);
    for (int d = 0; d < 2; ++d)
        for (int np = 0; np < 1024; ++np) {
            const int r = (np & 3) * 256 + (np >> 2);
            for (int k = 0; k < 256; ++k)
                w16[((size_t)d * 1024 + np) * 256 + k] = f32_to_act(whh[((size_t)d * 1024 + r) * 256 + k], act_dtype);
        }
    for (size_t row = 0; row < (size_t)B * T; ++row)
        for (int d = 0; d < 2; ++d)
            for (int np = 0; np < 1024; ++np)
                xp[row * 2048 + d * 1024 + np] = xproj[row * 2048 + d * 1024 + (np & 3) * 256 + (np >> 2)];
    DevBuf dx, dw, dy;
    const int pitch = split ? 1024 : 512;
    const size_t ny = (size_t)B * T * pitch;
    LOCR_CUDA_OK(dx.alloc(nx * 4));
    LOCR_CUDA_OK(dw.alloc(w16.size() * 2));
    LOCR_CUDA_OK(dy.alloc(ny * 2));
    LOCR_CUDA_OK(cudaMemcpy(dx.p, xp.data(), nx * 4, cudaMemcpyHostToDevice));
    LOCR_CUDA_OK(cudaMemcpy(dw.p, w16.data(), w16.size() * 2, cudaMemcpyHostToDevice));
    LOCR_CUDA_OK(cudaMemset(dy.p, 0xff, ny * 2));
    const int f16 = act_dtype == LOCR_ACT_F16 ? 1 : 0;
    cudaError_t e = launch_lstm_tc(dx.as<float>(), dw.p, dy.p, B, T, f16, 0, split);
    if (e != cudaSuccess) return fail(LOCR_ERR_CUDA, cudaGetErrorString(e));
    LOCR_CUDA_OK(cudaDeviceSynchronize());
    std::vector<uint16_t> hy(ny);
    LOCR_CUDA_OK(cudaMemcpy(hy.data(), dy.p, ny * 2, cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < (size_t)B * T; ++r)
        for (int j = 0; j < 512; ++j)
            out[r * 512 + j] = act_to_f32(hy[r * pitch + j], act_dtype) +
                               (split ? act_to_f32(hy[r * pitch + 512 + j], act_dtype) : 0.f);
    if (iters > 0 && ms_per_iter != nullptr) {
        cudaEvent_t e0, e1;
        LOCR_CUDA_OK(cudaEventCreate(&e0));
        LOCR_CUDA_OK(cudaEventCreate(&e1));
        LOCR_CUDA_OK(cudaEventRecord(e0, 0));
        for (int i = 0; i < iters; ++i) launch_lstm_tc(dx.as<float>(), dw.p, dy.p, B, T, f16, 0, split);
        LOCR_CUDA_OK(cudaEventRecord(e1, 0));
        LOCR_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0;
        LOCR_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        *ms_per_iter = ms / iters;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    return LOCR_OK;
}

/* The evaluation-loss kernels alone on host logits [n][26][C] (same kernels as locr_evaluate): greedy ids through the
 * decode kernel, then the CTC loss (head_attn = 0; targets concatenated, target_len [n]) or the attention cross entropy
 * (head_attn = 1; targets [n][targets_total / n], count [n] receives the counted steps). */
LOCR_API int locr_test_eval_loss(int head_attn, const float* logits, int n, int C, const int32_t* targets,
                                 const int32_t* target_len, int64_t targets_total, float* loss, int32_t* count,
                                 int32_t* correct) {
    if (logits == nullptr || targets == nullptr || target_len == nullptr || loss == nullptr || correct == nullptr ||
        n <= 0 || C <= 1 || targets_total < 0)
        return fail(LOCR_ERR_INVALID, "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(LOCR_ERR_CUDA, "no CUDA device: liblocr has no CPU fallback");
    std::vector<int32_t> off(n);
    int64_t tot = 0;
    for (int i = 0; i < n; ++i) {
        off[i] = (int32_t)tot;
        tot += target_len[i];
    }
    if (!head_attn && tot != targets_total) return fail(LOCR_ERR_INVALID, "target lengths do not add up");
    if (head_attn && (targets_total % n != 0 || targets_total / n < 2)) return fail(LOCR_ERR_INVALID, "bad target width");
    DevBuf dl, dt, dm, dloss, dids, dtext, dconf;
    LOCR_CUDA_OK(dl.alloc((size_t)n * 26 * C * 4));
    LOCR_CUDA_OK(dt.alloc((size_t)targets_total * 4 + 16));
    LOCR_CUDA_OK(dm.alloc((size_t)n * 5 * 4));
    LOCR_CUDA_OK(dloss.alloc((size_t)n * 4));
    LOCR_CUDA_OK(dids.alloc((size_t)n * 26 * 4));
    LOCR_CUDA_OK(dtext.alloc((size_t)n * 128));
    LOCR_CUDA_OK(dconf.alloc((size_t)n * 4));
    int32_t* m = dm.as<int32_t>();
    LOCR_CUDA_OK(cudaMemcpy(dl.p, logits, (size_t)n * 26 * C * 4, cudaMemcpyHostToDevice));
    if (targets_total) LOCR_CUDA_OK(cudaMemcpy(dt.p, targets, (size_t)targets_total * 4, cudaMemcpyHostToDevice));
    LOCR_CUDA_OK(cudaMemcpy(m, off.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    LOCR_CUDA_OK(cudaMemcpy(m + n, target_len, (size_t)n * 4, cudaMemcpyHostToDevice));
    LOCR_CUDA_OK(cudaMemset(m + 2 * n, 0, (size_t)n * 3 * 4));
    launch_decode(dl.as<float>(), n, C, head_attn, dids.as<int32_t>(), dtext.as<char>(), 128, m + 4 * n, dconf.as<float>(), 0);
    if (head_attn)
        launch_attn_ce(dl.as<float>(), n, C, dt.as<int32_t>(), (int)(targets_total / n), dids.as<int32_t>(),
                       dloss.as<float>(), m + 2 * n, m + 3 * n, 0);
    else
        launch_ctc_loss(dl.as<float>(), n, C, dt.as<int32_t>(), m, m + n, dids.as<int32_t>(), dloss.as<float>(), m + 3 * n, 0);
    LOCR_CUDA_OK(cudaGetLastError());
    LOCR_CUDA_OK(cudaDeviceSynchronize());
    LOCR_CUDA_OK(cudaMemcpy(loss, dloss.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    LOCR_CUDA_OK(cudaMemcpy(correct, m + 3 * n, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (count) LOCR_CUDA_OK(cudaMemcpy(count, m + 2 * n, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return LOCR_OK;
}

}  // extern "C"
