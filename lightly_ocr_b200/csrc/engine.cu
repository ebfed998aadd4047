#include "engine.cuh"

#include <math.h>

namespace locr {

namespace {

const char* kLoc = "Transformation.LocalizationNetwork.";
const char* kFe = "FeatureExtraction.ConvNet.";

const HostTensor* find(locr_handle* h, int model, const std::string& key) {
    auto it = h->host[model].find(key);
    return it == h->host[model].end() ? nullptr : &it->second;
}

template <typename T>
T* dev_upload(locr_handle* h, const std::vector<T>& v) {
    void* p = nullptr;
    if (cudaMalloc(&p, v.size() * sizeof(T) + 16) != cudaSuccess) return nullptr;
    if (cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    h->owned.push_back(p);
    return reinterpret_cast<T*>(p);
}

// Conv (+ optional BatchNorm, eval mode, eps 1e-5) -> weights with the BN scale folded in and a single fp32 bias.
// split: 0 = plain; 3 = split-precision input AND weights (K per tap [x_hi | x_lo | x_hi] against [w_hi | w_hi | w_lo]);
//        2 = split-precision weights only, plain 16-bit input (K per tap [x | x] against [w_hi | w_lo])
int fold_conv(locr_handle* h, int model, const std::string& prefix, const std::string& bn, bool direct,
              bool fold_image_std = false, int cin_pad = 0, int split = 0, bool window = false, bool im2col = false) {
    const HostTensor* w = find(h, model, prefix + ".weight");
    if (w == nullptr || (w->shape.size() != 4 && w->shape.size() != 2))
        return h->fail(LOCR_ERR_STATE, "missing or malformed tensor " + prefix + ".weight");
    const int cout = (int)w->shape[0], cin = (int)w->shape[1];
    const int kh = w->shape.size() == 4 ? (int)w->shape[2] : 1, kw = w->shape.size() == 4 ? (int)w->shape[3] : 1;
    const HostTensor* b = find(h, model, prefix + ".bias");
    std::vector<double> scale(cout, 1.0), shift(cout, 0.0);
    for (int n = 0; n < cout; ++n) shift[n] = b ? (double)b->data[n] : 0.0;
    if (!bn.empty()) {
        const HostTensor* g = find(h, model, bn + ".weight");
        const HostTensor* be = find(h, model, bn + ".bias");
        const HostTensor* mu = find(h, model, bn + ".running_mean");
        const HostTensor* var = find(h, model, bn + ".running_var");
        if (!g || !be || !mu || !var || g->numel() != cout)
            return h->fail(LOCR_ERR_STATE, "missing BatchNorm tensors for " + bn);
        for (int n = 0; n < cout; ++n) {
            const double s = (double)g->data[n] / sqrt((double)var->data[n] + 1e-5);
            scale[n] = s;
            shift[n] = (shift[n] - (double)mu->data[n]) * s + (double)be->data[n];
        }
    }
    ConvW cw;
    cw.cin = cin; cw.cin_real = cin; cw.cout = cout; cw.kh = kh; cw.kw = kw;
    cw.cout_pad = (cout + 15) / 16 * 16;
    std::vector<float> bias(cw.cout_pad, 0.f);
    for (int n = 0; n < cout; ++n) bias[n] = (float)shift[n];
    cw.bias = dev_upload(h, bias);
    const int taps = kh * kw;
    if (direct) {
        std::vector<float> w32((size_t)taps * cin * cout);
        for (int n = 0; n < cout; ++n)
            for (int c = 0; c < cin; ++c)
                for (int t = 0; t < taps; ++t)
                {
                    // u8 mode of the direct kernel: (x - mean)/std with 1/std folded into the weights (BGR order
                    // meets the RGB constants exactly as in the reference, imgproc.py:19-25)
                    static const double stdv[3] = {0.229 * 255.0, 0.224 * 255.0, 0.225 * 255.0};
                    const double istd = fold_image_std ? 1.0 / (double)(float)stdv[c] : 1.0;
                    w32[((size_t)t * cin + c) * cout + n] =
                        (float)(w->data[((size_t)n * cin + c) * taps + t] * scale[n] * istd);
                }
        cw.w32 = dev_upload(h, w32);
        if (!cw.w32) return h->fail(LOCR_ERR_CUDA, "weight upload failed");
    } else {
        if (split == 2) {
            if (cin % 32 != 0) return h->fail(LOCR_ERR_INVALID, prefix + ": split precision needs Cin % 32 == 0");
            const int K2 = 2 * cin;
            std::vector<uint16_t> w16((size_t)cw.cout_pad * taps * K2, 0);
            for (int n = 0; n < cout; ++n)
                for (int c = 0; c < cin; ++c)
                    for (int t = 0; t < taps; ++t) {
                        const float wv = (float)(w->data[((size_t)n * cin + c) * taps + t] * scale[n]);
                        const uint16_t hi = f32_to_act(wv, h->cfg.act_dtype);
                        uint16_t* row = &w16[((size_t)n * taps + t) * K2];
                        row[c] = hi;
                        row[cin + c] = f32_to_act(wv - act_to_f32(hi, h->cfg.act_dtype), h->cfg.act_dtype);
                    }
            cw.w = dev_upload(h, w16);
            if (!cw.w || !cw.bias) return h->fail(LOCR_ERR_CUDA, "weight upload failed");
            cw.cin = K2;
            cw.cin_wrap = cin;
            h->conv[prefix] = cw;
            return LOCR_OK;
        }
        if (split) {
            // split-precision layer: K channels per tap are [x_hi | x_lo | x_hi] against [w_hi | w_hi | w_lo], i.e.
            // x*w ~ x_hi*w_hi + x_lo*w_hi + x_hi*w_lo with ~22 significant bits on both operands
            if (cin % 32 != 0) return h->fail(LOCR_ERR_INVALID, prefix + ": split precision needs Cin % 32 == 0");
            const int K3 = 3 * cin;
            std::vector<uint16_t> w16((size_t)cw.cout_pad * taps * K3, 0);
            for (int n = 0; n < cout; ++n)
                for (int c = 0; c < cin; ++c)
                    for (int t = 0; t < taps; ++t) {
                        const float wv = (float)(w->data[((size_t)n * cin + c) * taps + t] * scale[n]);
                        const uint16_t hi = f32_to_act(wv, h->cfg.act_dtype);
                        const uint16_t lo = f32_to_act(wv - act_to_f32(hi, h->cfg.act_dtype), h->cfg.act_dtype);
                        uint16_t* row = &w16[((size_t)n * taps + t) * K3];
                        row[c] = hi;
                        row[cin + c] = hi;
                        row[2 * cin + c] = lo;
                    }
            cw.w = dev_upload(h, w16);
            if (!cw.w) return h->fail(LOCR_ERR_CUDA, "weight upload failed");
            cw.cin = K3;
            cw.cin_wrap = 2 * cin;
            if (!cw.bias) return h->fail(LOCR_ERR_CUDA, "bias upload failed");
            h->conv[prefix] = cw;
            return LOCR_OK;
        }
        if (im2col) {
            // first layer as an im2col GEMM (conv_tc.cuh: first_u8): K = 32 with k = (ky * 3 + kx) * cin + c, 27 used
            if (kh != 3 || kw != 3 || cin != 3) return h->fail(LOCR_ERR_INVALID, prefix + ": im2col layout is for the 3x3 / 3-channel first layer");
            const int K = 32;
            std::vector<uint16_t> w16((size_t)cw.cout_pad * K, 0);
            for (int n = 0; n < cout; ++n)
                for (int c = 0; c < cin; ++c)
                    for (int t = 0; t < 9; ++t)
                        w16[(size_t)n * K + t * cin + c] =
                            f32_to_act((float)(w->data[((size_t)n * cin + c) * 9 + t] * scale[n]), h->cfg.act_dtype);
            cw.cin = K; cw.cin_real = cin * 9; cw.kh = 1; cw.kw = 1;
            cw.w = dev_upload(h, w16);
            if (!cw.w || !cw.bias) return h->fail(LOCR_ERR_CUDA, "weight upload failed");
            h->conv[prefix] = cw;
            return LOCR_OK;
        }
        if (window) {
            // 3x3 conv as KH = 3 taps over a 4-pixel window (conv_tc.cuh: x_row_px): K per tap = 4 * cpp with
            // k = dx * cpp + c for the pixel x - 1 + dx; the 4th pixel and padded channels get zero weights
            const int cpp = cin_pad > 0 ? cin_pad : cin;
            if (kh != 3 || kw != 3 || (4 * cpp) % 32 != 0 || cin > cpp)
                return h->fail(LOCR_ERR_INVALID, prefix + ": window view needs a 3x3 conv with 8, 16 or 32 channels per pixel");
            const int K = 4 * cpp;
            std::vector<uint16_t> w16((size_t)cw.cout_pad * 3 * K, 0);
            for (int n = 0; n < cout; ++n)
                for (int c = 0; c < cin; ++c)
                    for (int ky = 0; ky < 3; ++ky)
                        for (int dx = 0; dx < 3; ++dx)
                            w16[((size_t)n * 3 + ky) * K + dx * cpp + c] = f32_to_act(
                                (float)(w->data[((size_t)n * cin + c) * 9 + ky * 3 + dx] * scale[n]), h->cfg.act_dtype);
            cw.cin = K; cw.cin_real = cin * 3; cw.kh = 3; cw.kw = 1; cw.window = 1;
            cw.w = dev_upload(h, w16);
            if (!cw.w || !cw.bias) return h->fail(LOCR_ERR_CUDA, "weight upload failed");
            h->conv[prefix] = cw;
            return LOCR_OK;
        }
        const int cp = cin_pad > 0 ? cin_pad : cin;   // input channels as laid out in memory (zero-padded)
        if (cp % 16 != 0) return h->fail(LOCR_ERR_INVALID, prefix + ": Cin must be a multiple of 16");
        std::vector<uint16_t> w16((size_t)cw.cout_pad * taps * cp, 0);
        for (int n = 0; n < cout; ++n)
            for (int c = 0; c < cin; ++c)
                for (int t = 0; t < taps; ++t)
                    w16[((size_t)n * taps + t) * cp + c] =
                        f32_to_act((float)(w->data[((size_t)n * cin + c) * taps + t] * scale[n]), h->cfg.act_dtype);
        cw.cin = cp;
        cw.w = dev_upload(h, w16);
        if (!cw.w) return h->fail(LOCR_ERR_CUDA, "weight upload failed");
    }
    if (!cw.bias) return h->fail(LOCR_ERR_CUDA, "bias upload failed");
    h->conv[prefix] = cw;
    return LOCR_OK;
}

float* upload_f32(locr_handle* h, const std::string& name, const std::vector<float>& v) {
    float* p = dev_upload(h, v);
    h->f32[name] = p;
    return p;
}

// [rows][cols] row-major -> [cols][rows]
std::vector<float> transpose(const std::vector<float>& a, int rows, int cols) {
    std::vector<float> t((size_t)rows * cols);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) t[(size_t)c * rows + r] = a[(size_t)r * cols + c];
    return t;
}

void dbg(locr_handle* h, const std::string& name, const void* p, int kind, std::vector<int64_t> shape, long pitch) {
    DebugTensor d;
    d.p = p; d.kind = kind; d.shape = std::move(shape); d.pitch = pitch;
    h->dbg[name] = d;
}

// split-K workspace: 1024 pixels x 8 slices x 1024 channels of fp32 partial sums at most
constexpr long kSplitKMaxPixels = 1024;
constexpr size_t kSplitKWorkspace = (size_t)kSplitKMaxPixels * 8 * 1024 * 4;

struct Ctx {
    locr_handle* h;
    int rc = LOCR_OK;
    // fused MaxPool2d(2, 2) request for the NEXT tc() call (consumed by it)
    void* pool_y = nullptr;
    long pool_pitch = 0;
    int pool_only = 0;
    void pool(void* y, long pitch, int only) { pool_y = y; pool_pitch = pitch; pool_only = only; }
    // row-padded input / output (pixels per memory row) for the NEXT tc() call (consumed by it)
    long x_row_px = 0, y_row_px = 0;
    void rows(long xr, long yr) { x_row_px = xr; y_row_px = yr; }
    // fused first layer (conv_tc.cuh first_u8) for the NEXT tc() call (consumed by it)
    const uint8_t* first_u8 = nullptr;
    int first_h = 0, first_w = 0;
    long first_row = 0, first_img = 0;
    void first(const uint8_t* p, int ih, int iw, long row, long img) { first_u8 = p; first_h = ih; first_w = iw; first_row = row; first_img = img; }
    // split-precision residual of the NEXT tc() call (consumed by it): offset of the lo halves in elements
    long res_lo_off = 0;
    void res_split(long off) { res_lo_off = off; }
    // fused 1x1 tail request for the NEXT tc() call (consumed by it)
    const float* tail_w = nullptr;
    float* tail_out = nullptr;
    void tail(const float* w, float* out) { tail_w = w; tail_out = out; }
    // y = act(conv(x)) through the tensor-core kernel.  Shapes are those of the INPUT; returns output dims.
    void tc(const std::string& layer, const void* x, int B, int H, int W, long x_pitch, void* y, long y_pitch, int relu,
            int pad_h, int pad_w, int dil = 1, int stride_h = 1, int out_fp32 = 0, const void* res = nullptr,
            long res_pitch = 0, int split_out = 0) {
        if (rc != LOCR_OK) return;
        auto it = h->conv.find(layer);
        if (it == h->conv.end() || it->second.w == nullptr) {
            rc = h->fail(LOCR_ERR_STATE, "layer not finalized: " + layer);
            return;
        }
        const ConvW& cw = it->second;
        ConvCall c;
        c.x = x; c.B = B; c.H = H; c.W = W; c.Cin = cw.cin; c.x_pitch = x_pitch;
        c.w = cw.w; c.Cout = cw.cout; c.Cout_pad = cw.cout_pad;
        c.KH = cw.kh; c.KW = cw.kw; c.dil_h = dil; c.dil_w = dil; c.pad_h = pad_h; c.pad_w = pad_w;
        c.stride_h = stride_h;
        c.OH = (H + 2 * pad_h - dil * (cw.kh - 1) - 1) / stride_h + 1;
        c.OW = (W + 2 * pad_w - dil * (cw.kw - 1) - 1) + 1;
        c.y = y; c.y_pitch = y_pitch; c.out_fp32 = out_fp32;
        c.bias = cw.bias; c.residual = res; c.res_pitch = res_pitch; c.relu = relu;
        c.res_lo_off = res_lo_off;
        res_lo_off = 0;
        c.first_u8 = first_u8; c.img_h = first_h; c.img_w = first_w; c.img_row_stride = first_row; c.img_stride = first_img;
        first_u8 = nullptr;
        c.dtype = h->cfg.act_dtype == LOCR_ACT_F16 ? ACT_F16 : ACT_BF16;
        c.cin_wrap = cw.cin_wrap;
        c.split_out = split_out;
        c.pool_y = pool_y; c.pool_pitch = pool_pitch; c.skip_full = pool_only;
        pool_y = nullptr; pool_pitch = 0; pool_only = 0;
        c.tail_w = tail_w; c.tail_out = tail_out;
        tail_w = nullptr; tail_out = nullptr;
        c.x_row_px = x_row_px; c.y_row_px = y_row_px;
        x_row_px = 0; y_row_px = 0;
        // single crops / a handful of them: deep layers split their K range over the machine (conv_tc.cuh splitk_ws)
        if ((long)B * c.OH * c.OW <= kSplitKMaxPixels && !out_fp32 && !split_out && cw.cin_wrap == 0 && c.pool_y == nullptr &&
            c.tail_out == nullptr && cw.cin % 64 == 0 && cw.kh * cw.kw * (cw.cin / 64) >= 8) {
            c.splitk_ws = engine_buffer(h, "splitk.ws", kSplitKWorkspace);     // null: the call simply runs unsplit
            c.splitk_ws_bytes = c.splitk_ws != nullptr ? kSplitKWorkspace : 0;
        }
        char err[256] = {0};
        cudaError_t e;
        const std::string shown = layer.substr(0, layer.find('#'));   // "layer#variant" is reported under the layer's name
        {
            ProfScope ps(h, shown, 2.0 * B * c.OH * c.OW * (double)cw.cout * cw.cin_real * cw.kh * cw.kw, true);
            e = conv_tc_launch(c, h->stream, err, sizeof(err));
        }
        h->launches++;
        if (e != cudaSuccess) rc = h->fail(LOCR_ERR_CUDA, layer + ": " + err);
        if (h->audit && e == cudaSuccess && !out_fp32 && c.tail_out == nullptr && h->audit_slots != nullptr &&
            (int)h->audit_names.size() < kAuditSlots) {
            // range audit: largest |activation| this layer stored (pooled tensor when only that one is written; a
            // split-precision tensor is judged by its hi halves; row pads are zero and do not matter)
            const bool pooled = c.pool_y != nullptr && c.skip_full;
            const void* t = pooled ? c.pool_y : y;
            const long pitch = pooled ? c.pool_pitch : y_pitch;
            const long rows = pooled ? (long)B * (c.OH / 2) * (c.OW / 2)
                                     : (long)B * c.OH * (c.y_row_px > 0 ? c.y_row_px : c.OW);
            if (t != nullptr) {
                launch_absmax(t, rows, cw.cout / 8 * 8, pitch, h->is_f16(), h->audit_slots + h->audit_names.size(), h->stream);
                h->audit_names.push_back(shown);
                h->launches++;
            }
        }
    }
    // Row-padded buffer whose pad pixels must read as zero: zero-filled whenever it is (re)allocated or its size changes
    // (the producers only ever write the interior).
    void* zbuf(const std::string& name, size_t bytes) {
        void* p = buf(name, bytes);
        if (p == nullptr) return p;
        auto& z = h->zeroed[name];
        if (z.first != p || z.second != bytes) {
            if (cudaMemsetAsync(p, 0, bytes, h->stream) != cudaSuccess && rc == LOCR_OK)
                rc = h->fail(LOCR_ERR_CUDA, "memset failed for " + name);
            z.first = p; z.second = bytes;
        }
        return p;
    }
    // Same for a buffer whose per-item geometry never changes (only the item count does): the part that was zero-filled
    // before keeps its zero pads, so only the newly covered tail is cleared.
    void* zbuf_grow(const std::string& name, size_t bytes) {
        void* p = buf(name, bytes);
        if (p == nullptr) return p;
        auto& z = h->zeroed[name];
        if (z.first != p) { z.first = p; z.second = 0; }
        if (bytes > z.second) {
            if (cudaMemsetAsync((char*)p + z.second, 0, bytes - z.second, h->stream) != cudaSuccess && rc == LOCR_OK)
                rc = h->fail(LOCR_ERR_CUDA, "memset failed for " + name);
            z.second = bytes;
        }
        return p;
    }
    void* buf(const std::string& name, size_t bytes) {
        void* p = engine_buffer(h, name, bytes);
        if (p == nullptr && rc == LOCR_OK) rc = h->fail(LOCR_ERR_CUDA, "device allocation failed for " + name);
        return p;
    }
};

}  // namespace

void* engine_buffer(locr_handle* h, const std::string& name, size_t bytes) {
    auto& e = h->buffers[name];
    if (e.first != nullptr && e.second >= bytes) return e.first;
    if (e.first != nullptr) {
        cudaStreamSynchronize(h->stream);
        cudaFree(e.first);
        e.first = nullptr;
    }
    h->zeroed.erase(name);   // a fresh allocation (possibly at the old address) holds no zero pads yet
    const size_t want = bytes + bytes / 8 + 256;
    if (cudaMalloc(&e.first, want) != cudaSuccess) {
        e.first = nullptr;
        e.second = 0;
        return nullptr;
    }
    e.second = want;
    return e.first;
}

void engine_release(locr_handle* h, const std::string& name) {
    auto it = h->buffers.find(name);
    if (it == h->buffers.end()) return;
    if (it->second.first != nullptr) {
        cudaStreamSynchronize(h->stream);
        cudaFree(it->second.first);
    }
    h->buffers.erase(it);
    h->zeroed.erase(name);
}

// ------------------------------------------------------------------------------------------------ CRAFT
// LOCR_FIRST_CH8=0 goes back to the 16-channel input tensor of the first CRAFT layer (A/B runs).
static bool first_c8() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("LOCR_FIRST_CH8"); v = e ? atoi(e) : 1; }
    return v != 0;
}

static const char* kCraftBn[][2] = {
    {"basenet.slice1.3", "basenet.slice1.4"},   {"basenet.slice1.7", "basenet.slice1.8"},
    {"basenet.slice1.10", "basenet.slice1.11"}, {"basenet.slice2.14", "basenet.slice2.15"},
    {"basenet.slice2.17", "basenet.slice2.18"}, {"basenet.slice3.20", "basenet.slice3.21"},
    {"basenet.slice3.24", "basenet.slice3.25"}, {"basenet.slice3.27", "basenet.slice3.28"},
    {"basenet.slice4.30", "basenet.slice4.31"}, {"basenet.slice4.34", "basenet.slice4.35"},
    {"basenet.slice4.37", "basenet.slice4.38"}, {"basenet.slice5.1", ""},
    {"basenet.slice5.2", ""},                   {"upconv1.conv.0", "upconv1.conv.1"},
    {"upconv1.conv.3", "upconv1.conv.4"},       {"upconv2.conv.0", "upconv2.conv.1"},
    {"upconv2.conv.3", "upconv2.conv.4"},       {"upconv3.conv.0", "upconv3.conv.1"},
    {"upconv3.conv.3", "upconv3.conv.4"},       {"upconv4.conv.0", "upconv4.conv.1"},
    {"upconv4.conv.3", "upconv4.conv.4"},       {"conv_cls.0", ""},
    {"conv_cls.2", ""},                         {"conv_cls.4", ""},
    {"conv_cls.6", ""},                         {"conv_cls.8", ""},
};

int engine_finalize_craft(locr_handle* h) {
    int rc = fold_conv(h, LOCR_MODEL_CRAFT, "basenet.slice1.0", "basenet.slice1.1", false, false, 16, false, true);
    if (rc != LOCR_OK) return rc;
    // the same layer over an 8-channel input tensor (16-byte pixels: half the pre-processing write and half the A traffic)
    h->conv["basenet.slice1.0#w"] = h->conv["basenet.slice1.0"];
    rc = fold_conv(h, LOCR_MODEL_CRAFT, "basenet.slice1.0", "basenet.slice1.1", false, false, 8, false, true);
    if (rc != LOCR_OK) return rc;
    h->conv["basenet.slice1.0#c8"] = h->conv["basenet.slice1.0"];
    // the same layer as an im2col GEMM fed straight from the uint8 image (experiment, LOCR_FIRST_FUSED=1)
    rc = fold_conv(h, LOCR_MODEL_CRAFT, "basenet.slice1.0", "basenet.slice1.1", false, false, 0, 0, false, true);
    if (rc != LOCR_OK) return rc;
    h->conv["basenet.slice1.0#im2col"] = h->conv["basenet.slice1.0"];
    h->conv["basenet.slice1.0"] = h->conv["basenet.slice1.0#w"];
    for (auto& e : kCraftBn) {
        if (std::string(e[0]) == "conv_cls.6" || std::string(e[0]) == "conv_cls.8") continue;   // fused tail, fp32
        const bool window = std::string(e[0]) == "conv_cls.0" || std::string(e[0]) == "conv_cls.2" ||
                            std::string(e[0]) == "conv_cls.4";
        rc = fold_conv(h, LOCR_MODEL_CRAFT, e[0], e[1], false, false, 0, false, window);
        if (rc != LOCR_OK) return rc;
        if (window) {
            // "#w": packed for the 4-pixel window view (LOCR_HEAD_HALO=0); the plain layout serves the haloed-patch form
            h->conv[std::string(e[0]) + "#w"] = h->conv[e[0]];
            rc = fold_conv(h, LOCR_MODEL_CRAFT, e[0], e[1], false, false, 0, false, false);
            if (rc != LOCR_OK) return rc;
        }
    }
    {
        const HostTensor* w6 = find(h, LOCR_MODEL_CRAFT, "conv_cls.6.weight");
        const HostTensor* b6 = find(h, LOCR_MODEL_CRAFT, "conv_cls.6.bias");
        const HostTensor* w8 = find(h, LOCR_MODEL_CRAFT, "conv_cls.8.weight");
        const HostTensor* b8 = find(h, LOCR_MODEL_CRAFT, "conv_cls.8.bias");
        if (!w6 || !b6 || !w8 || !b8 || w6->numel() != 256 || b6->numel() != 16 || w8->numel() != 32 || b8->numel() != 2)
            return h->fail(LOCR_ERR_STATE, "missing or malformed conv_cls.6 / conv_cls.8 tensors");
        std::vector<float> tail;
        tail.insert(tail.end(), w6->data.begin(), w6->data.end());
        tail.insert(tail.end(), b6->data.begin(), b6->data.end());
        tail.insert(tail.end(), w8->data.begin(), w8->data.end());
        tail.insert(tail.end(), b8->data.begin(), b8->data.end());
        if (upload_f32(h, "craft.cls_tail", tail) == nullptr) return h->fail(LOCR_ERR_CUDA, "weight upload failed");
    }
    h->host[LOCR_MODEL_CRAFT].clear();
    h->ready[LOCR_MODEL_CRAFT] = true;
    return LOCR_OK;
}

// VGG_UNet.forward (reference ocr/model.py:39-61) over vgg16_bn.forward (ocr/modules/vgg_bn.py:69-82).
// Concatenations never materialise: producers write straight into channel slices of the concat buffers.
int engine_craft_forward(locr_handle* h, const uint8_t* d_images, int B, int img_h, int img_w, int H, int W,
                         float** score) {
    if (!h->ready[LOCR_MODEL_CRAFT]) return h->fail(LOCR_ERR_STATE, "CRAFT weights not finalized");
    if (H % 32 != 0 || W % 32 != 0 || img_h > H || img_w > W) return h->fail(LOCR_ERR_INVALID, "bad canvas size");
    if ((long)B * H * W * 8 >= (1L << 31)) return h->fail(LOCR_ERR_CAPACITY, "CRAFT batch too large for 32-bit indices");
    Ctx c{h};
    const int f16 = h->is_f16();
    cudaStream_t s = h->stream;
    const size_t px = (size_t)B * H * W;
    void* a0 = c.buf("a0", px * 64 * 2);
    void* p1 = c.buf("p1", px / 4 * 64 * 2);
    void* a2 = c.buf("a2", px / 4 * 128 * 2);
    uint16_t* cat4 = (uint16_t*)c.buf("cat4", px / 4 * 192 * 2);
    void* p2 = c.buf("p2", px / 16 * 128 * 2);
    void* a3 = c.buf("a3", px / 16 * 256 * 2);
    uint16_t* cat3 = (uint16_t*)c.buf("cat3", px / 16 * 384 * 2);
    void* p3 = c.buf("p3", px / 64 * 256 * 2);
    void* a5 = c.buf("a5", px / 64 * 512 * 2);
    uint16_t* cat2 = (uint16_t*)c.buf("cat2", px / 64 * 768 * 2);
    void* p4 = c.buf("p4", px / 256 * 512 * 2);
    void* a7 = c.buf("a7", px / 256 * 512 * 2);
    uint16_t* cat1 = (uint16_t*)c.buf("cat1", px / 256 * 1536 * 2);
    void* p5 = c.buf("p5", px / 256 * 512 * 2);
    void* a8 = c.buf("a8", px / 256 * 1024 * 2);
    void* u1a = c.buf("u1a", px / 256 * 512 * 2);
    void* y1 = c.buf("y1", px / 256 * 256 * 2);
    void* u2a = c.buf("u2a", px / 64 * 256 * 2);
    void* y2 = c.buf("y2", px / 64 * 128 * 2);
    void* u3a = c.buf("u3a", px / 16 * 128 * 2);
    void* y3 = c.buf("y3", px / 16 * 64 * 2);
    void* u4a = c.buf("u4a", px / 4 * 64 * 2);
    // inputs of the 32-channel 3x3 head convs: rows padded to W/2 + 3 pixels (zero pixel left, two right), see conv_tc.cuh
    // (LOCR_HEAD_HALO=0, the round-1 form; by default these layers take the haloed-patch form of conv_tc and plain tensors)
    static int head_halo = -1;
    if (head_halo < 0) { const char* e = getenv("LOCR_HEAD_HALO"); head_halo = e ? atoi(e) : 1; }
    const long W2p = W / 2 + 3;
    const size_t padded32 = (size_t)B * (H / 2) * W2p * 32 * 2, plain32 = (size_t)B * (H / 2) * (W / 2) * 32 * 2;
    uint16_t* feat = head_halo ? (uint16_t*)c.buf("feature", plain32) : (uint16_t*)c.zbuf("feature_p", padded32);
    uint16_t* c0 = head_halo ? (uint16_t*)c.buf("c0", plain32) : (uint16_t*)c.zbuf("c0_p", padded32);
    uint16_t* c2 = head_halo ? (uint16_t*)c.buf("c2", plain32) : (uint16_t*)c.zbuf("c2_p", padded32);
    float* sc = (float*)c.buf("score", px / 4 * 2 * 4);
    if (c.rc != LOCR_OK) return c.rc;
    const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4, H8 = H / 8, W8 = W / 8, H16 = H / 16, W16 = W / 16;

    static int first_fused = -1;
    if (first_fused < 0) { const char* e = getenv("LOCR_FIRST_FUSED"); first_fused = e ? atoi(e) : 0; }
    if (first_fused) {
        // EXPERIMENT (off by default, LOCR_FIRST_FUSED=1): normalizeMeanVariance + basenet.slice1.0 + BN + ReLU in one
        // kernel, straight from the uint8 image.  Measured on B200: 1.15 ms per 8 canvases against 0.47 + 0.11 ms for
        // preproc + the window-view layer - four producer warps gathering 27 bytes per pixel from global memory are
        // latency-bound (~3300 cycles per 256-pixel tile); see DESIGN.md "measured and rejected".
        c.first(d_images, img_h, img_w, (long)img_w * 3, (long)img_h * img_w * 3);
        c.tc("basenet.slice1.0#im2col", d_images, B, H, W, 32, a0, 64, 1, 0, 0);
    } else if (first_c8()) {
        // 8 channels per pixel (3 used): the 4-pixel window is a 64-byte row (K = 32 per vertical tap)
        void* x8 = c.buf("x16", (size_t)B * H * (W + 3) * 8 * 2);   // rows padded: [zero | W pixels | zero zero]
        if (c.rc != LOCR_OK) return c.rc;
        { ProfScope ps_(h, "preproc_nhwc16", 0, false); launch_preproc_nhwc16(d_images, B, H, W, img_h, img_w, (long)img_w * 3, (long)img_h * img_w * 3, x8, f16, s, 8); }
        h->launches++;
        c.rows(W + 3, 0);
        c.tc("basenet.slice1.0#c8", x8, B, H, W, 8, a0, 64, 1, 1, 0);
    } else {
        void* x16 = c.buf("x16", (size_t)B * H * (W + 3) * 16 * 2);   // rows padded: [zero | W pixels | zero zero]
        if (c.rc != LOCR_OK) return c.rc;
        { ProfScope ps_(h, "preproc_nhwc16", 0, false); launch_preproc_nhwc16(d_images, B, H, W, img_h, img_w, (long)img_w * 3, (long)img_h * img_w * 3, x16, f16, s); }
        h->launches++;
        c.rows(W + 3, 0);
        c.tc("basenet.slice1.0", x16, B, H, W, 16, a0, 64, 1, 1, 0);   // 4-pixel window view: KW taps are one TMA row
    }
    c.pool(p1, 64, 1);   // MaxPool2d(2, 2) fused into the epilogue; the full-resolution tensor is never needed
    c.tc("basenet.slice1.3", a0, B, H, W, 64, nullptr, 64, 1, 1, 1);
    c.tc("basenet.slice1.7", p1, B, H2, W2, 64, a2, 128, 1, 1, 1);
    // relu2_2: the reference's slice ends on the BN, the next slice's in-place ReLU rectifies the tap as well
    c.pool(p2, 128, 0);
    c.tc("basenet.slice1.10", a2, B, H2, W2, 128, cat4 + 64, 192, 1, 1, 1);
    c.tc("basenet.slice2.14", p2, B, H4, W4, 128, a3, 256, 1, 1, 1);
    c.tc("basenet.slice2.17", a3, B, H4, W4, 256, cat3 + 128, 384, 1, 1, 1);               // relu3_2
    c.pool(p3, 256, 1);
    c.tc("basenet.slice3.20", cat3 + 128, B, H4, W4, 384, nullptr, 256, 1, 1, 1);
    c.tc("basenet.slice3.24", p3, B, H8, W8, 256, a5, 512, 1, 1, 1);
    c.tc("basenet.slice3.27", a5, B, H8, W8, 512, cat2 + 256, 768, 1, 1, 1);               // relu4_3
    c.pool(p4, 512, 1);
    c.tc("basenet.slice4.30", cat2 + 256, B, H8, W8, 768, nullptr, 512, 1, 1, 1);
    c.tc("basenet.slice4.34", p4, B, H16, W16, 512, a7, 512, 1, 1, 1);
    // relu5_3 keeps its negatives: slice5 starts with a non-in-place max-pool (vgg_bn.py:54)
    c.tc("basenet.slice4.37", a7, B, H16, W16, 512, cat1 + 1024, 1536, 0, 1, 1);
    { ProfScope ps_(h, "maxpool.craft5_3x3", 0, false); launch_maxpool(cat1 + 1024, 1536, B, H16, W16, 512, p5, 512, 3, 3, 1, 1, 1, 1, f16, s); }
    c.tc("basenet.slice5.1", p5, B, H16, W16, 512, a8, 1024, 0, 6, 6, 6);
    c.tc("basenet.slice5.2", a8, B, H16, W16, 1024, cat1, 1536, 0, 0, 0);                  // fc7
    c.tc("upconv1.conv.0", cat1, B, H16, W16, 1536, u1a, 512, 1, 0, 0);
    c.tc("upconv1.conv.3", u1a, B, H16, W16, 512, y1, 256, 1, 1, 1);
    { ProfScope ps_(h, "upsample.1", 0, false); launch_upsample2x(y1, 256, B, H16, W16, 256, cat2, 768, f16, s); }
    c.tc("upconv2.conv.0", cat2, B, H8, W8, 768, u2a, 256, 1, 0, 0);
    c.tc("upconv2.conv.3", u2a, B, H8, W8, 256, y2, 128, 1, 1, 1);
    { ProfScope ps_(h, "upsample.2", 0, false); launch_upsample2x(y2, 128, B, H8, W8, 128, cat3, 384, f16, s); }
    c.tc("upconv3.conv.0", cat3, B, H4, W4, 384, u3a, 128, 1, 0, 0);
    c.tc("upconv3.conv.3", u3a, B, H4, W4, 128, y3, 64, 1, 1, 1);
    { ProfScope ps_(h, "upsample.3", 0, false); launch_upsample2x(y3, 64, B, H4, W4, 64, cat4, 192, f16, s); }
    c.tc("upconv4.conv.0", cat4, B, H2, W2, 192, u4a, 64, 1, 0, 0);
    if (head_halo) {
        // 3x3 layers with <= 64 input and <= 64 output channels: one haloed patch per 16 x 16 tile (conv_tc.cu halo mode)
        c.tc("upconv4.conv.3", u4a, B, H2, W2, 64, feat, 32, 1, 1, 1);
        c.tc("conv_cls.0", feat, B, H2, W2, 32, c0, 32, 1, 1, 1);
        c.tc("conv_cls.2", c0, B, H2, W2, 32, c2, 32, 1, 1, 1);
        // conv_cls.4 + ReLU with conv_cls.6 + ReLU + conv_cls.8 (both 1x1) applied in the epilogue registers (fp32)
        c.tail(h->f32["craft.cls_tail"], sc);
        c.tc("conv_cls.4", c2, B, H2, W2, 32, nullptr, 16, 1, 1, 1);
    } else {
        c.rows(0, W2p);
        c.tc("upconv4.conv.3", u4a, B, H2, W2, 64, feat + 32, 32, 1, 1, 1);
        c.rows(W2p, W2p);
        c.tc("conv_cls.0#w", feat, B, H2, W2, 32, c0 + 32, 32, 1, 1, 0);
        c.rows(W2p, W2p);
        c.tc("conv_cls.2#w", c0, B, H2, W2, 32, c2 + 32, 32, 1, 1, 0);
        c.rows(W2p, 0);
        c.tail(h->f32["craft.cls_tail"], sc);
        c.tc("conv_cls.4#w", c2, B, H2, W2, 32, nullptr, 16, 1, 1, 0);
    }
    h->launches += 4;  // 1 max-pool (3x3 s1) + 3 up-samplings
    if (c.rc != LOCR_OK) return c.rc;
    LOCR_CUDA_OK(cudaGetLastError());
    dbg(h, "slice1.0", a0, 0, {B, H, W, 64}, 64);
    dbg(h, "relu2_2", cat4 + 64, 0, {B, H2, W2, 128}, 192);
    dbg(h, "relu3_2", cat3 + 128, 0, {B, H4, W4, 256}, 384);
    dbg(h, "relu4_3", cat2 + 256, 0, {B, H8, W8, 512}, 768);
    dbg(h, "relu5_3", cat1 + 1024, 0, {B, H16, W16, 512}, 1536);
    dbg(h, "fc7", cat1, 0, {B, H16, W16, 1024}, 1536);
    if (head_halo) dbg(h, "feature", feat, 0, {B, H2, W2, 32}, 32);
    else dbg(h, "feature", feat, 0, {B, H2, (int)W2p, 32}, 32);   // row-padded: columns 1 .. W/2 hold the tensor
    dbg(h, "score", sc, 1, {B, H2, W2, 2}, 2);
    *score = sc;
    return LOCR_OK;
}

// ------------------------------------------------------------------------------------------------ CRNN
int engine_finalize_crnn(locr_handle* h) {
    const int M = LOCR_MODEL_CRNN;
    const std::string loc = kLoc, fe = kFe;
    int rc;
    if ((rc = fold_conv(h, M, loc + "conv.0", loc + "conv.1", true))) return rc;
    // The localisation network runs in split precision (hi + lo 16-bit pairs, ~22 bits): its output moves the TPS
    // sampling grid, and a 1e-4 fiducial error of plain 16-bit storage becomes a 1e-2 error of the rectified crop.
    if ((rc = fold_conv(h, M, loc + "conv.4", loc + "conv.5", false, false, 0, 3))) return rc;
    if ((rc = fold_conv(h, M, loc + "conv.8", loc + "conv.9", false, false, 0, 3))) return rc;
    if ((rc = fold_conv(h, M, loc + "conv.12", loc + "conv.13", false, false, 0, 3))) return rc;
    // LOCR_PREC_EXACT: the whole recogniser in split precision (see include/locr.h); F folds accordingly from here on
    const int X = h->exact();
    auto F = [&](const std::string& p, const std::string& bn, bool direct = false) {
        return fold_conv(h, M, p, bn, direct, false, 0, (X && !direct) ? 3 : 0);
    };
    if ((rc = F(fe + "conv0_1", fe + "bn0_1", true))) return rc;
    if (X) { if ((rc = F(fe + "conv0_2", fe + "bn0_2"))) return rc; }
    else if ((rc = fold_conv(h, M, fe + "conv0_2", fe + "bn0_2", false, false, 0, 0, true))) return rc;   // window view
    const int nblocks[5] = {0, 1, 2, 5, 3};
    for (int l = 1; l <= 4; ++l) {
        for (int i = 0; i < nblocks[l]; ++i) {
            const std::string p = fe + "layer" + std::to_string(l) + "." + std::to_string(i) + ".";
            if ((rc = F(p + "conv1", p + "bn1"))) return rc;
            if ((rc = F(p + "conv2", p + "bn2"))) return rc;
            if (find(h, M, p + "downsample.0.weight") != nullptr)
                if ((rc = F(p + "downsample.0", p + "downsample.1"))) return rc;
        }
    }
    if ((rc = F(fe + "conv1", fe + "bn1"))) return rc;
    if ((rc = F(fe + "conv2", fe + "bn2"))) return rc;
    if ((rc = F(fe + "conv3", fe + "bn3"))) return rc;
    if ((rc = F(fe + "conv4_1", fe + "bn4_1"))) return rc;
    if ((rc = F(fe + "conv4_2", fe + "bn4_2"))) return rc;

    // localisation FC layers (fp32, transposed for coalesced reads) and the TPS buffers
    {
        const HostTensor* w1 = find(h, M, loc + "localization_fc1.0.weight");
        const HostTensor* b1 = find(h, M, loc + "localization_fc1.0.bias");
        const HostTensor* w2 = find(h, M, loc + "localization_fc2.weight");
        const HostTensor* b2 = find(h, M, loc + "localization_fc2.bias");
        const HostTensor* inv = find(h, M, "Transformation.GridGenerator.inv_delta_C");
        const HostTensor* ph = find(h, M, "Transformation.GridGenerator.P_hat");
        if (!w1 || !b1 || !w2 || !b2 || !inv || !ph || w1->numel() != 256 * 512 || w2->numel() != 40 * 256 ||
            inv->numel() != 23 * 23 || ph->numel() != 3200 * 23)
            return h->fail(LOCR_ERR_STATE, "missing or malformed TPS tensors");
        upload_f32(h, "loc.w1t", transpose(w1->data, 256, 512));
        upload_f32(h, "loc.b1", b1->data);
        upload_f32(h, "loc.w2t", transpose(w2->data, 40, 256));
        upload_f32(h, "loc.b2", b2->data);
        upload_f32(h, "tps.inv", inv->data);
        upload_f32(h, "tps.phat_t", transpose(ph->data, 3200, 23));
    }
    // BiLSTMs: stacked input projections as 1x1 "convs" (N = 2 dirs x 4 gates x 256), recurrent weights transposed
    for (int l = 0; l < 2; ++l) {
        const std::string p = "SequenceModeling." + std::to_string(l) + ".rnn.";
        const HostTensor* wf = find(h, M, p + "weight_ih_l0");
        const HostTensor* wb = find(h, M, p + "weight_ih_l0_reverse");
        const HostTensor* hf = find(h, M, p + "weight_hh_l0");
        const HostTensor* hb = find(h, M, p + "weight_hh_l0_reverse");
        const HostTensor* bif = find(h, M, p + "bias_ih_l0");
        const HostTensor* bhf = find(h, M, p + "bias_hh_l0");
        const HostTensor* bib = find(h, M, p + "bias_ih_l0_reverse");
        const HostTensor* bhb = find(h, M, p + "bias_hh_l0_reverse");
        if (!wf || !wb || !hf || !hb || !bif || !bhf || !bib || !bhb) return h->fail(LOCR_ERR_STATE, "missing LSTM tensors");
        const int nin = (int)wf->shape[1];
        // Row order of the recurrence kernel (lstm_tc.cu): n' = unit*4 + gate; PyTorch's is gate*256 + unit (i, f, g, o).
        auto perm = [](int np) { return (np & 3) * 256 + (np >> 2); };
        HostTensor stacked, sbias;
        stacked.shape = {2048, nin};
        stacked.data.resize((size_t)2048 * nin);
        sbias.shape = {2048};
        sbias.data.resize(2048);
        std::vector<uint16_t> whh((size_t)2048 * 256);
        for (int d = 0; d < 2; ++d) {
            const HostTensor* wi = d == 0 ? wf : wb;
            const HostTensor* hh = d == 0 ? hf : hb;
            const HostTensor* bi = d == 0 ? bif : bib;
            const HostTensor* bh = d == 0 ? bhf : bhb;
            for (int np = 0; np < 1024; ++np) {
                const int r = perm(np);
                memcpy(&stacked.data[(size_t)(d * 1024 + np) * nin], &wi->data[(size_t)r * nin], sizeof(float) * nin);
                sbias.data[d * 1024 + np] = bi->data[r] + bh->data[r];
                for (int k = 0; k < 256; ++k)
                    whh[(size_t)(d * 1024 + np) * 256 + k] = f32_to_act(hh->data[(size_t)r * 256 + k], h->cfg.act_dtype);
            }
        }
        const std::string name = "lstm" + std::to_string(l) + ".xproj";
        h->host[M][name + ".weight"] = stacked;
        h->host[M][name + ".bias"] = sbias;
        // exact: layer 0 reads the split-precision visual features, layer 1 the split-precision output of linear 0
        if ((rc = F(name, ""))) return rc;
        h->lstm_whh[l] = dev_upload(h, whh);
        // exact: the recurrence emits its hidden states as hi + lo pairs too (lstm_tc.cu SPLIT)
        if ((rc = F("SequenceModeling." + std::to_string(l) + ".linear", ""))) return rc;
    }
    if (h->cfg.head == LOCR_HEAD_CTC) {
        if ((rc = F("Prediction", ""))) return rc;
        if (h->conv["Prediction"].cout != h->cfg.num_classes) return h->fail(LOCR_ERR_STATE, "num_classes mismatch");
    } else {
        const std::string p = "Prediction.attention_cell.";
        const int C = h->cfg.num_classes;
        if ((rc = F(p + "i2h", ""))) return rc;
        const HostTensor* h2h = find(h, M, p + "h2h.weight");
        const HostTensor* h2hb = find(h, M, p + "h2h.bias");
        const HostTensor* sw = find(h, M, p + "score.weight");
        const HostTensor* wih = find(h, M, p + "rnn.weight_ih");
        const HostTensor* whh = find(h, M, p + "rnn.weight_hh");
        const HostTensor* bih = find(h, M, p + "rnn.bias_ih");
        const HostTensor* bhh = find(h, M, p + "rnn.bias_hh");
        const HostTensor* gw = find(h, M, "Prediction.generator.weight");
        const HostTensor* gb = find(h, M, "Prediction.generator.bias");
        if (!h2h || !h2hb || !sw || !wih || !whh || !bih || !bhh || !gw || !gb || wih->shape[1] != 256 + C ||
            gw->shape[0] != C)
            return h->fail(LOCR_ERR_STATE, "missing or malformed attention tensors");
        // the decoder's big matrices are stored in 16 bits like every other weight of the path (the kernel is bound by
        // streaming them from L2 once per step); biases, the score vector, the one-hot rows and the generator stay fp32
        {
            std::vector<uint16_t> h2h16((size_t)256 * 256);
            for (int k = 0; k < 256; ++k)
                for (int j = 0; j < 256; ++j)
                    h2h16[(size_t)k * 256 + j] = f32_to_act(h2h->data[(size_t)j * 256 + k], h->cfg.act_dtype);
            h->u16["att.h2h_wt"] = dev_upload(h, h2h16);
        }
        upload_f32(h, "att.h2h_b", h2hb->data);
        upload_f32(h, "att.score", sw->data);
        const int K = 256 + C;
        std::vector<uint16_t> wg((size_t)256 * 256 * 8);
        std::vector<float> woh((size_t)C * 256 * 4), gbias(1024);
        for (int k = 0; k < 256; ++k)
            for (int j = 0; j < 256; ++j)
                for (int q = 0; q < 4; ++q) {
                    wg[((size_t)k * 256 + j) * 8 + q] = f32_to_act(wih->data[(size_t)(q * 256 + j) * K + k], h->cfg.act_dtype);
                    wg[((size_t)k * 256 + j) * 8 + 4 + q] = f32_to_act(whh->data[(size_t)(q * 256 + j) * 256 + k], h->cfg.act_dtype);
                }
        for (int v = 0; v < C; ++v)
            for (int j = 0; j < 256; ++j)
                for (int q = 0; q < 4; ++q) woh[((size_t)v * 256 + j) * 4 + q] = wih->data[(size_t)(q * 256 + j) * K + 256 + v];
        for (int i = 0; i < 1024; ++i) gbias[i] = bih->data[i] + bhh->data[i];
        h->u16["att.wg"] = dev_upload(h, wg);
        if (X) {   // exact arithmetic: the decoder's matrices stay fp32 (same layouts)
            std::vector<float> h2h32((size_t)256 * 256), wg32((size_t)256 * 256 * 8);
            for (int k = 0; k < 256; ++k)
                for (int j = 0; j < 256; ++j) {
                    h2h32[(size_t)k * 256 + j] = h2h->data[(size_t)j * 256 + k];
                    for (int q = 0; q < 4; ++q) {
                        wg32[((size_t)k * 256 + j) * 8 + q] = wih->data[(size_t)(q * 256 + j) * K + k];
                        wg32[((size_t)k * 256 + j) * 8 + 4 + q] = whh->data[(size_t)(q * 256 + j) * 256 + k];
                    }
                }
            upload_f32(h, "att.h2h_wt32", h2h32);
            upload_f32(h, "att.wg32", wg32);
        }
        upload_f32(h, "att.woh", woh);
        if (!h->u16["att.h2h_wt"] || !h->u16["att.wg"]) return h->fail(LOCR_ERR_CUDA, "weight upload failed");
        upload_f32(h, "att.gate_b", gbias);
        upload_f32(h, "att.gen_w", gw->data);
        upload_f32(h, "att.gen_b", gb->data);
    }
    for (auto& kv : h->f32)
        if (kv.second == nullptr) return h->fail(LOCR_ERR_CUDA, "weight upload failed: " + kv.first);
    h->host[M].clear();
    h->ready[M] = true;
    return LOCR_OK;
}

// CRNNet.forward (reference ocr/model.py:103-118): TPS_STN -> ResNet -> 2x BiLSTM -> CTC / attention head.
// LOCR_PREC_EXACT (X = 1): every tensor from conv0_1 to the prediction head is a split-precision pair [hi | lo] with
// twice the channel pitch; the convolutions read [hi | lo | hi] against [w_hi | w_hi | w_lo] (see fold_conv).
int engine_crnn_forward(locr_handle* h, const float* d_x, int B, float** logits) {
    if (!h->ready[LOCR_MODEL_CRNN]) return h->fail(LOCR_ERR_STATE, "CRNN weights not finalized");
    if (B <= 0) return h->fail(LOCR_ERR_INVALID, "empty crop batch");
    if (B > 65536) return h->fail(LOCR_ERR_CAPACITY, "more than 65536 crops in one recognition batch");  // 32-bit indices
    Ctx c{h};
    const int f16 = h->is_f16();
    const int X = h->exact();
    auto P = [X](int ch) { return (long)(X ? 2 * ch : ch); };     // channel pitch of a tensor with `ch` logical channels
    cudaStream_t s = h->stream;
    const std::string loc = kLoc, fe = kFe;
    const size_t big = (size_t)B * 32 * 100 * 64 * 2 * (X ? 2 : 1);
    void* sA = c.buf("crnn.sA", big * 2);   // the split-precision localisation tensors have 2x the channels
    void* sB = c.buf("crnn.sB", big);
    void* sC = c.buf("crnn.sC", big / 2);
    void* sD = c.buf("crnn.sD", big / 2);
    float* fid = (float*)c.buf("crnn.fid", (size_t)B * 40 * 4);
    float* xr = (float*)c.buf("crnn.rectified", (size_t)B * 3200 * 4);
    float* grid = (float*)c.buf("crnn.grid", (size_t)B * 3200 * 2 * 4);
    void* vis = c.buf("crnn.visual", (size_t)B * 26 * P(512) * 2);
    float* xproj = (float*)c.buf("crnn.xproj", (size_t)B * 26 * 2048 * 4);
    void* hcat = c.buf("crnn.hcat", (size_t)B * 26 * P(512) * 2);
    void* s0 = c.buf("crnn.s0", (size_t)B * 26 * P(256) * 2);
    void* s1 = c.buf("crnn.contextual", (size_t)B * 26 * P(256) * 2);
    const int C = h->cfg.num_classes;
    float* lg = (float*)c.buf("crnn.logits", (size_t)B * 26 * C * 4);
    if (c.rc != LOCR_OK) return c.rc;

    // ---- TPS localisation network (TPS_STN.py:38-58), split precision: tensors hold [hi | lo] channel halves
    const ConvW& l0 = h->conv[loc + "conv.0"];
    // conv.0 + BN + ReLU + MaxPool2d(2, 2) in one pass (split-precision output [hi | lo])
    { ProfScope ps_(h, "direct_conv.loc0+pool", 0, false); launch_direct_conv3x3(d_x, 0, B, 32, 100, 32, 100, 0, 0, l0.w32, l0.bias, 1, 64, sB, 128, 1, f16, s, 1, 1); }
    c.pool(sC, 256, 1);
    c.tc(loc + "conv.4", sB, B, 16, 50, 128, nullptr, 256, 1, 1, 1, 1, 1, 0, nullptr, 0, 1);
    c.pool(sB, 512, 1);
    c.tc(loc + "conv.8", sC, B, 8, 25, 256, nullptr, 512, 1, 1, 1, 1, 1, 0, nullptr, 0, 1);
    c.tc(loc + "conv.12", sB, B, 4, 12, 512, sA, 1024, 1, 1, 1, 1, 1, 0, nullptr, 0, 1);
    { ProfScope ps_(h, "loc_head", 0, false); launch_loc_head(sA, B, 48, h->f32["loc.w1t"], h->f32["loc.b1"], h->f32["loc.w2t"], h->f32["loc.b2"], fid, f16, s, 1); }
    { ProfScope ps_(h, "tps_sample", 0, false); launch_tps_sample(fid, h->f32["tps.inv"], h->f32["tps.phat_t"], d_x, xr, grid, B, s); }
    h->launches += 3;

    // ---- ResNet feature extractor (resnet50v1.py:101-135)
    const ConvW& r0 = h->conv[fe + "conv0_1"];
    if (X) {
        void* r0out = c.buf("crnn.res0", (size_t)B * 32 * 100 * 64 * 2);
        if (c.rc != LOCR_OK) return c.rc;
        { ProfScope ps_(h, "direct_conv.res0", 0, false); launch_direct_conv3x3(xr, 0, B, 32, 100, 32, 100, 0, 0, r0.w32, r0.bias, 1, 32, r0out, 64, 1, f16, s, 1, 0, 0); }
        c.pool(sB, 128, 1);
        c.tc(fe + "conv0_2", r0out, B, 32, 100, 64, nullptr, 128, 1, 1, 1, 1, 1, 0, nullptr, 0, 1);
    } else {
        // conv0_1 writes a row-padded tensor (103 pixels per row, pads zero) so that conv0_2 takes the 4-pixel window view
        // (6 k-blocks of 128-byte rows instead of 9 of 64 bytes; conv_tc.cuh x_row_px)
        void* r0out = c.zbuf_grow("crnn.res0", (size_t)B * 32 * 103 * 32 * 2);
        if (c.rc != LOCR_OK || r0out == nullptr) return c.rc != LOCR_OK ? c.rc : h->fail(LOCR_ERR_CUDA, "allocation failed");
        { ProfScope ps_(h, "direct_conv.res0", 0, false); launch_direct_conv3x3(xr, 0, B, 32, 100, 32, 100, 0, 0, r0.w32, r0.bias, 1, 32, r0out, 32, 1, f16, s, 0, 0, 103); }
        c.pool(sB, 64, 1);
        c.rows(103, 0);
        c.tc(fe + "conv0_2", r0out, B, 32, 100, 32, nullptr, 64, 1, 1, 0);
    }
    h->launches += 1;
    // x lives in `cur`; BasicBlock (resnet50v1.py:33-48): relu(bn2(conv2(relu(bn1(conv1 x)))) + residual)
    void* cur = sB;
    void* other = sA;
    int Hc = 16, Wc = 50, Cc = 64;
    const int nblocks[5] = {0, 1, 2, 5, 3};
    const int planes[5] = {0, 128, 256, 512, 512};
    for (int l = 1; l <= 4; ++l) {
        for (int i = 0; i < nblocks[l]; ++i) {
            const std::string p = fe + "layer" + std::to_string(l) + "." + std::to_string(i) + ".";
            const int Pl = planes[l];
            c.tc(p + "conv1", cur, B, Hc, Wc, P(Cc), sC, P(Pl), 1, 1, 1, 1, 1, 0, nullptr, 0, X);
            const void* res = cur;
            long res_pitch = P(Cc);
            int res_ch = Cc;
            if (h->conv.count(p + "downsample.0")) {
                c.tc(p + "downsample.0", cur, B, Hc, Wc, P(Cc), sD, P(Pl), 0, 0, 0, 1, 1, 0, nullptr, 0, X);
                res = sD;
                res_pitch = P(Pl);
                res_ch = Pl;
            }
            if (X) c.res_split(res_ch);
            c.tc(p + "conv2", sC, B, Hc, Wc, P(Pl), other, P(Pl), 1, 1, 1, 1, 1, 0, res, res_pitch, X);
            std::swap(cur, other);
            Cc = Pl;
        }
        const std::string tail = fe + "conv" + std::to_string(l);
        if (l == 1) {
            // conv1 + BN + ReLU + MaxPool2d(2, 2) (resnet50v1.py:110-112): pooled in the conv epilogue, only the pooled tensor is written
            c.pool(other, P(Cc), 1);
            c.tc(tail, cur, B, Hc, Wc, P(Cc), nullptr, P(Cc), 1, 1, 1, 1, 1, 0, nullptr, 0, X);
            std::swap(cur, other);
            Hc /= 2; Wc /= 2;
        } else if (l <= 3) {
            c.tc(tail, cur, B, Hc, Wc, P(Cc), other, P(Cc), 1, 1, 1, 1, 1, 0, nullptr, 0, X);
            std::swap(cur, other);
        }
        if (l == 2) {
            { ProfScope ps_(h, "maxpool.res2", 0, false); launch_maxpool(cur, P(Cc), B, Hc, Wc, Cc, other, P(Cc), 2, 2, 2, 1, 0, 1, f16, s, X); }
            Hc = (Hc - 2) / 2 + 1; Wc = Wc + 2 - 2 + 1;
            h->launches++;
            std::swap(cur, other);
        }
    }
    dbg(h, "layer4", cur, 0, {B, Hc, Wc, 512}, P(512));
    if (X) h->dbg["layer4"].lo_off = 512;
    c.tc(fe + "conv4_1", cur, B, Hc, Wc, P(512), other, P(512), 1, 0, 1, 1, 2, 0, nullptr, 0, X);      // k2 s(2,1) p(0,1) -> 2x27
    c.tc(fe + "conv4_2", other, B, 2, 27, P(512), vis, P(512), 1, 0, 0, 1, 1, 0, nullptr, 0, X);        // k2 s1 p0     -> 1x26

    // ---- sequence modelling (model.py:107-112; AdaptiveAvgPool over H=1 is the identity) as [B*26, C] GEMMs
    const int R = B * 26;
    c.tc("lstm0.xproj", vis, 1, 1, R, P(512), xproj, 2048, 0, 0, 0, 1, 1, 1);
    if (c.rc == LOCR_OK) {
        ProfScope ps_(h, "lstm", 0, false);
        if (launch_lstm_tc(xproj, h->lstm_whh[0], hcat, B, 26, f16, s, X) != cudaSuccess)
            c.rc = h->fail(LOCR_ERR_CUDA, "BiLSTM launch failed");
    }
    c.tc("SequenceModeling.0.linear", hcat, 1, 1, R, P(512), s0, P(256), 0, 0, 0, 1, 1, 0, nullptr, 0, X);
    c.tc("lstm1.xproj", s0, 1, 1, R, P(256), xproj, 2048, 0, 0, 0, 1, 1, 1);
    if (c.rc == LOCR_OK) {
        ProfScope ps_(h, "lstm", 0, false);
        if (launch_lstm_tc(xproj, h->lstm_whh[1], hcat, B, 26, f16, s, X) != cudaSuccess)
            c.rc = h->fail(LOCR_ERR_CUDA, "BiLSTM launch failed");
    }
    c.tc("SequenceModeling.1.linear", hcat, 1, 1, R, P(512), s1, P(256), 0, 0, 0, 1, 1, 0, nullptr, 0, X);
    h->launches += 2;
    if (h->cfg.head == LOCR_HEAD_CTC) {
        c.tc("Prediction", s1, 1, 1, R, P(256), lg, C, 0, 0, 0, 1, 1, 1);
    } else {
        float* fproj = (float*)c.buf("crnn.fproj", (size_t)R * 256 * 4);
        c.tc("Prediction.attention_cell.i2h", s1, 1, 1, R, P(256), fproj, 256, 0, 0, 0, 1, 1, 1);
        if (c.rc == LOCR_OK) {
            AttnWeights w;
            w.h2h_wt = (const uint16_t*)h->u16["att.h2h_wt"]; w.h2h_b = h->f32["att.h2h_b"]; w.score_w = h->f32["att.score"];
            w.wg = (const uint16_t*)h->u16["att.wg"]; w.woh = h->f32["att.woh"]; w.gate_b = h->f32["att.gate_b"];
            w.gen_w = h->f32["att.gen_w"]; w.gen_b = h->f32["att.gen_b"];
            if (X) { w.h2h_wt32 = h->f32["att.h2h_wt32"]; w.wg32 = h->f32["att.wg32"]; }
            { ProfScope ps_(h, "attention", 0, false); launch_attention(s1, fproj, w, lg, B, C, f16, s, P(256), X ? 256 : 0); }
            h->launches++;
        }
    }
    if (c.rc != LOCR_OK) return c.rc;
    LOCR_CUDA_OK(cudaGetLastError());
    dbg(h, "fiducials", fid, 1, {B, 20, 2}, 2);
    dbg(h, "grid", grid, 1, {B, 3200, 2}, 2);
    dbg(h, "rectified", xr, 1, {B, 32, 100}, 100);
    dbg(h, "visual", vis, 0, {B, 26, 512}, P(512));
    dbg(h, "contextual", s1, 0, {B, 26, 256}, P(256));
    if (X) { h->dbg["visual"].lo_off = 512; h->dbg["contextual"].lo_off = 256; }
    dbg(h, "logits", lg, 1, {B, 26, C}, C);
    *logits = lg;
    return LOCR_OK;
}

}  // namespace locr
