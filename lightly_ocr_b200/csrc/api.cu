// C ABI of liblocr (include/locr.h): handle lifecycle, weight loading, and the debug read-back used by tests.
#include <algorithm>

#include "engine.cuh"

using namespace locr;

extern "C" {

LOCR_API int locr_create(const locr_config* cfg, locr_handle** out) {
    if (cfg == nullptr || out == nullptr) return fail(LOCR_ERR_INVALID, "locr_create: null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(LOCR_ERR_CUDA, "locr_create: no CUDA device (liblocr has no CPU fallback)");
    if (cfg->device_id < 0 || cfg->device_id >= ndev) return fail(LOCR_ERR_INVALID, "locr_create: bad device id");
    LOCR_CUDA_OK(cudaSetDevice(cfg->device_id));
    cudaDeviceProp prop;
    LOCR_CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device_id));
    if (prop.major != 10)
        return fail(LOCR_ERR_CUDA, std::string("locr_create: sm_100a kernels need a Blackwell B200, found ") + prop.name);
    locr_handle* h = new locr_handle();
    h->cfg = *cfg;
    if (h->cfg.canvas_size <= 0) h->cfg.canvas_size = 1280;
    if (h->cfg.mag_ratio <= 0.f) h->cfg.mag_ratio = 1.5f;
    if (h->cfg.text_threshold <= 0.f) h->cfg.text_threshold = 0.7f;
    if (h->cfg.link_threshold <= 0.f) h->cfg.link_threshold = 0.4f;
    if (h->cfg.low_text <= 0.f) h->cfg.low_text = 0.4f;
    if (h->cfg.num_classes <= 0) h->cfg.num_classes = h->cfg.head == LOCR_HEAD_CTC ? 37 : 38;
    if (h->cfg.crnn_precision != LOCR_PREC_EXACT) h->cfg.crnn_precision = LOCR_PREC_FAST;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return fail(LOCR_ERR_CUDA, "locr_create: stream creation failed");
    }
    *out = h;
    return LOCR_OK;
}

LOCR_API void locr_destroy(locr_handle* h) {
    if (h == nullptr) return;
    cudaSetDevice(h->cfg.device_id);
    cudaStreamSynchronize(h->stream);
    for (void* p : h->owned) cudaFree(p);
    for (auto& kv : h->buffers)
        if (kv.second.first) cudaFree(kv.second.first);
    for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
    if (h->timer0) cudaEventDestroy(h->timer0);
    if (h->timer1) cudaEventDestroy(h->timer1);
    cudaStreamDestroy(h->stream);
    delete h;
}

LOCR_API int locr_load_tensor(locr_handle* h, int model, const char* key, const float* data, const int64_t* shape,
                              int ndim) {
    if (h == nullptr || key == nullptr || data == nullptr || (model != 0 && model != 1) || ndim < 0 || ndim > 8)
        return fail(LOCR_ERR_INVALID, "locr_load_tensor: bad argument");
    std::string k(key);
    if (k.rfind("module.", 0) == 0) k = k.substr(7);  // copyStateDict (reference net.py:24-34)
    HostTensor t;
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        t.shape.push_back(shape[i]);
        n *= shape[i];
    }
    t.data.assign(data, data + n);
    h->host[model][k] = std::move(t);
    h->ready[model] = false;
    return LOCR_OK;
}

LOCR_API int locr_finalize(locr_handle* h, int model) {
    if (h == nullptr || (model != 0 && model != 1)) return fail(LOCR_ERR_INVALID, "locr_finalize: bad argument");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    return model == LOCR_MODEL_CRAFT ? engine_finalize_craft(h) : engine_finalize_crnn(h);
}

LOCR_API int64_t locr_launch_count(const locr_handle* h) { return h ? h->launches : 0; }

LOCR_API const char* locr_last_error(const locr_handle* h) {
    // Every failure path (with or without a handle) records its text for the calling thread; the handle's copy only
    // serves a caller that asks from another thread than the one that made the failing call.
    if (!tls_error().empty()) return tls_error().c_str();
    return h != nullptr ? h->err.c_str() : "";
}

LOCR_API int locr_timer_start(locr_handle* h) {
    if (h == nullptr) return fail(LOCR_ERR_INVALID, "null handle");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    if (h->timer0 == nullptr) {
        LOCR_CUDA_OK(cudaEventCreate(&h->timer0));
        LOCR_CUDA_OK(cudaEventCreate(&h->timer1));
    }
    LOCR_CUDA_OK(cudaStreamSynchronize(h->stream));
    LOCR_CUDA_OK(cudaEventRecord(h->timer0, h->stream));
    return LOCR_OK;
}

LOCR_API int locr_timer_stop(locr_handle* h, float* ms) {
    if (h == nullptr || ms == nullptr || h->timer0 == nullptr) return fail(LOCR_ERR_INVALID, "timer not started");
    LOCR_CUDA_OK(cudaEventRecord(h->timer1, h->stream));
    LOCR_CUDA_OK(cudaEventSynchronize(h->timer1));
    LOCR_CUDA_OK(cudaEventElapsedTime(ms, h->timer0, h->timer1));
    return LOCR_OK;
}

LOCR_API int locr_profile(locr_handle* h, int enable) {
    if (h == nullptr) return fail(LOCR_ERR_INVALID, "null handle");
    h->profile = enable != 0;
    return LOCR_OK;
}

LOCR_API int locr_profile_read(locr_handle* h, double* conv_ms, double* conv_flops, int64_t* conv_launches) {
    if (h == nullptr) return fail(LOCR_ERR_INVALID, "null handle");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    LOCR_CUDA_OK(cudaStreamSynchronize(h->stream));
    double ms = 0, fl = 0;
    int64_t n = 0;
    for (auto& r : h->prof) {
        float t = 0;
        cudaEventElapsedTime(&t, r.e0, r.e1);
        if (r.is_conv) {
            ms += t;
            fl += r.flops;
            ++n;
        }
        auto& a = h->prof_layers[r.name];
        a.ms += t;
        a.flops += r.flops;
        a.n += 1;
        h->event_pool.push_back(r.e0);      // reused by the next ProfScope: creating events per launch costs host time
        h->event_pool.push_back(r.e1);      // that shows up as gaps between the kernels of a profiled pass
    }
    if (conv_ms) *conv_ms = ms;
    if (conv_flops) *conv_flops = fl;
    if (conv_launches) *conv_launches = n;
    h->prof.clear();
    return LOCR_OK;
}

/* Per-layer totals accumulated by locr_profile_read since the last call, as text lines "name ms flops launches". */
LOCR_API int locr_profile_layers(locr_handle* h, char* out, int64_t capacity) {
    if (h == nullptr || out == nullptr || capacity <= 0) return fail(LOCR_ERR_INVALID, "bad argument");
    std::string s;
    char line[256];
    for (auto& kv : h->prof_layers) {
        snprintf(line, sizeof(line), "%s %.6f %.0f %lld\n", kv.first.c_str(), kv.second.ms, kv.second.flops,
                 (long long)kv.second.n);
        s += line;
    }
    h->prof_layers.clear();
    if ((int64_t)s.size() + 1 > capacity) return h->fail(LOCR_ERR_CAPACITY, "locr_profile_layers: buffer too small");
    memcpy(out, s.c_str(), s.size() + 1);
    return LOCR_OK;
}

/* Range audit of the 16-bit activations: locr_audit(h, 1), run any forward passes, locr_audit_read. */
LOCR_API int locr_audit(locr_handle* h, int enable) {
    if (h == nullptr) return fail(LOCR_ERR_INVALID, "null handle");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    if (enable) {
        if (h->audit_slots == nullptr) {
            void* p = nullptr;
            LOCR_CUDA_OK(cudaMalloc(&p, kAuditSlots * sizeof(float)));
            h->owned.push_back(p);
            h->audit_slots = (float*)p;
        }
        LOCR_CUDA_OK(cudaMemsetAsync(h->audit_slots, 0, kAuditSlots * sizeof(float), h->stream));
        h->audit_names.clear();
    }
    h->audit = enable != 0;
    return LOCR_OK;
}

/* Text lines "layer-name abs-max" for every audited launch since locr_audit(h, 1), in launch order (a layer that ran
 * several times appears several times).  65504 (fp16) means the layer saturated. */
LOCR_API int locr_audit_read(locr_handle* h, char* out, int64_t capacity) {
    if (h == nullptr || out == nullptr || capacity <= 0) return fail(LOCR_ERR_INVALID, "bad argument");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    std::vector<float> v(h->audit_names.size());
    if (!v.empty()) {
        LOCR_CUDA_OK(cudaStreamSynchronize(h->stream));
        LOCR_CUDA_OK(cudaMemcpy(v.data(), h->audit_slots, v.size() * sizeof(float), cudaMemcpyDeviceToHost));
    }
    std::string s;
    char line[256];
    for (size_t i = 0; i < v.size(); ++i) {
        snprintf(line, sizeof(line), "%s %.9g\n", h->audit_names[i].c_str(), v[i]);
        s += line;
    }
    if ((int64_t)s.size() + 1 > capacity) return h->fail(LOCR_ERR_CAPACITY, "locr_audit_read: buffer too small");
    memcpy(out, s.c_str(), s.size() + 1);
    return LOCR_OK;
}

/* ---- debug / test entry points ---- */

LOCR_API int locr_debug_craft_scores(locr_handle* h, const uint8_t* bgr, int B, int img_h, int img_w, float* score) {
    if (h == nullptr || bgr == nullptr || score == nullptr) return fail(LOCR_ERR_INVALID, "null argument");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    const int H = (img_h + 31) / 32 * 32, W = (img_w + 31) / 32 * 32;
    const size_t nbytes = (size_t)B * img_h * img_w * 3;
    h->resident.clear();   // the image buffer is about to be overwritten (and possibly moved)
    uint8_t* d = (uint8_t*)engine_buffer(h, "images", nbytes);
    if (!d) return h->fail(LOCR_ERR_CUDA, "allocation failed");
    LOCR_CUDA_OK(cudaMemcpyAsync(d, bgr, nbytes, cudaMemcpyHostToDevice, h->stream));
    float* sc = nullptr;
    int rc = engine_craft_forward(h, d, B, img_h, img_w, H, W, &sc);
    if (rc != LOCR_OK) return rc;
    LOCR_CUDA_OK(cudaMemcpyAsync(score, sc, (size_t)B * (H / 2) * (W / 2) * 2 * 4, cudaMemcpyDeviceToHost, h->stream));
    LOCR_CUDA_OK(cudaStreamSynchronize(h->stream));
    return LOCR_OK;
}

LOCR_API int locr_debug_crnn(locr_handle* h, const uint8_t* u8, int n, float* logits, int32_t* ids, char* text,
                             int text_stride, int32_t* has_eos, float* conf) {
    if (h == nullptr || u8 == nullptr || n <= 0) return fail(LOCR_ERR_INVALID, "bad argument");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    const size_t np = (size_t)n * 3200;
    std::vector<float> x(np);
    // ToTensor (/255) then sub_(0.5).div_(0.5)  (reference tools/dataset.py:45-46), fp32 in that order
    for (size_t i = 0; i < np; ++i) x[i] = ((float)u8[i] / 255.0f - 0.5f) / 0.5f;
    float* dx = (float*)engine_buffer(h, "crnn.x", np * 4);
    if (!dx) return h->fail(LOCR_ERR_CUDA, "allocation failed");
    LOCR_CUDA_OK(cudaMemcpyAsync(dx, x.data(), np * 4, cudaMemcpyHostToDevice, h->stream));
    float* lg = nullptr;
    int rc = engine_crnn_forward(h, dx, n, &lg);
    if (rc != LOCR_OK) return rc;
    const int C = h->cfg.num_classes;
    int32_t* d_ids = (int32_t*)engine_buffer(h, "dec.ids", (size_t)n * 26 * 4);
    char* d_text = (char*)engine_buffer(h, "dec.text", (size_t)n * text_stride);
    int32_t* d_eos = (int32_t*)engine_buffer(h, "dec.eos", (size_t)n * 4);
    float* d_conf = (float*)engine_buffer(h, "dec.conf", (size_t)n * 4);
    if (!d_ids || !d_text || !d_eos || !d_conf) return h->fail(LOCR_ERR_CUDA, "allocation failed");
    launch_decode(lg, n, C, h->cfg.head == LOCR_HEAD_ATTN, d_ids, d_text, text_stride, d_eos, d_conf, h->stream);
    h->launches++;
    LOCR_CUDA_OK(cudaGetLastError());
    h->last_logits = lg; h->last_ids = d_ids; h->last_n = n;
    if (logits) LOCR_CUDA_OK(cudaMemcpyAsync(logits, lg, (size_t)n * 26 * C * 4, cudaMemcpyDeviceToHost, h->stream));
    if (ids) LOCR_CUDA_OK(cudaMemcpyAsync(ids, d_ids, (size_t)n * 26 * 4, cudaMemcpyDeviceToHost, h->stream));
    if (text) LOCR_CUDA_OK(cudaMemcpyAsync(text, d_text, (size_t)n * text_stride, cudaMemcpyDeviceToHost, h->stream));
    if (has_eos) LOCR_CUDA_OK(cudaMemcpyAsync(has_eos, d_eos, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (conf) LOCR_CUDA_OK(cudaMemcpyAsync(conf, d_conf, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    LOCR_CUDA_OK(cudaStreamSynchronize(h->stream));
    return LOCR_OK;
}

/* Copies a named intermediate of the last forward pass to the host as fp32 (dense, pitch removed). */
LOCR_API int locr_debug_read(locr_handle* h, const char* name, float* out, int64_t capacity, int64_t* shape,
                             int* ndim) {
    if (h == nullptr || name == nullptr) return fail(LOCR_ERR_INVALID, "null argument");
    auto it = h->dbg.find(name);
    if (it == h->dbg.end()) return h->fail(LOCR_ERR_INVALID, std::string("unknown debug tensor ") + name);
    const DebugTensor& d = it->second;
    int64_t n = 1;
    for (size_t i = 0; i < d.shape.size(); ++i) {
        if (shape) shape[i] = d.shape[i];
        n *= d.shape[i];
    }
    if (ndim) *ndim = (int)d.shape.size();
    if (out == nullptr) return LOCR_OK;
    if (capacity < n) return h->fail(LOCR_ERR_CAPACITY, "debug read: capacity too small");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    LOCR_CUDA_OK(cudaStreamSynchronize(h->stream));
    const int64_t inner = d.shape.back();
    const int64_t rows = n / inner;
    const size_t esz = d.kind == 0 ? 2 : (d.kind == 3 ? 1 : 4);
    std::vector<uint8_t> raw((size_t)rows * inner * esz), raw_lo;
    LOCR_CUDA_OK(cudaMemcpy2D(raw.data(), inner * esz, d.p, d.pitch * esz, inner * esz, rows, cudaMemcpyDeviceToHost));
    if (d.kind == 0 && d.lo_off > 0) {
        raw_lo.resize(raw.size());
        LOCR_CUDA_OK(cudaMemcpy2D(raw_lo.data(), inner * esz, (const uint16_t*)d.p + d.lo_off, d.pitch * esz, inner * esz,
                                  rows, cudaMemcpyDeviceToHost));
    }
    for (int64_t i = 0; i < n; ++i) {
        if (d.kind == 0 && d.lo_off > 0)
            out[i] = act_to_f32(reinterpret_cast<uint16_t*>(raw.data())[i], h->cfg.act_dtype) +
                     act_to_f32(reinterpret_cast<uint16_t*>(raw_lo.data())[i], h->cfg.act_dtype);
        else if (d.kind == 0) out[i] = act_to_f32(reinterpret_cast<uint16_t*>(raw.data())[i], h->cfg.act_dtype);
        else if (d.kind == 1) out[i] = reinterpret_cast<float*>(raw.data())[i];
        else if (d.kind == 2) out[i] = (float)reinterpret_cast<int32_t*>(raw.data())[i];
        else out[i] = (float)raw[i];
    }
    return LOCR_OK;
}

}  // extern "C"
