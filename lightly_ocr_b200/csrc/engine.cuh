// The per-GPU engine behind locr_handle: weight store, BN folding / layout conversion, and the CRAFT and CRNN
// forward passes expressed as sequences of kernels on one CUDA stream.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "conv_tc.cuh"
#include "nn_kernels.cuh"
#include "util.cuh"

namespace locr {

struct HostTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
    int64_t numel() const {
        int64_t n = 1;
        for (auto s : shape) n *= s;
        return n;
    }
};

// One folded convolution / linear layer resident on the device.
struct ConvW {
    void* w = nullptr;       // 16-bit [cout_pad][kh][kw][cin] (tensor-core path)
    float* w32 = nullptr;    // fp32 [kh*kw*cin][cout]         (direct path, tiny cin)
    float* bias = nullptr;   // fp32 [cout_pad]
    int cin = 0, cout = 0, cout_pad = 0, kh = 1, kw = 1;
    int cin_real = 0;        // input channels of the reference layer (cin may be zero-padded to 16)
    int cin_wrap = 0;        // split-precision input: K channels [hi|lo|hi] = 3C read from a 2C-channel tensor
    int window = 0;          // 1: weights packed for the 4-pixel window view of a 3x3 conv (conv_tc.cuh: x_row_px)
};

struct DebugTensor {
    const void* p;
    int kind;  // 0 = 16-bit activation, 1 = fp32, 2 = int32, 3 = uint8
    std::vector<int64_t> shape;
    long pitch;  // elements per innermost row (>= shape.back())
    long lo_off = 0;  // > 0: split-precision tensor, the lo halves live lo_off elements further (value = hi + lo)
};

}  // namespace locr

struct ResidentImage {
    const uint8_t* p;  // device, packed rows of w*3 bytes
    int h, w;
};

struct locr_handle {
    locr_config cfg;
    std::vector<ResidentImage> resident;  // images of the last locr_detect call (for locr_recognize_boxes)
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    bool ready[2] = {false, false};
    std::map<std::string, locr::HostTensor> host[2];
    std::map<std::string, locr::ConvW> conv;
    std::map<std::string, float*> f32;   // misc fp32 device arrays (TPS buffers, FC weights, attention weights)
    void* lstm_whh[2] = {nullptr, nullptr};
    std::map<std::string, void*> u16;    // misc 16-bit device arrays (attention decoder weights)
    std::map<std::string, std::pair<void*, size_t>> buffers;  // named activation buffers, grown on demand
    std::map<std::string, std::pair<void*, size_t>> zeroed;   // row-padded buffers: (pointer, bytes) last zero-filled
    std::map<std::string, locr::DebugTensor> dbg;
    std::vector<void*> owned;  // weight allocations
    // per-launch profiling of the tensor-core conv kernel (bench.py roofline): event pairs + algorithmic FLOPs
    bool profile = false;
    struct ProfRec { cudaEvent_t e0, e1; double flops; std::string name; bool is_conv; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> event_pool;   // events of read-out records, reused
    struct ProfAgg { double ms = 0, flops = 0; int64_t n = 0; };
    std::map<std::string, ProfAgg> prof_layers;  // per-layer totals accumulated by locr_profile_read
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;
    // range audit (locr_audit): while enabled every 16-bit conv output is followed by an abs-max reduction
    bool audit = false;
    float* audit_slots = nullptr;            // device, kAuditSlots floats
    std::vector<std::string> audit_names;    // slot i <- layer name of the i-th audited launch since locr_audit(1)
    // logits / greedy ids of the last recognition forward (device), for locr_evaluate
    const float* last_logits = nullptr;
    const int32_t* last_ids = nullptr;
    int last_n = 0;

    int fail(int code, const std::string& m) {
        err = m;
        locr::tls_error() = m;
        return code;
    }
    int is_f16() const { return cfg.act_dtype == LOCR_ACT_F16 ? 1 : 0; }
    int exact() const { return cfg.crnn_precision == LOCR_PREC_EXACT ? 1 : 0; }
};

namespace locr {

constexpr int kAuditSlots = 512;

// Brackets one kernel launch with CUDA events on the handle's stream while profiling is enabled.
struct ProfScope {
    locr_handle* h;
    locr_handle::ProfRec r;
    bool on;
    ProfScope(locr_handle* h_, const std::string& name, double flops, bool is_conv) : h(h_), on(h_->profile) {
        if (!on) return;
        r.name = name; r.flops = flops; r.is_conv = is_conv;
        auto take = [&](cudaEvent_t* e) {
            if (!h->event_pool.empty()) { *e = h->event_pool.back(); h->event_pool.pop_back(); }
            else cudaEventCreate(e);
        };
        take(&r.e0);
        take(&r.e1);
        cudaEventRecord(r.e0, h->stream);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(r.e1, h->stream);
        h->prof.push_back(r);
    }
};

int engine_finalize_craft(locr_handle* h);
int engine_finalize_crnn(locr_handle* h);

// images: device uint8 [B][img_h][img_w][3] (packed rows).  Canvas H x W (multiples of 32, >= image).
// On success *score points at fp32 [B][H/2][W/2][2].
int engine_craft_forward(locr_handle* h, const uint8_t* d_images, int B, int img_h, int img_w, int H, int W,
                         float** score);

// x: device fp32 [B][32][100] normalised crops.  On success *logits points at fp32 [B][26][num_classes].
int engine_crnn_forward(locr_handle* h, const float* d_x, int B, float** logits);

// Named activation buffer (device), at least `bytes` large; contents are undefined after growth.
void* engine_buffer(locr_handle* h, const std::string& name, size_t bytes);
// Frees a named buffer (it is re-created on next use).
void engine_release(locr_handle* h, const std::string& name);

}  // namespace locr
