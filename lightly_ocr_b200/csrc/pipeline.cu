// The product entry points of the C ABI: locr_detect, locr_recognize, locr_recognize_boxes.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "engine.cuh"
#include "imgops.cuh"
#include "jpeg.cuh"
#include "png.cuh"
#include "polys.cuh"
#include "postproc.cuh"

using namespace locr;

namespace {

constexpr int kTextStride = 128;
constexpr int kMaxChunkCap = 32;
// images per CRAFT forward (activations: ~0.6 GB per 1280x960 image); LOCR_CRAFT_CHUNK overrides (experiments)
int max_chunk() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("LOCR_CRAFT_CHUNK");
        v = e ? atoi(e) : 8;
        if (v < 1) v = 1;
        if (v > kMaxChunkCap) v = kMaxChunkCap;
    }
    return v;
}

struct Pinned {
    void* p = nullptr;
    size_t cap = 0;
    ~Pinned() { if (p) cudaFreeHost(p); }
    void* get(size_t n) {
        if (n <= cap) return p;
        if (p) cudaFreeHost(p);
        cap = n + n / 4 + 4096;
        if (cudaMallocHost(&p, cap) != cudaSuccess) { p = nullptr; cap = 0; }
        return p;
    }
};

// Per-thread pinned staging (one host thread drives one handle).
Pinned& staging(int which) {
    static thread_local Pinned s[4];
    return s[which];
}

// resizeAspectRatio's size arithmetic (reference ocr/tools/imgproc.py:38-57), Python float semantics.
void craft_geometry(const locr_config& cfg, int h, int w, int* th, int* tw, int* H32, int* W32, double* ratio) {
    const int mx = h > w ? h : w;
    double target = (double)cfg.mag_ratio * mx;
    if (target > cfg.canvas_size) target = cfg.canvas_size;
    *ratio = target / mx;
    *th = (int)(h * *ratio);
    *tw = (int)(w * *ratio);
    *H32 = *th % 32 ? *th + (32 - *th % 32) : *th;
    *W32 = *tw % 32 ? *tw + (32 - *tw % 32) : *tw;
}

int run_crnn_and_decode(locr_handle* h, const float* d_x, int n, float* logits, int32_t* ids, char* text,
                        int32_t* has_eos, float* conf) {
    float* lg = nullptr;
    int rc = engine_crnn_forward(h, d_x, n, &lg);
    if (rc != LOCR_OK) return rc;
    const int C = h->cfg.num_classes;
    int32_t* d_ids = (int32_t*)engine_buffer(h, "dec.ids", (size_t)n * 26 * 4);
    char* d_text = (char*)engine_buffer(h, "dec.text", (size_t)n * kTextStride);
    int32_t* d_eos = (int32_t*)engine_buffer(h, "dec.eos", (size_t)n * 4);
    float* d_conf = (float*)engine_buffer(h, "dec.conf", (size_t)n * 4);
    if (!d_ids || !d_text || !d_eos || !d_conf) return h->fail(LOCR_ERR_CUDA, "allocation failed");
    { ProfScope ps_(h, "decode", 0, false);
    launch_decode(lg, n, C, h->cfg.head == LOCR_HEAD_ATTN, d_ids, d_text, kTextStride, d_eos, d_conf, h->stream); }
    h->launches++;
    LOCR_CUDA_OK(cudaGetLastError());
    h->last_logits = lg; h->last_ids = d_ids; h->last_n = n;
    cudaStream_t s = h->stream;
    if (logits) LOCR_CUDA_OK(cudaMemcpyAsync(logits, lg, (size_t)n * 26 * C * 4, cudaMemcpyDeviceToHost, s));
    if (ids) LOCR_CUDA_OK(cudaMemcpyAsync(ids, d_ids, (size_t)n * 26 * 4, cudaMemcpyDeviceToHost, s));
    if (text) LOCR_CUDA_OK(cudaMemcpyAsync(text, d_text, (size_t)n * kTextStride, cudaMemcpyDeviceToHost, s));
    if (has_eos) LOCR_CUDA_OK(cudaMemcpyAsync(has_eos, d_eos, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (conf) LOCR_CUDA_OK(cudaMemcpyAsync(conf, d_conf, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    LOCR_CUDA_OK(cudaStreamSynchronize(s));
    return LOCR_OK;
}

// Uploads crop descriptors, runs crop -> gray -> bicubic -> normalise, then the recogniser.
int run_crops(locr_handle* h, std::vector<CropDesc>& descs, float* logits, int32_t* ids, char* text,
              int32_t* has_eos, float* conf, uint8_t* resized_u8) {
    const int n = (int)descs.size();
    long coef_total = 0, inter_total = 0;
    for (auto& d : descs) {
        long ci, ib;
        crop_scratch_sizes(d.h, d.w, &d.ksh, &d.ksv, &ci, &ib);
        d.coef_off = coef_total;
        d.inter_off = inter_total;
        coef_total += ci;
        inter_total += ib;
    }
    CropDesc* d_desc = (CropDesc*)engine_buffer(h, "crop.desc", (size_t)n * sizeof(CropDesc));
    int32_t* d_coef = (int32_t*)engine_buffer(h, "crop.coef", (size_t)coef_total * 4);
    uint8_t* d_inter = (uint8_t*)engine_buffer(h, "crop.inter", (size_t)inter_total + 16);
    float* d_x = (float*)engine_buffer(h, "crnn.x", (size_t)n * 3200 * 4);
    uint8_t* d_u8 = (uint8_t*)engine_buffer(h, "crop.u8", (size_t)n * 3200);
    if (!d_desc || !d_coef || !d_inter || !d_x || !d_u8) return h->fail(LOCR_ERR_CUDA, "allocation failed");
    void* hd = staging(2).get((size_t)n * sizeof(CropDesc));
    if (!hd) return h->fail(LOCR_ERR_CUDA, "pinned allocation failed");
    memcpy(hd, descs.data(), (size_t)n * sizeof(CropDesc));
    LOCR_CUDA_OK(cudaMemcpyAsync(d_desc, hd, (size_t)n * sizeof(CropDesc), cudaMemcpyHostToDevice, h->stream));
    { ProfScope ps_(h, "crop_resize", 0, false); launch_crop_resize(d_desc, n, d_coef, d_inter, d_x, d_u8, h->stream); }
    h->launches++;
    {
        DebugTensor dt;
        dt.p = d_u8; dt.kind = 3; dt.shape = {n, 32, 100}; dt.pitch = 100;
        h->dbg["crop_u8"] = dt;
    }
    LOCR_CUDA_OK(cudaGetLastError());
    if (resized_u8)
        LOCR_CUDA_OK(cudaMemcpyAsync(resized_u8, d_u8, (size_t)n * 3200, cudaMemcpyDeviceToHost, h->stream));
    return run_crnn_and_decode(h, d_x, n, logits, ids, text, has_eos, conf);
}

}  // namespace

extern "C" {

static int detect_resident(locr_handle* h, int max_boxes_total, int32_t* rects, float* boxes,
                           int32_t* box_counts, float* score_maps) {
    cudaStream_t s = h->stream;
    const int n = (int)h->resident.size();
    std::vector<int> heights(n), widths(n);
    std::vector<size_t> img_off(n);
    const uint8_t* d_img = n ? h->resident[0].p : nullptr;
    for (int i = 0; i < n; ++i) {
        heights[i] = h->resident[i].h;
        widths[i] = h->resident[i].w;
        img_off[i] = (size_t)(h->resident[i].p - d_img);
    }
    int out_base = 0;
    size_t score_base = 0;
    int i0 = 0;
    while (i0 < n) {
        int i1 = i0 + 1;
        while (i1 < n && i1 - i0 < max_chunk() && heights[i1] == heights[i0] && widths[i1] == widths[i0] &&
               img_off[i1] - img_off[i1 - 1] == (size_t)heights[i0] * widths[i0] * 3)
            ++i1;
        const int B = i1 - i0, ih = heights[i0], iw = widths[i0];
        int th, tw, H32, W32;
        double ratio;
        craft_geometry(h->cfg, ih, iw, &th, &tw, &H32, &W32, &ratio);
        if (th <= 0 || tw <= 0) return h->fail(LOCR_ERR_INVALID, "locr_detect: image too small");
        const uint8_t* src = d_img + img_off[i0];
        if (th != ih || tw != iw) {
            uint8_t* d_rs = (uint8_t*)engine_buffer(h, "resized", (size_t)B * th * tw * 3);
            if (!d_rs) return h->fail(LOCR_ERR_CUDA, "allocation failed");
            launch_resize_linear_bgr(src, B, ih, iw, d_rs, th, tw, s);
            h->launches++;
            src = d_rs;
        }
        float* sc = nullptr;
        int rc = engine_craft_forward(h, src, B, th, tw, H32, W32, &sc);
        if (rc != LOCR_OK) return rc;
        const int mh = H32 / 2, mw = W32 / 2;
        const int cap = 4096;
        void* ws = engine_buffer(h, "pp.ws", postproc_workspace_bytes(B, mh, mw));
        float* d_boxes = (float*)engine_buffer(h, "pp.boxes", (size_t)B * cap * 8 * 4);
        int32_t* d_rects = (int32_t*)engine_buffer(h, "pp.rects", (size_t)B * cap * 4 * 4);
        int32_t* d_lab = (int32_t*)engine_buffer(h, "pp.lab", (size_t)B * cap * 4);
        int32_t* d_counts = (int32_t*)engine_buffer(h, "pp.counts", (size_t)B * 2 * 4);
        if (!ws || !d_boxes || !d_rects || !d_lab || !d_counts) return h->fail(LOCR_ERR_CUDA, "allocation failed");
        PostprocParams pp;
        pp.B = B; pp.H = mh; pp.W = mw;
        pp.low_text = h->cfg.low_text; pp.link_threshold = h->cfg.link_threshold;
        pp.text_threshold = h->cfg.text_threshold;
        const double inv = 1.0 / ratio;  // ratio_w = ratio_h = 1 / target_ratio (net.py:75)
        pp.scale_x = inv * 2;
        pp.scale_y = inv * 2;
        pp.max_boxes = cap;
        int nl;
        { ProfScope ps_(h, "postproc", 0, false); nl = launch_postproc(sc, pp, ws, d_boxes, d_rects, d_lab, d_counts, nullptr, s); }
        if (nl < 0) return h->fail(LOCR_ERR_INVALID, "locr_detect: score map larger than 1024 x 1024");
        h->launches += nl;
        LOCR_CUDA_OK(cudaGetLastError());
        int32_t counts[kMaxChunkCap * 2];
        LOCR_CUDA_OK(cudaMemcpyAsync(counts, d_counts, (size_t)B * 2 * 4, cudaMemcpyDeviceToHost, s));
        LOCR_CUDA_OK(cudaStreamSynchronize(s));
        for (int b = 0; b < B; ++b) {
            if (counts[b * 2 + 1] >= 65535) return h->fail(LOCR_ERR_CAPACITY, "locr_detect: more than 65534 components");
            const int k = counts[b * 2] < cap ? counts[b * 2] : cap;
            if (out_base + k > max_boxes_total) return h->fail(LOCR_ERR_CAPACITY, "locr_detect: max_boxes_total too small");
            LOCR_CUDA_OK(cudaMemcpyAsync(rects + (size_t)out_base * 4, d_rects + (size_t)b * cap * 4, (size_t)k * 16,
                                         cudaMemcpyDeviceToHost, s));
            if (boxes)
                LOCR_CUDA_OK(cudaMemcpyAsync(boxes + (size_t)out_base * 8, d_boxes + (size_t)b * cap * 8,
                                             (size_t)k * 32, cudaMemcpyDeviceToHost, s));
            box_counts[i0 + b] = k;
            out_base += k;
        }
        if (score_maps) {
            const size_t cnt = (size_t)B * mh * mw * 2;
            LOCR_CUDA_OK(cudaMemcpyAsync(score_maps + score_base, sc, cnt * 4, cudaMemcpyDeviceToHost, s));
            score_base += cnt;
        }
        LOCR_CUDA_OK(cudaStreamSynchronize(s));
        i0 = i1;
    }
    return LOCR_OK;
}

LOCR_API int locr_detect(locr_handle* h, const uint8_t* const* bgr, const int* heights, const int* widths,
                         const int* strides, int n, int max_boxes_total, int32_t* rects, float* boxes,
                         int32_t* box_counts, float* score_maps) {
    if (h == nullptr || bgr == nullptr || heights == nullptr || widths == nullptr || n <= 0 || rects == nullptr ||
        box_counts == nullptr)
        return fail(LOCR_ERR_INVALID, "locr_detect: bad argument");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    cudaStream_t s = h->stream;
    // all images stay resident (packed) for a following locr_recognize_boxes / locr_detect_resident
    size_t total_bytes = 0;
    std::vector<size_t> img_off(n);
    for (int i = 0; i < n; ++i) {
        if (heights[i] <= 0 || widths[i] <= 0 || bgr[i] == nullptr) return h->fail(LOCR_ERR_INVALID, "locr_detect: empty image");
        img_off[i] = total_bytes;
        total_bytes += (size_t)heights[i] * widths[i] * 3;
    }
    h->resident.clear();   // stale after this point: the buffer may move and is overwritten
    uint8_t* d_img = (uint8_t*)engine_buffer(h, "images", total_bytes);
    if (!d_img) return h->fail(LOCR_ERR_CUDA, "image buffer allocation failed");
    // Images that already live in pinned (page-locked) host memory with packed rows are copied straight from the
    // caller's buffer; everything else goes through the per-thread pinned staging buffer first.
    uint8_t* h_img = nullptr;
    for (int i = 0; i < n; ++i) {
        const size_t row = (size_t)widths[i] * 3;
        const size_t st = strides ? (size_t)strides[i] : row;
        cudaPointerAttributes attr;
        const bool pinned = st == row && cudaPointerGetAttributes(&attr, bgr[i]) == cudaSuccess &&
                            attr.type == cudaMemoryTypeHost;
        if (pinned) {
            LOCR_CUDA_OK(cudaMemcpyAsync(d_img + img_off[i], bgr[i], row * heights[i], cudaMemcpyHostToDevice, s));
            continue;
        }
        cudaGetLastError();   // cudaPointerGetAttributes on plain pageable memory may leave an error behind
        if (h_img == nullptr) {
            h_img = (uint8_t*)staging(0).get(total_bytes);
            if (!h_img) return h->fail(LOCR_ERR_CUDA, "pinned staging allocation failed");
        }
        if (st == row) memcpy(h_img + img_off[i], bgr[i], row * heights[i]);
        else for (int y = 0; y < heights[i]; ++y) memcpy(h_img + img_off[i] + y * row, bgr[i] + y * st, row);
        LOCR_CUDA_OK(cudaMemcpyAsync(d_img + img_off[i], h_img + img_off[i], row * heights[i], cudaMemcpyHostToDevice, s));
    }
    h->resident.clear();
    for (int i = 0; i < n; ++i) h->resident.push_back({d_img + img_off[i], heights[i], widths[i]});
    return detect_resident(h, max_boxes_total, rects, boxes, box_counts, score_maps);
}

/* cv2.imread / cv2.imdecode(IMREAD_COLOR) of one JPEG file (pipeline.py:68): header only. */
LOCR_API int locr_jpeg_info(const uint8_t* data, int64_t nbytes, int* height, int* width, int* components) {
    if (data == nullptr || nbytes <= 0 || height == nullptr || width == nullptr || components == nullptr)
        return fail(LOCR_ERR_INVALID, "locr_jpeg_info: bad argument");
    std::string err;
    const int rc = jpeg_probe(data, (size_t)nbytes, height, width, components, &err);
    return rc == LOCR_OK ? LOCR_OK : fail(rc, err);
}

/* The same for any format the GPU ingest covers (JPEG or PNG, told apart by the file signature). */
LOCR_API int locr_image_info(const uint8_t* data, int64_t nbytes, int* height, int* width, int* components,
                             int* format) {
    if (data == nullptr || nbytes <= 0 || height == nullptr || width == nullptr || components == nullptr)
        return fail(LOCR_ERR_INVALID, "locr_image_info: bad argument");
    std::string err;
    const bool png = png_is_png(data, (size_t)nbytes);
    const int rc = png ? png_probe(data, (size_t)nbytes, height, width, components, &err)
                       : jpeg_probe(data, (size_t)nbytes, height, width, components, &err);
    if (rc == LOCR_OK && format) *format = png ? LOCR_FORMAT_PNG : LOCR_FORMAT_JPEG;
    return rc == LOCR_OK ? LOCR_OK : fail(rc, err);
}

/* Decodes n encoded files (JPEG or PNG) on the GPU and leaves the BGR images resident (packed) like locr_detect does. */
static int decode_resident(locr_handle* h, const uint8_t* const* blob, const int64_t* nbytes, int n, int* heights,
                           int* widths) {
    std::vector<size_t> img_off(n);
    std::vector<int> hh(n), ww(n);
    std::vector<char> is_png(n);
    size_t total_bytes = 0;
    for (int i = 0; i < n; ++i) {
        int comps = 0;
        std::string err;
        if (blob[i] == nullptr || nbytes[i] <= 0) return h->fail(LOCR_ERR_INVALID, "imdecode: empty input");
        is_png[i] = png_is_png(blob[i], (size_t)nbytes[i]) ? 1 : 0;
        const int rc = is_png[i] ? png_probe(blob[i], (size_t)nbytes[i], &hh[i], &ww[i], &comps, &err)
                                 : jpeg_probe(blob[i], (size_t)nbytes[i], &hh[i], &ww[i], &comps, &err);
        if (rc != LOCR_OK) return h->fail(rc, err);
        img_off[i] = total_bytes;
        total_bytes += (size_t)hh[i] * ww[i] * 3;
    }
    h->resident.clear();   // stale after this point: the buffer may move and is overwritten
    uint8_t* d_img = (uint8_t*)engine_buffer(h, "images", total_bytes);
    if (!d_img) return h->fail(LOCR_ERR_CUDA, "image buffer allocation failed");
    for (int kind = 0; kind < 2; ++kind) {
        std::vector<const uint8_t*> bl;
        std::vector<int64_t> nb;
        std::vector<uint8_t*> outs;
        for (int i = 0; i < n; ++i)
            if (is_png[i] == kind) {
                bl.push_back(blob[i]);
                nb.push_back(nbytes[i]);
                outs.push_back(d_img + img_off[i]);
            }
        if (bl.empty()) continue;
        const int rc = kind ? png_decode_to_device(h, bl.data(), nb.data(), (int)bl.size(), outs.data())
                            : jpeg_decode_to_device(h, bl.data(), nb.data(), (int)bl.size(), outs.data());
        if (rc != LOCR_OK) return rc;
    }
    for (int i = 0; i < n; ++i) {
        h->resident.push_back({d_img + img_off[i], hh[i], ww[i]});
        if (heights) heights[i] = hh[i];
        if (widths) widths[i] = ww[i];
    }
    return LOCR_OK;
}

LOCR_API int locr_imdecode(locr_handle* h, const uint8_t* jpeg, int64_t nbytes, uint8_t* bgr, int64_t capacity,
                           int* height, int* width) {
    if (h == nullptr || jpeg == nullptr || nbytes <= 0 || bgr == nullptr || height == nullptr || width == nullptr)
        return fail(LOCR_ERR_INVALID, "locr_imdecode: bad argument");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    const int rc = decode_resident(h, &jpeg, &nbytes, 1, height, width);
    if (rc != LOCR_OK) return rc;
    const size_t bytes = (size_t)*height * *width * 3;
    if ((int64_t)bytes > capacity) return h->fail(LOCR_ERR_CAPACITY, "locr_imdecode: output buffer too small");
    LOCR_CUDA_OK(cudaMemcpyAsync(bgr, h->resident[0].p, bytes, cudaMemcpyDeviceToHost, h->stream));
    LOCR_CUDA_OK(cudaStreamSynchronize(h->stream));
    return LOCR_OK;
}

LOCR_API int locr_detect_encoded(locr_handle* h, const uint8_t* const* jpeg, const int64_t* nbytes, int n,
                                 int max_boxes_total, int32_t* rects, float* boxes, int32_t* box_counts,
                                 float* score_maps, int* heights, int* widths) {
    if (h == nullptr || jpeg == nullptr || nbytes == nullptr || n <= 0 || rects == nullptr || box_counts == nullptr)
        return fail(LOCR_ERR_INVALID, "locr_detect_encoded: bad argument");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    const int rc = decode_resident(h, jpeg, nbytes, n, heights, widths);
    if (rc != LOCR_OK) return rc;
    return detect_resident(h, max_boxes_total, rects, boxes, box_counts, score_maps);
}

/* Detection on the images already resident on the device from the previous locr_detect call (no host-to-device copy):
 * the HBM-resident variant bench.py times as `value`. */
LOCR_API int locr_detect_resident(locr_handle* h, int max_boxes_total, int32_t* rects, float* boxes,
                                  int32_t* box_counts, float* score_maps) {
    if (h == nullptr || rects == nullptr || box_counts == nullptr) return fail(LOCR_ERR_INVALID, "bad argument");
    if (h->resident.empty()) return h->fail(LOCR_ERR_STATE, "locr_detect_resident: no resident images");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    return detect_resident(h, max_boxes_total, rects, boxes, box_counts, score_maps);
}

LOCR_API int locr_recognize(locr_handle* h, const uint8_t* const* img, const int* heights, const int* widths,
                            const int* strides, const int* channels, int n, float* logits, int32_t* token_ids,
                            char* text, int32_t* has_eos, float* conf) {
    if (h == nullptr || img == nullptr || heights == nullptr || widths == nullptr || n <= 0)
        return fail(LOCR_ERR_INVALID, "locr_recognize: bad argument");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    size_t total = 0;
    std::vector<size_t> off(n);
    for (int i = 0; i < n; ++i) {
        const int ch = channels ? channels[i] : 1;
        if (ch != 1 && ch != 3) return h->fail(LOCR_ERR_INVALID, "locr_recognize: channels must be 1 or 3");
        if (heights[i] <= 0 || widths[i] <= 0)
            return h->fail(LOCR_ERR_INVALID, "locr_recognize: empty crop (the reference's cv2.cvtColor raises here)");
        off[i] = total;
        total += ((size_t)heights[i] * widths[i] * ch + 15) / 16 * 16;
    }
    uint8_t* d_buf = (uint8_t*)engine_buffer(h, "crop.src", total);
    uint8_t* h_buf = (uint8_t*)staging(1).get(total);
    if (!d_buf || !h_buf) return h->fail(LOCR_ERR_CUDA, "crop buffer allocation failed");
    std::vector<CropDesc> descs(n);
    for (int i = 0; i < n; ++i) {
        const int ch = channels ? channels[i] : 1;
        const size_t row = (size_t)widths[i] * ch;
        const size_t st = strides ? (size_t)strides[i] : row;
        for (int y = 0; y < heights[i]; ++y) memcpy(h_buf + off[i] + y * row, img[i] + y * st, row);
        CropDesc& d = descs[i];
        d.src = d_buf + off[i];
        d.stride = (long)row;
        d.h = heights[i];
        d.w = widths[i];
        d.channels = ch;
    }
    LOCR_CUDA_OK(cudaMemcpyAsync(d_buf, h_buf, total, cudaMemcpyHostToDevice, h->stream));
    return run_crops(h, descs, logits, token_ids, text, has_eos, conf, nullptr);
}

/* evaluation() of the reference's training script (ocr/train/crnn.py:142-240) for one validation batch: recognises the
 * crops like locr_recognize, then computes on the GPU, from the logits still in HBM, the loss of every crop and whether
 * its greedy prediction spells the label; *cost is the scalar the reference's loss_fn returns for the batch. */
LOCR_API int locr_evaluate(locr_handle* h, const uint8_t* const* img, const int* heights, const int* widths,
                           const int* strides, const int* channels, int n, const int32_t* targets,
                           const int32_t* target_len, int64_t targets_total, float* loss, int32_t* correct,
                           int32_t* token_ids, char* text, float* conf, float* cost) {
    if (h == nullptr || targets == nullptr || target_len == nullptr || n <= 0 || targets_total < 0)
        return fail(LOCR_ERR_INVALID, "locr_evaluate: bad argument");
    const int attn = h->cfg.head == LOCR_HEAD_ATTN;
    const int C = h->cfg.num_classes;
    std::vector<int32_t> off(n);
    int tw = 0;
    if (attn) {
        if (targets_total % n != 0 || targets_total / n < 2)
            return h->fail(LOCR_ERR_INVALID, "locr_evaluate: attention targets must be [n][batch_max_len + 2]");
        tw = (int)(targets_total / n);
    } else {
        int64_t tot = 0;
        for (int i = 0; i < n; ++i) {
            if (target_len[i] < 0) return h->fail(LOCR_ERR_INVALID, "locr_evaluate: negative target length");
            off[i] = (int32_t)tot;
            tot += target_len[i];
        }
        if (tot != targets_total) return h->fail(LOCR_ERR_INVALID, "locr_evaluate: target lengths do not add up");
    }
    for (int64_t i = 0; i < targets_total; ++i)
        if (targets[i] < 0 || targets[i] >= C || (!attn && targets[i] == 0))
            return h->fail(LOCR_ERR_INVALID, "locr_evaluate: target class out of range");
    std::vector<int32_t> eos(n);
    int rc = locr_recognize(h, img, heights, widths, strides, channels, n, nullptr, token_ids, text, eos.data(), conf);
    if (rc != LOCR_OK) return rc;
    if (h->last_n != n || h->last_logits == nullptr) return h->fail(LOCR_ERR_STATE, "locr_evaluate: no logits");
    cudaStream_t s = h->stream;
    const size_t tb = (size_t)(targets_total > 0 ? targets_total : 1) * 4;
    int32_t* d_tg = (int32_t*)engine_buffer(h, "eval.targets", tb);
    int32_t* d_meta = (int32_t*)engine_buffer(h, "eval.meta", (size_t)n * 4 * 4);   // off | len | count | correct
    float* d_loss = (float*)engine_buffer(h, "eval.loss", (size_t)n * 4);
    if (!d_tg || !d_meta || !d_loss) return h->fail(LOCR_ERR_CUDA, "allocation failed");
    int32_t* d_off = d_meta, *d_len = d_meta + n, *d_cnt = d_meta + 2 * n, *d_ok = d_meta + 3 * n;
    if (targets_total > 0) LOCR_CUDA_OK(cudaMemcpyAsync(d_tg, targets, (size_t)targets_total * 4, cudaMemcpyHostToDevice, s));
    LOCR_CUDA_OK(cudaMemcpyAsync(d_off, off.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
    LOCR_CUDA_OK(cudaMemcpyAsync(d_len, target_len, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    { ProfScope ps_(h, "eval_loss", 0, false);
    if (attn) launch_attn_ce(h->last_logits, n, C, d_tg, tw, h->last_ids, d_loss, d_cnt, d_ok, s);
    else launch_ctc_loss(h->last_logits, n, C, d_tg, d_off, d_len, h->last_ids, d_loss, d_ok, s); }
    h->launches++;
    LOCR_CUDA_OK(cudaGetLastError());
    std::vector<float> hl(n);
    std::vector<int32_t> hc(n), hk(n);
    LOCR_CUDA_OK(cudaMemcpyAsync(hl.data(), d_loss, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    LOCR_CUDA_OK(cudaMemcpyAsync(hk.data(), d_ok, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (attn) LOCR_CUDA_OK(cudaMemcpyAsync(hc.data(), d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    LOCR_CUDA_OK(cudaStreamSynchronize(s));
    // the batch scalar: CTCLoss 'mean' = mean over crops of loss / max(target length, 1); CrossEntropyLoss 'mean' =
    // sum over counted steps / their number (fp32 like torch)
    float acc = 0.f;
    int64_t cnt = 0;
    for (int i = 0; i < n; ++i) {
        if (attn) { acc += hl[i]; cnt += hc[i]; }
        else acc += hl[i] / (float)(target_len[i] > 1 ? target_len[i] : 1);
    }
    if (cost) *cost = attn ? acc / (float)cnt : acc / (float)n;   // 0 / 0 = NaN like torch for an all-padding batch
    if (loss) memcpy(loss, hl.data(), (size_t)n * 4);
    if (correct) memcpy(correct, hk.data(), (size_t)n * 4);
    return LOCR_OK;
}

LOCR_API int locr_recognize_boxes(locr_handle* h, const int32_t* image_index, const int32_t* rects, int n,
                                  float* logits, int32_t* token_ids, char* text, int32_t* has_eos, float* conf,
                                  uint8_t* resized_u8) {
    if (h == nullptr || image_index == nullptr || rects == nullptr || n <= 0)
        return fail(LOCR_ERR_INVALID, "locr_recognize_boxes: bad argument");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    std::vector<CropDesc> descs(n);
    auto py_slice = [](int start, int stop, int size, int* s0, int* len) {  // a[start:stop] for ints, step 1
        int a = start < 0 ? std::max(start + size, 0) : std::min(start, size);
        int b = stop < 0 ? std::max(stop + size, 0) : std::min(stop, size);
        *s0 = a;
        *len = b > a ? b - a : 0;
    };
    for (int i = 0; i < n; ++i) {
        const int ii = image_index[i];
        if (ii < 0 || ii >= (int)h->resident.size())
            return h->fail(LOCR_ERR_STATE, "locr_recognize_boxes: image index not resident (call locr_detect first)");
        const auto& im = h->resident[ii];
        int y0, hh, x0, ww;
        // sub = image[min_y:max_y, min_x:max_x, :]  (net.py:109-111, numpy slice semantics)
        py_slice(rects[i * 4 + 0], rects[i * 4 + 2], im.h, &y0, &hh);
        py_slice(rects[i * 4 + 1], rects[i * 4 + 3], im.w, &x0, &ww);
        CropDesc& d = descs[i];
        d.stride = (long)im.w * 3;
        d.src = im.p + (size_t)y0 * d.stride + (size_t)x0 * 3;
        d.h = (hh > 0 && ww > 0) ? hh : 0;
        d.w = (hh > 0 && ww > 0) ? ww : 0;
        d.channels = 3;
    }
    int rc = run_crops(h, descs, logits, token_ids, text, has_eos, conf, resized_u8);
    if (rc != LOCR_OK) return rc;
    if (has_eos)
        for (int i = 0; i < n; ++i)
            if (descs[i].h == 0) has_eos[i] = -2;  // empty crop: the reference's cv2.cvtColor would raise
    return LOCR_OK;
}

/* Post-processing only, on host score maps [B][H][W][2] (tests: bit-exact parity with det_boxes_core). */
LOCR_API int locr_debug_postproc(locr_handle* h, const float* score, int B, int H, int W, double ratio_w,
                                 double ratio_h, int max_boxes, float* boxes, int32_t* rects, int32_t* box_label,
                                 int32_t* counts, int32_t* labels) {
    if (h == nullptr || score == nullptr || B <= 0) return fail(LOCR_ERR_INVALID, "bad argument");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    cudaStream_t s = h->stream;
    const size_t npix = (size_t)B * H * W;
    float* d_sc = (float*)engine_buffer(h, "score", npix * 2 * 4);
    void* ws = engine_buffer(h, "pp.ws", postproc_workspace_bytes(B, H, W));
    float* d_boxes = (float*)engine_buffer(h, "pp.boxes", (size_t)B * max_boxes * 8 * 4);
    int32_t* d_rects = (int32_t*)engine_buffer(h, "pp.rects", (size_t)B * max_boxes * 4 * 4);
    int32_t* d_lab = (int32_t*)engine_buffer(h, "pp.lab", (size_t)B * max_boxes * 4);
    int32_t* d_counts = (int32_t*)engine_buffer(h, "pp.counts", (size_t)B * 2 * 4);
    int32_t* d_labels = labels ? (int32_t*)engine_buffer(h, "pp.labels", npix * 4) : nullptr;
    if (!d_sc || !ws || !d_boxes || !d_rects || !d_lab || !d_counts || (labels && !d_labels))
        return h->fail(LOCR_ERR_CUDA, "allocation failed");
    LOCR_CUDA_OK(cudaMemcpyAsync(d_sc, score, npix * 8, cudaMemcpyHostToDevice, s));
    PostprocParams pp;
    pp.B = B; pp.H = H; pp.W = W;
    pp.low_text = h->cfg.low_text; pp.link_threshold = h->cfg.link_threshold; pp.text_threshold = h->cfg.text_threshold;
    pp.scale_x = ratio_w * 2; pp.scale_y = ratio_h * 2; pp.max_boxes = max_boxes;
    const int nl = launch_postproc(d_sc, pp, ws, d_boxes, d_rects, d_lab, d_counts, d_labels, s);
    if (nl < 0) return h->fail(LOCR_ERR_INVALID, "score map larger than 1024 x 1024");
    h->launches += nl;
    LOCR_CUDA_OK(cudaGetLastError());
    LOCR_CUDA_OK(cudaMemcpyAsync(boxes, d_boxes, (size_t)B * max_boxes * 32, cudaMemcpyDeviceToHost, s));
    LOCR_CUDA_OK(cudaMemcpyAsync(rects, d_rects, (size_t)B * max_boxes * 16, cudaMemcpyDeviceToHost, s));
    LOCR_CUDA_OK(cudaMemcpyAsync(box_label, d_lab, (size_t)B * max_boxes * 4, cudaMemcpyDeviceToHost, s));
    LOCR_CUDA_OK(cudaMemcpyAsync(counts, d_counts, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
    if (labels) LOCR_CUDA_OK(cudaMemcpyAsync(labels, d_labels, npix * 4, cudaMemcpyDeviceToHost, s));
    LOCR_CUDA_OK(cudaStreamSynchronize(s));
    return LOCR_OK;
}

/* getDetBoxes(textmap, linkmap, text_threshold, link_threshold, low_text, poly) of the reference's tools package
 * (det_utils.py:248-256) on host score maps: boxes in score-map coordinates (det_boxes_core) and, with poly != 0, the
 * polygon of every box (poly_core, :97-245) or "none". */
LOCR_API int locr_get_det_boxes(locr_handle* h, const float* score, int B, int H, int W, float text_threshold,
                                float link_threshold, float low_text, int poly, int max_boxes, float* boxes,
                                int32_t* counts, double* polys, int32_t* poly_valid) {
    if (h == nullptr || score == nullptr || B <= 0 || max_boxes <= 0 || boxes == nullptr || counts == nullptr ||
        (poly && (polys == nullptr || poly_valid == nullptr)))
        return fail(LOCR_ERR_INVALID, "locr_get_det_boxes: bad argument");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    cudaStream_t s = h->stream;
    const size_t npix = (size_t)B * H * W;
    float* d_sc = (float*)engine_buffer(h, "score", npix * 2 * 4);
    void* ws = engine_buffer(h, "pp.ws", postproc_workspace_bytes(B, H, W));
    float* d_boxes = (float*)engine_buffer(h, "pp.boxes", (size_t)B * max_boxes * 8 * 4);
    int32_t* d_rects = (int32_t*)engine_buffer(h, "pp.rects", (size_t)B * max_boxes * 4 * 4);
    int32_t* d_lab = (int32_t*)engine_buffer(h, "pp.lab", (size_t)B * max_boxes * 4);
    int32_t* d_counts = (int32_t*)engine_buffer(h, "pp.counts", (size_t)B * 2 * 4);
    int32_t* d_labels = poly ? (int32_t*)engine_buffer(h, "pp.labels", npix * 4) : nullptr;
    double* d_polys = poly ? (double*)engine_buffer(h, "pp.polys", (size_t)B * max_boxes * 28 * 8) : nullptr;
    int32_t* d_pvalid = poly ? (int32_t*)engine_buffer(h, "pp.pvalid", (size_t)B * max_boxes * 4) : nullptr;
    if (!d_sc || !ws || !d_boxes || !d_rects || !d_lab || !d_counts || (poly && (!d_labels || !d_polys || !d_pvalid)))
        return h->fail(LOCR_ERR_CUDA, "allocation failed");
    LOCR_CUDA_OK(cudaMemcpyAsync(d_sc, score, npix * 8, cudaMemcpyHostToDevice, s));
    PostprocParams pp;
    pp.B = B; pp.H = H; pp.W = W;
    pp.low_text = low_text; pp.link_threshold = link_threshold; pp.text_threshold = text_threshold;
    pp.scale_x = 2.0; pp.scale_y = 2.0; pp.max_boxes = max_boxes;
    const int nl = launch_postproc(d_sc, pp, ws, d_boxes, d_rects, d_lab, d_counts, d_labels, s);
    if (nl < 0) return h->fail(LOCR_ERR_INVALID, "score map larger than 1024 x 1024");
    h->launches += nl;
    if (poly) {
        ProfScope ps_(h, "polys", 0, false);
        launch_polys(d_boxes, d_lab, d_counts, d_labels, B, H, W, max_boxes, d_polys, d_pvalid, s);
        h->launches++;
    }
    LOCR_CUDA_OK(cudaGetLastError());
    std::vector<int32_t> c2((size_t)B * 2);
    LOCR_CUDA_OK(cudaMemcpyAsync(c2.data(), d_counts, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
    LOCR_CUDA_OK(cudaMemcpyAsync(boxes, d_boxes, (size_t)B * max_boxes * 32, cudaMemcpyDeviceToHost, s));
    if (poly) {
        LOCR_CUDA_OK(cudaMemcpyAsync(polys, d_polys, (size_t)B * max_boxes * 28 * 8, cudaMemcpyDeviceToHost, s));
        LOCR_CUDA_OK(cudaMemcpyAsync(poly_valid, d_pvalid, (size_t)B * max_boxes * 4, cudaMemcpyDeviceToHost, s));
    }
    LOCR_CUDA_OK(cudaStreamSynchronize(s));
    for (int b = 0; b < B; ++b) {
        if (c2[2 * b] > max_boxes) return h->fail(LOCR_ERR_CAPACITY, "locr_get_det_boxes: max_boxes too small");
        counts[b] = c2[2 * b];
    }
    return LOCR_OK;
}

/* cv2.resize(INTER_LINEAR) alone (tests). */
LOCR_API int locr_debug_resize(locr_handle* h, const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw) {
    if (h == nullptr || src == nullptr || dst == nullptr) return fail(LOCR_ERR_INVALID, "bad argument");
    LOCR_CUDA_OK(cudaSetDevice(h->cfg.device_id));
    h->resident.clear();   // the image buffer is about to be overwritten (and possibly moved)
    uint8_t* d_s = (uint8_t*)engine_buffer(h, "images", (size_t)sh * sw * 3);
    uint8_t* d_d = (uint8_t*)engine_buffer(h, "resized", (size_t)dh * dw * 3);
    if (!d_s || !d_d) return h->fail(LOCR_ERR_CUDA, "allocation failed");
    LOCR_CUDA_OK(cudaMemcpyAsync(d_s, src, (size_t)sh * sw * 3, cudaMemcpyHostToDevice, h->stream));
    launch_resize_linear_bgr(d_s, 1, sh, sw, d_d, dh, dw, h->stream);
    h->launches++;
    LOCR_CUDA_OK(cudaMemcpyAsync(dst, d_d, (size_t)dh * dw * 3, cudaMemcpyDeviceToHost, h->stream));
    LOCR_CUDA_OK(cudaStreamSynchronize(h->stream));
    return LOCR_OK;
}

}  // extern "C"
