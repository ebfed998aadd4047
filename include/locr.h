/* liblocr - B200-native detect-then-recognize path of lightly-ocr behind a C ABI.
 *
 * The reference has no FFI of its own: its boundary is the Python class contract of ocr/net.py
 * (CRAFT :37-113, CRNN :116-193) as driven by ocr/pipeline.py:47-87.  Each entry point below names the reference
 * code it replaces; the Python module lightly_ocr_b200/net.py binds them with ctypes and re-exposes the reference's
 * CRAFT / CRNN classes unchanged.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative locr_status and
 * never throws; outputs are caller-allocated host buffers; one handle = one GPU = one CUDA stream; a handle is not
 * thread-safe, distinct handles are.  All calls are synchronous on return.  There is no CPU fallback: without a
 * CUDA device locr_create fails.
 */
#ifndef LOCR_H_
#define LOCR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LOCR_API __attribute__((visibility("default")))
#else
#define LOCR_API
#endif

typedef struct locr_handle locr_handle;

typedef enum {
    LOCR_OK = 0,
    LOCR_ERR_INVALID = -1,   /* bad argument / unsupported shape */
    LOCR_ERR_CUDA = -2,      /* CUDA runtime or driver error (message in locr_last_error) */
    LOCR_ERR_STATE = -3,     /* weights missing / not finalized */
    LOCR_ERR_CAPACITY = -4   /* caller-provided output capacity too small */
} locr_status;

enum { LOCR_MODEL_CRAFT = 0, LOCR_MODEL_CRNN = 1 };
enum { LOCR_HEAD_CTC = 0, LOCR_HEAD_ATTN = 1 };
enum { LOCR_ACT_F16 = 0, LOCR_ACT_BF16 = 1 };
/* Arithmetic of the recogniser (CRNN):
 *   LOCR_PREC_FAST  : 16-bit operands, fp32 accumulate - one tensor-core pass per layer (north-star precision policy);
 *   LOCR_PREC_EXACT : split precision - every activation and folded weight of the ResNet, the BiLSTM projections and
 *                     the prediction head is carried as a hi + lo pair of 16-bit numbers (~22 significant bits, three
 *                     tensor-core passes per product), which reproduces the reference's fp32 logits to ~1e-2 of their
 *                     trained scale at about twice the recognition time.  Detection (CRAFT) is unaffected. */
enum { LOCR_PREC_FAST = 0, LOCR_PREC_EXACT = 1 };

/* Mirrors the knobs of CRAFT.__init__ (ocr/net.py:45-50) and the config.yml keys read by CRNN (ocr/config.yml:27-45). */
typedef struct {
    int device_id;
    int act_dtype;        /* LOCR_ACT_*: storage type of activations and weights; accumulation is always fp32 */
    int head;             /* LOCR_HEAD_* (config.yml `prediction`) */
    int num_classes;      /* 37 for CTC, 38 for Attention (config.yml `num_classes`) */
    int canvas_size;      /* 1280  (net.py:45) */
    float mag_ratio;      /* 1.5   (net.py:46) */
    float text_threshold; /* 0.7   (net.py:47) */
    float link_threshold; /* 0.4   (net.py:48) */
    float low_text;       /* 0.4   (net.py:49) */
    int crnn_precision;   /* LOCR_PREC_* */
} locr_config;

LOCR_API const char* locr_version(void);
/* Error text of the last failing call on this thread (handle may be NULL for create/test hooks). */
LOCR_API const char* locr_last_error(const locr_handle* h);

/* CRAFT.__init__/CRNN.__init__ (net.py:38-51, :117-126): create the per-GPU engine. */
LOCR_API int locr_create(const locr_config* cfg, locr_handle** out);
LOCR_API void locr_destroy(locr_handle* h);

/* CRAFT.load / CRNN.load (net.py:59-69, :134-150): one call per state-dict entry, host fp32 data, reference key names
 * (an optional leading "module." is stripped like copyStateDict, net.py:24-34). */
LOCR_API int locr_load_tensor(locr_handle* h, int model, const char* key, const float* data, const int64_t* shape,
                              int ndim);
/* Folds BatchNorm into the convolutions, converts/lays out weights on the device.  Must follow the loads. */
LOCR_API int locr_finalize(locr_handle* h, int model);

/* CRAFT.process up to the rect list (net.py:100-107 = preproc :71-80, VGG_UNet forward model.py:39-61,
 * getDetBoxes det_utils.py:248-256, adjustResultCoordinates :259-265, getCoords net.py:82-98) for n images.
 *   bgr[i]     : uint8 HxWx3 (cv2.imread layout), row stride strides[i] bytes; packed images in pinned (page-locked)
 *                host memory are copied to the device straight from the caller's buffer, others are staged first
 *   rects      : [max_boxes_total][4] = (min_y, min_x, max_y, max_x) in image pixels, label order (unsorted; the
 *                reading-order sort compare_rects stays on the host), images concatenated
 *   boxes      : [max_boxes_total][4][2] float32 score-map-space corners as det_boxes_core returns them (may be NULL)
 *   box_counts : [n] number of boxes of each image
 *   score_maps : optional [sum over images of (H/2 * W/2 * 2)] fp32 (text, link interleaved like y[0,:,:,c]) */
LOCR_API int locr_detect(locr_handle* h, const uint8_t* const* bgr, const int* heights, const int* widths,
                         const int* strides, int n, int max_boxes_total, int32_t* rects, float* boxes,
                         int32_t* box_counts, float* score_maps);

/* pipeline.py:74-79 + CRNN.getPreds/process (net.py:152-193) for n crops at once.
 *   img[i] : uint8 crop, channels[i] = 3 (BGR, converted like cv2.cvtColor BGR2GRAY) or 1 (gray)
 *   logits : [n][26][num_classes] fp32 (`preds`) or NULL; token_ids [n][26]; text [n][128] NUL-terminated decoded
 *   string, stride 128 (CTC: collapsed; Attention: characters before the first [s]; has_eos = 1 found, 0 none, -1 [s]
 *   first = the reference raises IndexError);
 *   conf   : [n] cumulative product of per-step max softmax probabilities (net.py:190). */
LOCR_API int locr_recognize(locr_handle* h, const uint8_t* const* img, const int* heights, const int* widths,
                            const int* strides, const int* channels, int n, float* logits, int32_t* token_ids,
                            char* text, int32_t* has_eos, float* conf);

/* evaluation() of the reference's training script for ONE validation batch (ocr/train/crnn.py:142-240; SURVEY 8f row
 * 4): the crops are recognised like locr_recognize, then the loss and the accuracy flags are computed on the GPU from
 * the logits still in HBM.
 *   CTC head       : targets = concatenated class indices, target_len [n] (CTCLabelConverter.encode,
 *                    tools/recog_utils.py:24-30); loss[i] = torch.nn.CTCLoss(zero_infinity=True, reduction='none') of
 *                    preds.log_softmax(2) (crnn.py:119,190); *cost = its 'mean' reduction (loss / target length,
 *                    averaged over the batch)
 *   Attention head : targets = [n][batch_max_len + 2] rows of AttnLabelConverter.encode (:84-96: [GO], tokens, [s],
 *                    [GO] padding), target_len unused but required; loss[i] = sum over the non-padding steps of the
 *                    cross entropy against targets[i][1:], *cost = CrossEntropyLoss(ignore_index=0) of the batch
 *                    (crnn.py:121,203-208)
 *   correct[i]     : 1 iff the greedy prediction equals the label under evaluation()'s comparison (:222-230)
 *   token_ids / text / conf : as locr_recognize (any may be NULL). */
LOCR_API int locr_evaluate(locr_handle* h, const uint8_t* const* img, const int* heights, const int* widths,
                           const int* strides, const int* channels, int n, const int32_t* targets,
                           const int32_t* target_len, int64_t targets_total, float* loss, int32_t* correct,
                           int32_t* token_ids, char* text, float* conf, float* cost);

/* Second half of the fused throughput path (getText, pipeline.py:65-87): recognises boxes of the images of the LAST
 * locr_detect call without the pixels leaving the GPU.  The caller sorts the rects in between exactly like the
 * reference does on the host (sorted(rects, key=cmp_to_key(compare_rects)), net.py:108 - the comparator is not a
 * consistent order, so the result is defined by CPython's sort and stays in Python).
 *   image_index[i] : index into the image list passed to locr_detect;  rects[i] = (min_y, min_x, max_y, max_x)
 *   crops follow numpy slicing image[min_y:max_y, min_x:max_x] (net.py:109-111); an empty crop yields has_eos = -2
 *   (the reference's cv2.cvtColor raises on it).  resized_u8 (optional): [n][32][100] output of the BICUBIC resize. */
LOCR_API int locr_recognize_boxes(locr_handle* h, const int32_t* image_index, const int32_t* rects, int n,
                                  float* logits, int32_t* token_ids, char* text, int32_t* has_eos, float* conf,
                                  uint8_t* resized_u8);

/* tools.getDetBoxes(textmap, linkmap, text_threshold, link_threshold, low_text, poly) (det_utils.py:248-256) on B host
 * score maps [B][H][W][2] (text, link interleaved): `boxes` receives the float32 corners of det_boxes_core (:35-94) in
 * score-map coordinates, [B][max_boxes][4][2], `counts[b]` their number.  With poly != 0 the polygon path poly_core
 * (:97-245, what CRAFT.enablePoly switches on) runs as well: `polys` [B][max_boxes][14][2] float64 and `poly_valid`
 * [B][max_boxes] (1 = polygon, 0 = the reference's None).  Note: the reference's CRAFT.getCoords overwrites the
 * polygons with the boxes before using them (net.py:86-87), so enablePoly never changes what CRAFT.process returns;
 * this entry point serves callers of getDetBoxes itself. */
LOCR_API int locr_get_det_boxes(locr_handle* h, const float* score, int B, int H, int W, float text_threshold,
                                float link_threshold, float low_text, int poly, int max_boxes, float* boxes,
                                int32_t* counts, double* polys, int32_t* poly_valid);

/* locr_detect on the images left resident on the device by the previous locr_detect call (no host-to-device copy). */
LOCR_API int locr_detect_resident(locr_handle* h, int max_boxes_total, int32_t* rects, float* boxes,
                                  int32_t* box_counts, float* score_maps);

/* cv2.imread(path) / cv2.imdecode(buf, IMREAD_COLOR) for baseline JPEG files (pipeline.py:68), bit-exact
 * with OpenCV's libjpeg defaults (ISLOW inverse DCT, fancy up-sampling, 16-bit fixed-point YCbCr -> BGR).  The Huffman
 * entropy decoding runs on host threads, everything else on the GPU.  Covered: 8-bit Huffman-coded DCT files - baseline,
 * extended sequential (one or several scans) and progressive (spectral selection + successive approximation) - with 1
 * or 3 components, any integral sampling ratios, restart intervals, EXIF orientations 1-8 (applied like OpenCV's
 * ExifTransform; sizes are those of the rotated image).  Arithmetic-coded, lossless, 12-bit, CMYK and truncated files
 * return LOCR_ERR_INVALID (the caller falls back to its own reader and hands the pixels to locr_detect). */
LOCR_API int locr_jpeg_info(const uint8_t* data, int64_t nbytes, int* height, int* width, int* components);
/* PNG files (the reference server accepts .png uploads, ocr/server.py:11; cv2.imread at pipeline.py:68 hands them to
 * libpng): locr_imdecode / locr_detect_encoded take JPEG and PNG files alike, told apart by their signature.  Byte-exact
 * with cv2.imdecode(IMREAD_COLOR): every colour type and bit depth (gray / RGB / palette / alpha, 1-16 bit), both
 * interlace methods; 16-bit samples -> high byte, alpha and tRNS dropped, gamma chunks ignored, as OpenCV asks of
 * libpng.  zlib inflate runs on host threads, scanline un-filtering and sample conversion on the GPU.  Animated,
 * truncated or CRC-damaged files return LOCR_ERR_INVALID (the caller falls back to its own reader).
 * locr_image_info: header query for either format; *format (optional) receives LOCR_FORMAT_*. */
enum { LOCR_FORMAT_JPEG = 0, LOCR_FORMAT_PNG = 1 };
LOCR_API int locr_image_info(const uint8_t* data, int64_t nbytes, int* height, int* width, int* components,
                             int* format);
/* One file -> packed uint8 [height][width][3] BGR in the caller's host buffer (capacity in bytes). */
LOCR_API int locr_imdecode(locr_handle* h, const uint8_t* jpeg, int64_t nbytes, uint8_t* bgr, int64_t capacity,
                           int* height, int* width);
/* locr_detect on n encoded files: decoded on the GPU, the pixels never visit the host and stay resident for
 * locr_recognize_boxes.  heights / widths (optional, [n]) receive the image sizes. */
LOCR_API int locr_detect_encoded(locr_handle* h, const uint8_t* const* jpeg, const int64_t* nbytes, int n,
                                 int max_boxes_total, int32_t* rects, float* boxes, int32_t* box_counts,
                                 float* score_maps, int* heights, int* widths);

/* CUDA-event timing on the handle's own stream (bench.py; torch.cuda.Event only sees torch's streams). */
LOCR_API int locr_timer_start(locr_handle* h);
LOCR_API int locr_timer_stop(locr_handle* h, float* ms);
/* Per-launch CUDA-event profiling of the tensor-core convolution kernel: enable, run, then read the totals since the
 * last read (kernel milliseconds, algorithmic FLOPs = 2*M*N*K of the unpadded layers, launches). */
LOCR_API int locr_profile(locr_handle* h, int enable);
LOCR_API int locr_profile_read(locr_handle* h, double* conv_ms, double* conv_flops, int64_t* conv_launches);
/* Per-kernel totals accumulated by the locr_profile_read calls since the last call: text lines
 * "layer-or-kernel-name milliseconds algorithmic-flops launches" (tools/prof_pipeline.py). */
LOCR_API int locr_profile_layers(locr_handle* h, char* out, int64_t capacity);

/* Range audit of the 16-bit activation storage.  Conversions to fp16 / bf16 saturate (an activation beyond the fp16
 * range is stored as +-65504, never as an infinity), so an overflowing checkpoint degrades instead of turning into
 * NaNs - and this audit finds it: enable, run detection / recognition on representative inputs, read text lines
 * "layer-name abs-max" (one per convolution launch, in launch order; 16-bit outputs only).  A layer close to 65504 has
 * no head-room in fp16: run the handle with act_dtype = LOCR_ACT_BF16 (8 exponent bits) instead.  Costs one extra pass
 * over every activation while enabled. */
LOCR_API int locr_audit(locr_handle* h, int enable);
LOCR_API int locr_audit_read(locr_handle* h, char* out, int64_t capacity);

/* Kernel launches issued by this handle since creation (bench.py reports the per-step delta as gpu_launches). */
LOCR_API int64_t locr_launch_count(const locr_handle* h);

/* ---- kernel-level test hooks (used by tests/ only; same kernels as the product path) ---- */
typedef struct {
    int B, H, W, Cin, Cout;
    int KH, KW, dil_h, dil_w, pad_h, pad_w, stride_h;
    int x_pitch, y_pitch;      /* elements per pixel of the input / output buffers (>= Cin / Cout) */
    int relu, out_fp32, act_dtype, n_tile;
} locr_conv_desc;
/* x [B,H,W,x_pitch] fp32 NHWC, w [Cout,KH,KW,Cin] fp32, bias [Cout] or NULL, residual [B,OH,OW,Cout] or NULL,
 * y [B,OH,OW,y_pitch] fp32.  Inputs are rounded to the 16-bit activation type on the way in. */
LOCR_API int locr_test_conv(const locr_conv_desc* d, const float* x, const float* w, const float* bias,
                            const float* residual, float* y);

/* Same with the fused MaxPool2d(2, 2) of the activated output (vgg_bn.py / resnet50v1.py / TPS_STN.py max-pools that
 * follow a conv+BN+ReLU): y_pool [B,OH/2,OW/2,Cout] fp32; y may be NULL (only the pooled tensor is written). */
LOCR_API int locr_test_conv_pool(const locr_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                                 float* y_pool);

/* The BiLSTM recurrence kernel alone (replaces the cuDNN RNN behind nn.LSTM, reference ocr/modules/biLSTM.py:18,24).
 * xproj [B][T][2048] fp32 = W_ih x + b_ih + b_hh and whh [2][1024][256] fp32 in PyTorch's row order
 * (direction*1024 + gate*256 + unit, gates i, f, g, o); out [B][T][512] fp32 (forward | backward hidden states).
 * iters > 0 also times `iters` launches with CUDA events (ms per launch). */
LOCR_API int locr_test_lstm(const float* xproj, const float* whh, int B, int T, int act_dtype, float* out, int iters,
                            float* ms_per_iter, int split);

/* The evaluation-loss kernels of locr_evaluate alone, on host logits [n][26][C] (greedy ids come from the decode
 * kernel): head_attn = 0 CTC (targets concatenated, target_len [n]); head_attn = 1 attention cross entropy (targets
 * [n][targets_total / n]; count [n] = steps counted, may be NULL for CTC). */
LOCR_API int locr_test_eval_loss(int head_attn, const float* logits, int n, int C, const int32_t* targets,
                                 const int32_t* target_len, int64_t targets_total, float* loss, int32_t* count,
                                 int32_t* correct);

/* CRAFT forward only: bgr uint8 [B][img_h][img_w][3] packed -> score fp32 [B][H32/2][W32/2][2]. */
LOCR_API int locr_debug_craft_scores(locr_handle* h, const uint8_t* bgr, int B, int img_h, int img_w, float* score);
/* CRNN forward + decode on already resized crops: u8 [n][32][100] (output of ResizeNormalize's BICUBIC resize). */
LOCR_API int locr_debug_crnn(locr_handle* h, const uint8_t* u8, int n, float* logits, int32_t* ids, char* text,
                             int text_stride, int32_t* has_eos, float* conf);
/* Named intermediate of the last forward pass as dense fp32 (out may be NULL to query the shape only). */
LOCR_API int locr_debug_read(locr_handle* h, const char* name, float* out, int64_t capacity, int64_t* shape,
                             int* ndim);

/* det_boxes_core + adjustResultCoordinates + getCoords alone, on host score maps [B][H][W][2]; outputs per image with
 * capacity max_boxes: boxes [B][max_boxes][8], rects [B][max_boxes][4], box_label [B][max_boxes], counts [B][2]
 * (boxes kept, components found), labels (optional) [B][H][W] like cv2.connectedComponents. */
LOCR_API int locr_debug_postproc(locr_handle* h, const float* score, int B, int H, int W, double ratio_w,
                                 double ratio_h, int max_boxes, float* boxes, int32_t* rects, int32_t* box_label,
                                 int32_t* counts, int32_t* labels);
/* cv2.resize(src [sh][sw][3], (dw, dh), INTER_LINEAR) alone. */
LOCR_API int locr_debug_resize(locr_handle* h, const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw);

/* Host half of the JPEG reader alone (parsing + Huffman entropy decoding, no GPU): out = int16
 * [component][block row][block col][64] quantised coefficients in natural order (NULL: geometry only), info[19] = H, W,
 * components, hmax, vmax, MCUs per row, MCU rows, then (h, v, blocks per row, block rows) per component. */
LOCR_API int locr_test_jpeg_coefficients(const uint8_t* data, int64_t nbytes, int16_t* out, int64_t capacity, int* info);
/* Host half of the PNG reader alone (chunk walk, CRC, zlib inflate; no GPU): the filtered scanlines, pass after pass.
 * out may be NULL to query *need (bytes) only. */
LOCR_API int locr_test_png_scanlines(const uint8_t* data, int64_t nbytes, uint8_t* out, int64_t capacity, int64_t* need);

/* Number of convolutions this process has run in the split-K form (few output pixels, deep K: the K range is cut into
 * slices that run as separate CTAs, a second kernel adds the partial sums; lightly_ocr_b200/csrc/conv_tc.cuh).  Tests
 * use it to prove that the path under test was taken.  LOCR_TEST_SPLITK=0 makes the two conv harnesses above run
 * unsplit, LOCR_CONV_SPLITK=0 switches the form off everywhere. */
LOCR_API int64_t locr_test_splitk_calls(void);

/* Times one conv layer in isolation (zero-filled device buffers, CUDA events, `iters` launches after 3 warm-ups). */
LOCR_API int locr_bench_conv(const locr_conv_desc* d, int iters, float* ms_per_iter);
/* In-kernel timeline of the conv kernel (tools/conv_trace.py): only a library built with -DLOCR_CONV_EXPERIMENTS=1 and
 * run with LOCR_CONV_DBG bit 32 records anything.  out [3][8192] = (clock64 << 4 | event) of block 0's TMA producer,
 * MMA issuer and first epilogue warp; counts [3] = valid entries per role.  Resets the counters. */
LOCR_API int locr_conv_trace(unsigned long long* out, int* counts);

#ifdef __cplusplus
}
#endif
#endif /* LOCR_H_ */
