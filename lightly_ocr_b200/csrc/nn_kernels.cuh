// Non-GEMM neural-network kernels of the path (all HBM- or latency-bound): tiny-Cin direct convolution with the
// CRAFT normalisation fused, max-pool, bilinear 2x up-sampling into concat views, the TPS localisation head, TPS grid
// generation + grid_sample, the BiLSTM recurrence, the attention decoder and the CTC / attention token decode.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace locr {

// Split-precision tensors (hi + lo halves in channels [c] and [C + c], see conv_tc.cuh) are produced with split_out /
// consumed with split = 1; C is then the number of LOGICAL channels.
//
// 3x3 / pad 1 convolution for Cin in {1, 3} (reference layers basenet.slice1.0, LocalizationNetwork.conv.0,
// ConvNet.conv0_1).  w: fp32 [9*Cin][Cout] (BN scale folded), bias fp32 [Cout], out 16-bit NHWC.
//   u8 mode  : `in` is uint8 BGR [B][img_h][img_w][3] (row stride in bytes); the canvas is H x W, pixels outside the
//              image but inside the canvas read as 0 and everything is normalised (x - mean)/std on the fly exactly
//              like normalizeMeanVariance (imgproc.py:19-25) on the zero-padded canvas of resizeAspectRatio (:58-60).
//   f32 mode : `in` is fp32 [B][H][W] (Cin = 1).
void launch_direct_conv3x3(const void* in, int u8_mode, int B, int H, int W, int img_h, int img_w, long row_stride,
                           long img_stride, const float* w, const float* bias, int Cin, int Cout, void* out,
                           long out_pitch, int relu, int is_f16, cudaStream_t s, int split_out = 0, int pool = 0,
                           int out_row_px = 0);
//   pool     : f32 mode only; writes MaxPool2d(2, 2) of the activated output instead, [B][H/2][W/2][Cout]
//   out_row_px : f32 mode without pool, even H and W only; > 0 writes a row-padded tensor (pixel x lands at x + 1 of a
//              row of out_row_px = W + 3 pixels; the pad pixels are the caller's, see conv_tc.cuh x_row_px)

// CRAFT.preproc's normalisation (imgproc.py:19-25) of the zero-padded canvas: uint8 BGR -> 16-bit NHWC with 16
// channels (3 used), the input layout of the tensor-core path of basenet.slice1.0.  Output rows are padded to W + 3
// pixels (one zero pixel on the left, two on the right): out is [B][H][W + 3][16] (conv_tc.cuh: x_row_px).
// channels = 8: the same with 8 channels per pixel (16-byte pixels; half the bytes written and read back).
void launch_preproc_nhwc16(const uint8_t* in, int B, int H, int W, int img_h, int img_w, long row_stride,
                           long img_stride, void* out, int is_f16, cudaStream_t s, int channels = 16);

void launch_maxpool(const void* in, long in_pitch, int B, int H, int W, int C, void* out, long out_pitch, int kh,
                    int kw, int sh, int sw, int ph, int pw, int is_f16, cudaStream_t s, int split = 0);

// F.interpolate(scale 2, bilinear, align_corners=False) of [B,H,W,C] into a [B,2H,2W,*] view (model.py:47,51,55).
void launch_upsample2x(const void* in, long in_pitch, int B, int H, int W, int C, void* out, long out_pitch,
                       int is_f16, cudaStream_t s);

// AdaptiveAvgPool2d(1) + Linear(512,256)+ReLU + Linear(256,40)  (TPS_STN.py:57-60,70-76).
// feat 16-bit [B][hw][512]; w1t fp32 [512][256]; w2t fp32 [256][40]; fid fp32 [B][40].
void launch_loc_head(const void* feat, int B, int hw, const float* w1t, const float* b1, const float* w2t,
                     const float* b2, float* fid, int is_f16, cudaStream_t s, int split = 0);

// build_P_prime + grid_sample(bilinear, border, align_corners=True)  (TPS_STN.py:142-150, :27).
// inv_delta_c fp32 [23][23]; p_hat_t fp32 [23][3200] (transposed); x, out fp32 [B][32][100]; grid (optional) [B][3200][2].
void launch_tps_sample(const float* fid, const float* inv_delta_c, const float* p_hat_t, const float* x, float* out,
                       float* grid, int B, cudaStream_t s);

// BiLSTM recurrence (modules/biLSTM.py:18,24): cluster kernel with W_hh resident in shared memory (lstm_tc.cu).
//   xproj    fp32 [B][T][2048] = x W_ih^T + b_ih + b_hh, column = dir*1024 + unit*4 + gate (gate order i, f, g, o)
//   whh_perm 16-bit [2 dirs * 1024][256], rows in the same (unit, gate) order
//   out      16-bit [B][T][512] (forward | backward); also the medium through which the cluster exchanges h_t
//   split    LOCR_PREC_EXACT: out is [B][T][1024] = [hi 512 | lo 512], h_t is fed back as hi + lo (two GEMMs per step)
cudaError_t launch_lstm_tc(const float* xproj, const void* whh_perm, void* out, int B, int T, int is_f16,
                           cudaStream_t s, int split = 0);

struct AttnWeights {
    const uint16_t* h2h_wt;  // 16-bit [256 k][256 j]
    const float* h2h_b;      // [256]
    const float* score_w;    // [256]
    const uint16_t* wg;      // 16-bit [256 k][256 j][8] = LSTMCell weight_ih (context part) gates i,f,g,o then weight_hh
    const float* woh;        // fp32 [C][256 j][4 gates]: the one-hot part of weight_ih (one row is read per step)
    const float* gate_b;     // [4][256]  b_ih + b_hh
    const float* gen_w;    // [C][256]
    const float* gen_b;    // [C]
    // LOCR_PREC_EXACT: the two big matrices in fp32 (same layouts); used instead of the 16-bit copies when non-null
    const float* h2h_wt32 = nullptr;
    const float* wg32 = nullptr;
};
// Attention.forward inference branch (attention.py:46-59) with B=1 semantics per crop.
// feats 16-bit [B][26][256] (contextual features), fproj fp32 [B][26][256] (= i2h(feats), hoisted out of the loop),
// preds fp32 [B][26][C].
// feat_pitch: elements per time step of feats (0 = 256); feat_lo_off > 0: feats is a split-precision tensor whose lo
// halves live feat_lo_off elements further (value = hi + lo).
void launch_attention(const void* feats, const float* fproj, AttnWeights w, float* preds, int B, int C, int is_f16,
                      cudaStream_t s, long feat_pitch = 0, long feat_lo_off = 0);

// Range audit: largest |value| of a 16-bit tensor of `rows` pixels x `C` channels (channel pitch `pitch`), combined
// into *slot (fp32, must be zeroed by the caller) with atomicMax on the bit pattern of the non-negative maximum.
void launch_absmax(const void* t, long rows, int C, long pitch, int is_f16, float* slot, cudaStream_t s);

// Token decode + confidence (net.py:162-167,177-190; recog_utils.py:32-47,113-119).  logits fp32 [B][26][C].
// ids int32 [B][26]; text char [B][text_stride]; has_eos int32 [B] (CTC: always 1; Attention: 0 when no [s] was
// produced, -1 when the reference would raise IndexError because [s] is the first character); conf fp32 [B].
void launch_decode(const float* logits, int B, int C, int head_attn, int32_t* ids, char* text, int text_stride,
                   int32_t* has_eos, float* conf, cudaStream_t s);

// Evaluation losses of the reference's training script (ocr/train/crnn.py:142-240) on logits fp32 [B][26][C] and the
// greedy ids of launch_decode: per-crop CTC loss (CTCLoss(zero_infinity=True), unreduced) / attention cross-entropy
// sums and counts (CrossEntropyLoss(ignore_index=0)), and label == prediction flags.
void launch_ctc_loss(const float* logits, int B, int C, const int32_t* targets, const int32_t* tgt_off,
                     const int32_t* tgt_len, const int32_t* ids, float* loss, int32_t* correct, cudaStream_t s);
void launch_attn_ce(const float* logits, int B, int C, const int32_t* targets, int tw, const int32_t* ids, float* loss,
                    int32_t* count, int32_t* correct, cudaStream_t s);

}  // namespace locr
