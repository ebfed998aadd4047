// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA-fed).
// Host-side description of one layer invocation; see conv_tc.cu for the kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace locr {

enum ActDType { ACT_F16 = 0, ACT_BF16 = 1 };

// One convolution call: y = act( conv(x, w) + bias [+ residual] ).  BN scale is folded into w on load.
//   x : NHWC activation view, 16-bit elements, channel pitch x_pitch (>= Cin; concat buffers are views)
//   w : [Cout_pad][KH*KW*Cin] K-major, k = (kh*KW + kw)*Cin + c, 16-bit elements (same type as x)
//   y : NHWC, channel pitch y_pitch, 16-bit (same type as x) or fp32 when out_fp32
struct ConvCall {
    const void* x = nullptr;
    int B = 0, H = 0, W = 0, Cin = 0;
    long x_pitch = 0;
    const void* w = nullptr;
    int Cout = 0, Cout_pad = 0;
    int KH = 1, KW = 1, dil_h = 1, dil_w = 1, pad_h = 0, pad_w = 0, stride_h = 1;  // stride_w is always 1 on this path
    void* y = nullptr;
    int OH = 0, OW = 0;
    long y_pitch = 0;
    int out_fp32 = 0;
    const float* bias = nullptr;   // [Cout_pad] fp32
    const void* residual = nullptr;  // NHWC at output resolution, 16-bit, pitch res_pitch
    long res_pitch = 0;
    long res_lo_off = 0;  // > 0: the residual is a split-precision tensor, its lo halves live res_lo_off channels further
    int relu = 0;
    int dtype = ACT_BF16;
    int n_tile = 0;  // 0 = choose automatically
    // Split-precision (hi + lo) tensors: a value v is stored as two 16-bit numbers hi = round(v), lo = round(v - hi) in
    // channels [c] and [C + c] of a 2C-channel tensor (~22 significant bits).
    int cin_wrap = 0;   // > 0: the K channels of a tap are [hi | lo | hi] = 3C, read from a tensor of cin_wrap = 2C channels
    int split_out = 0;  // 1: write hi to channel n and lo to channel Cout + n of a 2*Cout-channel output
    // Fused MaxPool2d(2, 2) of the activated output (nn.MaxPool2d(kernel_size=2, stride=2) right after the conv's ReLU
    // in vgg16_bn / ResNet / LocalizationNetwork): pooled tensor [B][OH/2][OW/2][Cout] with channel pitch pool_pitch.
    void* pool_y = nullptr;
    long pool_pitch = 0;
    int skip_full = 0;  // 1: only the pooled tensor is written (y may be null)
    // Fused 1x1 tail (CRAFT conv_cls.6 + ReLU + conv_cls.8 after conv_cls.4 + ReLU, model.py:33-36): with Cout = 16 the
    // epilogue thread that owns a pixel applies relu(W6 h + b6) and W8 . + b8 in fp32 registers and writes the two
    // score channels.  tail_w: fp32 [16*16 W6 row-major | 16 b6 | 2*16 W8 | 2 b8]; tail_out: fp32 [B][OH][OW][2].
    const float* tail_w = nullptr;
    float* tail_out = nullptr;
    // Row-padded tensors and the "window" input view (small-Cin 3x3 layers).  The TMA unit spends ~2.5 cycles per box
    // row whatever the row's size, so a 3x3 conv over 16 or 32 channels (32/64-byte rows, 9 taps) is bound by the
    // row rate.  With x_row_px = W + 3 (one zero pixel left, two right of every row) the three horizontal taps of an
    // output pixel are ONE contiguous run of memory: the tensor map's innermost dimension becomes Cin = 4 pixels x
    // channels (window starting at pixel x - 1; the 4th pixel meets zero weights) with the pixel pitch as its stride,
    // KW = 1 and pad_w = 0: 3 (or 6) k-blocks of 128-byte rows instead of 9.
    long x_row_px = 0;  // pixels per memory row of x (0 = W)
    long y_row_px = 0;  // pixels per memory row of y (0 = OW); TMA-store epilogue only
    // First layer of CRAFT straight from the uint8 image (basenet.slice1.0 with normalizeMeanVariance, reference
    // tools/imgproc.py:19-25, fused): when first_u8 is set, x is unused and the A operand is built inside the kernel - four
    // producer warps gather the 3x3x3 neighbourhood of every output pixel from the packed BGR image (conv padding reads
    // as 0, canvas padding outside the image as the normalised value of 0), normalise through a 768-entry table and
    // write K = 27 (+5 zero) 16-bit values per pixel in the tensor core's 64-byte-swizzled K-major layout.
    // w is then [Cout_pad][32] with k = (ky * 3 + kx) * 3 + c; H x W is the canvas (= output) size.
    const uint8_t* first_u8 = nullptr;
    int img_h = 0, img_w = 0;
    long img_row_stride = 0, img_stride = 0;   // bytes per image row / per image
    // Split-K for calls with very few output pixels (one crop or a handful: M <= 8 tiles of 128 pixels), where a deep
    // layer (e.g. 512 -> 512, 3x3: 72 k-blocks) would otherwise run as one or two CTAs walking the whole K range while
    // 146 SMs idle: with a workspace the launcher cuts the input channels into up to 8 slices, runs slices x n-tiles
    // (N = 64) CTAs that store fp32 partial sums, and a second small kernel adds the slices, the bias and the residual,
    // applies the ReLU and rounds to 16 bits.  Plain 16-bit-output layers only (no pooling, split precision, fused tail).
    void* splitk_ws = nullptr;       // fp32 workspace, >= B * OH * OW * slices * Cout_pad * 4 bytes; null = never split
    size_t splitk_ws_bytes = 0;
    int ksplit = 0;                  // internal: set by conv_tc_launch on the partial-sum call it issues
};

// Returns cudaSuccess or the launch/encode error; writes a human-readable reason into err (if non-null).
cudaError_t conv_tc_launch(const ConvCall& c, cudaStream_t stream, char* err, int errlen);

// Experiments build (-DLOCR_CONV_EXPERIMENTS=1, LOCR_CONV_DBG bit 32): clock64 stamps of block 0's producer / MMA issuer /
// first epilogue warp, [3][8192] entries of (clock << 4 | event); counts[3] entries are valid.  Resets the counters.
int conv_tc_trace_read(unsigned long long* out, int* counts);

// Split-K convolutions launched by this process so far (tests: proves that the path under test was taken).
long long conv_tc_splitk_calls();

// Number of SMs used for the persistent grid (queried once).
int device_sm_count();

}  // namespace locr
