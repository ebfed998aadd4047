// Byte-exact image operators of the path, batched on the GPU:
//   * crop -> cv2 BGR2GRAY -> PIL BICUBIC resize to 100x32 -> ToTensor -> (x-0.5)/0.5   (reference ocr/net.py:109-111,
//     ocr/pipeline.py:75, ocr/net.py:155-158, ocr/tools/dataset.py:43-47)
//   * cv2.resize(INTER_LINEAR) of the uint8 BGR image for CRAFT's resizeAspectRatio (ocr/tools/imgproc.py:51)
// Compiled with --fmad=false: coefficient arithmetic must match Pillow's / OpenCV's x86 code bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace locr {

struct CropDesc {
    const uint8_t* src;  // first pixel of the crop (device)
    long stride;         // bytes per source row
    int h, w;            // crop size in pixels (0 = empty crop: output is zero-filled)
    int channels;        // 3 = BGR (converted to gray), 1 = gray
    int ksh, ksv;        // Pillow kernel sizes of the horizontal / vertical pass
    long coef_off;       // int32 offset into the coefficient scratch: [100*ksh][32*ksv][100*2][32*2]
    long inter_off;      // byte offset into the intermediate scratch: [h][100] uint8
};

// Fills ksh/ksv and returns the scratch needs of one crop (ints of coefficient scratch, bytes of intermediate).
void crop_scratch_sizes(int h, int w, int* ksh, int* ksv, long* coef_ints, long* inter_bytes);

// out_x: fp32 [n][32][100] normalised; out_u8 (optional): uint8 [n][32][100] resized image before normalisation.
void launch_crop_resize(const CropDesc* d_descs, int n, int32_t* coef_scratch, uint8_t* inter_scratch, float* out_x,
                        uint8_t* out_u8, cudaStream_t s);

// cv2.resize(src [B][sh][sw][3] uint8 packed, (dw, dh), INTER_LINEAR) -> dst [B][dh][dw][3] packed.
void launch_resize_linear_bgr(const uint8_t* src, int B, int sh, int sw, uint8_t* dst, int dh, int dw, cudaStream_t s);

}  // namespace locr
