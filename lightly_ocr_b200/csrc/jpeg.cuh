// cv2.imread / cv2.imdecode(IMREAD_COLOR) for baseline JPEG files (reference ocr/pipeline.py:68), bit-exact with
// libjpeg's default decoder: Huffman entropy decoding on host threads, dequantisation + "ISLOW" integer inverse DCT,
// "fancy" chroma up-sampling and YCbCr -> BGR conversion on the GPU.  See jpeg.cu.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>

struct locr_handle;

namespace locr {

// Size and component count from the file header alone.  Returns 0 or LOCR_ERR_INVALID (reason in *err).
int jpeg_probe(const uint8_t* data, size_t nbytes, int* height, int* width, int* components, std::string* err);

// Host half alone (tests): parses the file and entropy-decodes it into out = int16 [component][block row][block col][64]
// (natural order).  info[19] = H, W, components, hmax, vmax, MCUs per row, MCU rows, then (h, v, blocks per row, block
// rows) per component.  out may be NULL to query the geometry only.  No GPU involved.
int jpeg_host_coefficients(const uint8_t* data, size_t nbytes, int16_t* out, size_t capacity, int* info,
                           std::string* err);

// Decodes n files; image i lands as packed uint8 [H][W][3] BGR at d_out[i] (device memory, the caller sized it with
// jpeg_probe).  Work is queued on the handle's stream; the host-side entropy decoding is finished on return.
int jpeg_decode_to_device(locr_handle* h, const uint8_t* const* blobs, const int64_t* nbytes, int n,
                          uint8_t* const* d_out);

}  // namespace locr
