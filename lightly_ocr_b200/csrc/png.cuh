// cv2.imread / cv2.imdecode(IMREAD_COLOR) for PNG files (reference ocr/pipeline.py:68; the server accepts .png uploads,
// ocr/server.py:11), byte-exact with OpenCV 4.13 + libpng 1.6: chunk parsing + zlib inflate on host threads (one
// serial bit stream per file), scanline un-filtering (anti-diagonal wavefront) and sample -> BGR conversion on the GPU.
// See png.cu.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>

struct locr_handle;

namespace locr {

// true when the buffer starts with the PNG signature
bool png_is_png(const uint8_t* data, size_t nbytes);

// Size from the header (all critical chunks are CRC-checked).  Returns 0 or LOCR_ERR_INVALID (reason in *err).
int png_probe(const uint8_t* data, size_t nbytes, int* height, int* width, int* components, std::string* err);

// Host half alone (tests): inflates the image data of one file into out = filtered scanlines, pass after pass
// (Adam7) or top to bottom; *need receives the byte count.  out may be NULL to query the size only.  No GPU involved.
int png_host_scanlines(const uint8_t* data, size_t nbytes, uint8_t* out, size_t capacity, size_t* need,
                       std::string* err);

// Decodes n files; image i lands as packed uint8 [H][W][3] BGR at d_out[i] (device memory, sized with png_probe).
// Work is queued on the handle's stream; the host-side inflate is finished and the stream synchronised on return.
int png_decode_to_device(locr_handle* h, const uint8_t* const* blobs, const int64_t* nbytes, int n,
                         uint8_t* const* d_out);

}  // namespace locr
