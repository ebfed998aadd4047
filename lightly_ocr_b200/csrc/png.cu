// PNG reader of the ingest path: `cv2.imread(path)` of reference ocr/pipeline.py:68 for the `.png` uploads the reference
// server accepts (ocr/server.py:11).  OpenCV 4.13 hands PNG files to libpng 1.6 and asks for: 16-bit samples chopped to
// their high byte, palette -> RGB, gray 1/2/4 bit scaled to 8, alpha / tRNS dropped without blending, gray replicated to
// three channels, RGB -> BGR; gamma / colour-profile chunks are ignored.  None of that code is under /root/reference:
// the PNG specification (chunk layout, CRC-32, the five scanline filters, Adam7) is restated here in its own form and
// pinned byte for byte against the live cv2.imdecode (oracle/png_ref.py, tests/test_png_*.py).
//
// Split of the work:
//   host  : chunk walk + CRC of the critical chunks, zlib inflate of the IDAT stream (one serial bit stream per file,
//           so one host thread per image of the batch) straight into pinned memory;
//   GPU   : png_unfilter_kernel - the filters make byte (x, y) depend on (x-bpp, y), (x, y-1) and (x-bpp, y-1); all
//           units on an anti-diagonal are independent, so one thread per scanline runs one step behind the thread
//           above it and the reconstructed unit travels down through shared memory (1024 rows per block and sweep);
//           png_color_kernel - one thread per output pixel: Adam7 pass lookup, sample extraction (1/2/4/8/16 bit),
//           palette lookup, gray replication, RGB -> BGR, packed 3-byte store into the resident image buffer.
#include "png.cuh"

#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <atomic>
#include <thread>
#include <vector>

#include "engine.cuh"

namespace locr {

namespace {

const uint8_t kSig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
// Adam7 pass geometry: x start, y start, x step, y step
const int kAdam[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};

struct Span { const uint8_t* p; size_t n; };

// Device layout of one pass: every filtered scanline sits in a 16-byte aligned slot of in_pitch bytes with its filter
// byte at offset 15 and its data from offset 16 on; every unfiltered scanline in a slot of out_pitch bytes.  The aligned
// slots are what lets one thread move its scanline with 16-byte loads and stores.
struct PassGeo {
    int w = 0, h = 0;        // pixels of this pass (0: pass absent)
    size_t rowbytes = 0;     // unfiltered bytes per scanline
    size_t in_pitch = 0, out_pitch = 0;
    size_t in_off = 0;       // offset of the pass in the padded filtered buffer
    size_t out_off = 0;      // offset of the pass in the unfiltered buffer
};

struct Header {
    int W = 0, H = 0, depth = 0, color = 0, interlace = 0, channels = 0, bpp = 1;
    int npal = 0;
    uint8_t pal[768];
    std::vector<Span> idat;
    int npass = 1;
    PassGeo pass[7];
    size_t filtered_bytes = 0;   // contiguous stream: sum of (1 + rowbytes) over all scanlines
    size_t padded_bytes = 0;     // the same scanlines in their aligned slots
    size_t plain_bytes = 0;      // unfiltered slots
};

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

size_t max_image_pixels() {
    static size_t v = 0;
    if (v == 0) {
        const char* e = getenv("LOCR_MAX_IMAGE_PIXELS");
        v = e ? (size_t)strtoull(e, nullptr, 10) : 0;
        if (v == 0) v = (size_t)1 << 30;     // cv2.imread refuses larger images (CV_IO_MAX_IMAGE_PIXELS)
    }
    return v;
}

// Chunk walk up to IEND.  Critical chunks (IHDR, PLTE, IDAT, IEND) must carry a valid CRC, like in libpng; ancillary
// chunks are skipped (tRNS only adds an alpha channel that IMREAD_COLOR drops again; gAMA / sRGB / iCCP are not applied).
bool parse(const uint8_t* d, size_t n, Header* hd, std::string* err) {
    auto bad = [&](const char* m) { *err = std::string("PNG: ") + m; return false; };
    if (n < 8 || memcmp(d, kSig, 8) != 0) return bad("not a PNG file (bad signature)");
    size_t pos = 8;
    bool have_hdr = false, have_iend = false, have_plte = false, idat_closed = false;
    while (pos + 12 <= n) {
        const size_t len = be32(d + pos);
        const uint8_t* type = d + pos + 4;
        if (len > 0x7fffffffu || pos + 12 + len > n) return bad("truncated chunk");
        const uint8_t* body = d + pos + 8;
        const bool critical = (type[0] & 0x20) == 0;
        if (critical) {
            const uint32_t crc = (uint32_t)crc32(crc32(0L, type, 4), body, (uInt)len);
            if (crc != be32(body + len)) return bad("CRC mismatch in a critical chunk");
        }
        pos += 12 + len;
        if (memcmp(type, "IHDR", 4) == 0) {
            if (have_hdr || len != 13) return bad("bad IHDR");
            const uint32_t w = be32(body), h = be32(body + 4);
            hd->depth = body[8]; hd->color = body[9]; hd->interlace = body[12];
            if (w == 0 || h == 0 || w > 0x7fffffffu || h > 0x7fffffffu || body[10] != 0 || body[11] != 0 || hd->interlace > 1)
                return bad("bad IHDR");
            const int dpt = hd->depth;
            bool ok = false;
            switch (hd->color) {
                case 0: hd->channels = 1; ok = dpt == 1 || dpt == 2 || dpt == 4 || dpt == 8 || dpt == 16; break;
                case 2: hd->channels = 3; ok = dpt == 8 || dpt == 16; break;
                case 3: hd->channels = 1; ok = dpt == 1 || dpt == 2 || dpt == 4 || dpt == 8; break;
                case 4: hd->channels = 2; ok = dpt == 8 || dpt == 16; break;
                case 6: hd->channels = 4; ok = dpt == 8 || dpt == 16; break;
                default: break;
            }
            if (!ok) return bad("bad colour type / bit depth");
            if ((size_t)w * (size_t)h > max_image_pixels() || w > (1u << 20) || h > (1u << 20))
                return bad("image larger than the pixel limit (OpenCV's CV_IO_MAX_IMAGE_PIXELS, 2^30; LOCR_MAX_IMAGE_PIXELS)");
            hd->W = (int)w; hd->H = (int)h;
            have_hdr = true;
        } else if (!have_hdr) {
            return bad("IHDR is not the first chunk");
        } else if (memcmp(type, "PLTE", 4) == 0) {
            if (have_plte || !hd->idat.empty() || len == 0 || len % 3 != 0 || len > 768) return bad("bad PLTE");
            memcpy(hd->pal, body, len);
            hd->npal = (int)(len / 3);
            have_plte = true;
        } else if (memcmp(type, "IDAT", 4) == 0) {
            if (idat_closed) return bad("IDAT chunks are not consecutive");
            hd->idat.push_back({body, len});
        } else if (memcmp(type, "IEND", 4) == 0) {
            have_iend = true;
            break;
        } else {
            if (critical) return bad("unknown critical chunk");
            if (memcmp(type, "acTL", 4) == 0) return bad("animated PNG files are not supported");
            if (!hd->idat.empty()) idat_closed = true;
        }
    }
    if (!have_hdr || hd->idat.empty()) return bad("missing IHDR / IDAT");
    if (!have_iend) return bad("missing IEND (truncated file)");
    if (hd->color == 3 && !have_plte) return bad("palette image without PLTE");
    const int bits_pp = hd->depth * hd->channels;
    hd->bpp = bits_pp >= 8 ? bits_pp / 8 : 1;
    size_t stream = 0, in_off = 0, out_off = 0;
    hd->npass = hd->interlace ? 7 : 1;
    for (int p = 0; p < hd->npass; ++p) {
        PassGeo& g = hd->pass[p];
        const int pw = hd->interlace ? (hd->W - kAdam[p][0] + kAdam[p][2] - 1) / kAdam[p][2] : hd->W;
        const int ph = hd->interlace ? (hd->H - kAdam[p][1] + kAdam[p][3] - 1) / kAdam[p][3] : hd->H;
        if (pw <= 0 || ph <= 0) continue;
        g.w = pw; g.h = ph;
        g.rowbytes = ((size_t)pw * bits_pp + 7) / 8;
        g.in_pitch = (g.rowbytes + 16 + 15) / 16 * 16;
        g.out_pitch = (g.rowbytes + 15) / 16 * 16;
        g.in_off = in_off;
        g.out_off = out_off;
        stream += (g.rowbytes + 1) * (size_t)ph;
        in_off += g.in_pitch * (size_t)ph;
        out_off += g.out_pitch * (size_t)ph;
    }
    hd->filtered_bytes = stream;
    hd->padded_bytes = in_off;
    hd->plain_bytes = out_off;
    // deflate cannot expand by more than ~1032 : 1: a file whose IDAT data is too short for the declared size is
    // refused here, before any buffer is sized from the header
    size_t zbytes = 0;
    for (auto& s : hd->idat) zbytes += s.n;
    if (hd->filtered_bytes / 1032 > zbytes + 16) return bad("not enough image data for the declared size");
    return true;
}

// zlib inflate of the concatenated IDAT bodies.  padded = false: the contiguous stream out[0 .. filtered_bytes)
// (tests); padded = true: every scanline into its aligned slot (filter byte at offset 15 of the slot).
bool inflate_idat(const Header& hd, uint8_t* out, bool padded, std::string* err) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit(&zs) != Z_OK) { *err = "PNG: inflateInit failed"; return false; }
    size_t span = 0, off = 0;
    bool fail = false, short_data = false, ended = false;
    auto segment = [&](uint8_t* dst, size_t n) {
        while (n > 0 && !fail && !short_data) {
            const size_t now = n > (1u << 30) ? (1u << 30) : n;
            zs.next_out = dst;
            zs.avail_out = (uInt)now;
            while (zs.avail_out > 0) {
                if (ended) { short_data = true; break; }
                if (zs.avail_in == 0) {
                    while (span < hd.idat.size() && off >= hd.idat[span].n) { ++span; off = 0; }
                    if (span >= hd.idat.size()) { short_data = true; break; }
                    const size_t in_now = hd.idat[span].n - off > (1u << 30) ? (1u << 30) : hd.idat[span].n - off;
                    zs.next_in = const_cast<Bytef*>(hd.idat[span].p + off);
                    zs.avail_in = (uInt)in_now;
                    off += in_now;
                }
                const int rc = inflate(&zs, Z_NO_FLUSH);
                if (rc == Z_STREAM_END) ended = true;
                else if (rc != Z_OK && rc != Z_BUF_ERROR) { fail = true; break; }
            }
            const size_t got = now - zs.avail_out;
            dst += got;
            n -= got;
            if (zs.avail_out > 0) break;
        }
    };
    if (!padded) {
        segment(out, hd.filtered_bytes);
    } else {
        for (int p = 0; p < hd.npass && !fail && !short_data; ++p) {
            const PassGeo& g = hd.pass[p];
            for (int y = 0; y < g.h && !fail && !short_data; ++y)
                segment(out + g.in_off + (size_t)y * g.in_pitch + 15, g.rowbytes + 1);
        }
    }
    inflateEnd(&zs);
    if (fail) { *err = "PNG: corrupt compressed data"; return false; }
    if (short_data) { *err = "PNG: not enough image data"; return false; }   // libpng: too MUCH data is only a warning
    return true;
}

// ---------------------------------------------------------------------------------------------------- device side
struct UnfilterJob {
    const uint8_t* in;    // filtered scanlines of one pass in aligned slots: [h][in_pitch], filter byte at 15, data at 16
    uint8_t* out;         // unfiltered: [h][out_pitch]
    int h;
    int rowbytes;
    int bpp;
    int in_pitch, out_pitch;
};

__device__ __forceinline__ int paeth(int a, int b, int c) {
    const int p = a + b - c;
    const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// One block per (image, pass).  Thread r owns scanline y = g * 1024 + r of sweep g and walks it unit by unit (unit = bpp
// bytes), one step behind the thread above: at step s it reconstructs unit x = s - r from its own previous unit (a), the
// unit above (b, published by thread r - 1 one step earlier through shared memory) and the one above-left (c = last
// step's b).  Every step ends in a block-wide barrier and the 1024 scanlines of a sweep are ~3 KB apart in memory, so
// byte-wide accesses would cost one memory transaction per thread, byte and step (measured: 3 us per step, 12.7 ms per
// 1280x960 image).  Each thread therefore streams its scanline through registers in 16-byte pieces: aligned 16-byte
// loads two pieces ahead of the byte it is working on (the latency of a miss hides behind ~10 steps), bytes shifted
// out of / into a 128-bit window, one aligned 16-byte store per 16 reconstructed bytes.
struct ByteWindow {
    unsigned long long lo, hi;
    __device__ __forceinline__ unsigned get(int p) const {
        return (unsigned)((p < 8 ? lo >> (8 * p) : hi >> (8 * (p - 8))) & 255ull);
    }
    __device__ __forceinline__ void put(int p, unsigned v) {
        if (p < 8) lo |= (unsigned long long)v << (8 * p);
        else hi |= (unsigned long long)v << (8 * (p - 8));
    }
};

__device__ __forceinline__ ByteWindow load16(const uint8_t* p) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    ByteWindow w;
    w.lo = (unsigned long long)u.x | ((unsigned long long)u.y << 32);
    w.hi = (unsigned long long)u.z | ((unsigned long long)u.w << 32);
    return w;
}

__global__ void __launch_bounds__(1024)
png_unfilter_kernel(const UnfilterJob* __restrict__ jobs, int* __restrict__ error_flag) {
    __shared__ unsigned long long xch[2][1024];
    const UnfilterJob J = jobs[blockIdx.x];
    const int r = threadIdx.x;
    const int units = (J.rowbytes + J.bpp - 1) / J.bpp;
    const int pieces = (J.rowbytes + 15) >> 4;
    for (int y0 = 0; y0 < J.h; y0 += 1024) {
        const int rows = min(1024, J.h - y0);
        const int y = y0 + r;
        const bool active = r < rows;
        const uint8_t* slot = J.in + (size_t)(active ? y : 0) * (size_t)J.in_pitch;
        const uint8_t* src = slot + 16;
        uint8_t* dst = J.out + (size_t)(active ? y : 0) * (size_t)J.out_pitch;
        const uint8_t* above = J.out + (size_t)(y > 0 && active ? y - 1 : 0) * (size_t)J.out_pitch;  // r == 0 of a later sweep
        const int ft = active ? slot[15] : 0;
        if (active && ft > 4) atomicExch(error_flag, 1);
        ByteWindow cur, n1, n2, outw;
        cur.lo = cur.hi = n1.lo = n1.hi = n2.lo = n2.hi = outw.lo = outw.hi = 0ull;
        if (active) {
            cur = load16(src);
            if (pieces > 1) n1 = load16(src + 16);
            if (pieces > 2) n2 = load16(src + 32);
        }
        int ip = 0, piece = 0, op = 0, opiece = 0, i = 0;
        unsigned long long a = 0, c = 0;
        const int steps = units + rows - 1;
        for (int s = 0; s < steps; ++s) {
            const int x = s - r;
            if (active && x >= 0 && x < units) {
                unsigned long long b = 0;
                if (r > 0) b = xch[(s - 1) & 1][r - 1];
                else if (y > 0) {
#pragma unroll 1
                    for (int k = 0; k < J.bpp; ++k)
                        if (i + k < J.rowbytes) b |= (unsigned long long)above[i + k] << (8 * k);
                }
                unsigned long long v = 0;
                for (int k = 0; k < J.bpp && i < J.rowbytes; ++k, ++i) {
                    const int raw = (int)cur.get(ip);
                    if (++ip == 16) {           // next 16-byte piece; fetch the one after the two already in flight
                        ip = 0;
                        ++piece;
                        cur = n1;
                        n1 = n2;
                        if (piece + 2 < pieces) n2 = load16(src + (size_t)(piece + 2) * 16);
                    }
                    const int ak = (int)((a >> (8 * k)) & 255), bk = (int)((b >> (8 * k)) & 255), ck = (int)((c >> (8 * k)) & 255);
                    int pred = 0;
                    if (ft == 1) pred = ak;
                    else if (ft == 2) pred = bk;
                    else if (ft == 3) pred = (ak + bk) >> 1;
                    else if (ft == 4) pred = paeth(ak, bk, ck);
                    const unsigned val = (unsigned)(raw + pred) & 255u;
                    v |= (unsigned long long)val << (8 * k);
                    outw.put(op, val);
                    if (++op == 16 || i + 1 == J.rowbytes) {
                        *reinterpret_cast<uint4*>(dst + (size_t)opiece * 16) =
                            make_uint4((unsigned)outw.lo, (unsigned)(outw.lo >> 32), (unsigned)outw.hi, (unsigned)(outw.hi >> 32));
                        outw.lo = outw.hi = 0ull;
                        op = 0;
                        ++opiece;
                    }
                }
                xch[s & 1][r] = v;
                a = v;
                c = b;
            }
            __syncthreads();
        }
        // the next sweep's first thread reads this sweep's last scanline from global memory (same block: the barrier
        // of the final step has made the stores visible)
    }
}

struct ColorJob {
    const uint8_t* plain;     // unfiltered passes of this image
    uint8_t* out;             // packed BGR [H][W][3]
    const uint8_t* pal;       // palette (RGB triples) or nullptr
    int npal;
    int W, H, depth, color, channels, interlace;
    int pass_pitch[7];        // bytes per unfiltered scanline slot of each pass
    long pass_off[7];
};

__global__ void __launch_bounds__(256)
png_color_kernel(const ColorJob* __restrict__ jobs, int* __restrict__ error_flag) {
    const ColorJob& J = jobs[blockIdx.z];
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= J.W || y >= J.H) return;
    int p = 0, px = x, py = y;
    if (J.interlace) {
        // Adam7: the pass of a pixel depends on (x mod 8, y mod 8) only
        const int xm = x & 7, ym = y & 7;
        if (ym & 1) { p = 6; px = x; py = y >> 1; }
        else if (xm & 1) { p = 5; px = x >> 1; py = y >> 1; }
        else if (ym & 2) { p = 4; px = x >> 1; py = y >> 2; }
        else if (xm & 2) { p = 3; px = x >> 2; py = y >> 2; }
        else if (ym & 4) { p = 2; px = x >> 2; py = y >> 3; }
        else if (xm & 4) { p = 1; px = x >> 3; py = y >> 3; }
        else { p = 0; px = x >> 3; py = y >> 3; }
    }
    const uint8_t* row = J.plain + J.pass_off[p] + (size_t)py * (size_t)J.pass_pitch[p];
    int s[3];
    if (J.depth == 8) {
        const uint8_t* q = row + (size_t)px * J.channels;
        s[0] = q[0];
        if (J.color == 2 || J.color == 6) { s[1] = q[1]; s[2] = q[2]; }
    } else if (J.depth == 16) {
        const uint8_t* q = row + (size_t)px * J.channels * 2;        // png_set_strip_16: the high byte
        s[0] = q[0];
        if (J.color == 2 || J.color == 6) { s[1] = q[2]; s[2] = q[4]; }
    } else {
        const int bit = px * J.depth;
        const int v = (row[bit >> 3] >> (8 - J.depth - (bit & 7))) & ((1 << J.depth) - 1);
        s[0] = J.color == 3 ? v : v * (255 / ((1 << J.depth) - 1));  // png_set_expand_gray_1_2_4_to_8
    }
    uint8_t b, g, r_;
    if (J.color == 3) {
        if (s[0] >= J.npal) { atomicExch(error_flag, 2); s[0] = 0; }
        const uint8_t* e = J.pal + 3 * s[0];
        r_ = e[0]; g = e[1]; b = e[2];
    } else if (J.color == 2 || J.color == 6) {
        r_ = (uint8_t)s[0]; g = (uint8_t)s[1]; b = (uint8_t)s[2];
    } else {
        r_ = g = b = (uint8_t)s[0];
    }
    uint8_t* o = J.out + ((size_t)y * J.W + x) * 3;
    o[0] = b; o[1] = g; o[2] = r_;
}

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    void* get(size_t n) {
        if (n <= cap) return p;
        if (p) cudaFreeHost(p);
        cap = n + n / 4 + 4096;
        if (cudaMallocHost(&p, cap) != cudaSuccess) { p = nullptr; cap = 0; }
        return p;
    }
};

}  // namespace

bool png_is_png(const uint8_t* data, size_t nbytes) { return data != nullptr && nbytes >= 8 && memcmp(data, kSig, 8) == 0; }

int png_probe(const uint8_t* data, size_t nbytes, int* height, int* width, int* components, std::string* err) {
    Header hd;
    if (!parse(data, nbytes, &hd, err)) return LOCR_ERR_INVALID;
    *height = hd.H;
    *width = hd.W;
    *components = hd.channels;
    return LOCR_OK;
}

int png_host_scanlines(const uint8_t* data, size_t nbytes, uint8_t* out, size_t capacity, size_t* need,
                       std::string* err) {
    Header hd;
    if (!parse(data, nbytes, &hd, err)) return LOCR_ERR_INVALID;
    if (need) *need = hd.filtered_bytes;
    if (out == nullptr) return LOCR_OK;
    if (capacity < hd.filtered_bytes) { *err = "PNG: output capacity too small"; return LOCR_ERR_CAPACITY; }
    return inflate_idat(hd, out, false, err) ? LOCR_OK : LOCR_ERR_INVALID;
}

int png_decode_to_device(locr_handle* h, const uint8_t* const* blobs, const int64_t* nbytes, int n,
                         uint8_t* const* d_out) {
    static thread_local PinnedBuf host_buf;
    std::vector<Header> hd((size_t)n);
    std::vector<size_t> in_off((size_t)n), plain_off((size_t)n);
    size_t in_total = 0, plain_total = 0;
    std::string err;
    int njobs = 0;
    for (int i = 0; i < n; ++i) {
        if (blobs[i] == nullptr || nbytes[i] <= 0) return h->fail(LOCR_ERR_INVALID, "PNG: empty input");
        if (!parse(blobs[i], (size_t)nbytes[i], &hd[i], &err)) return h->fail(LOCR_ERR_INVALID, err);
        in_off[i] = in_total;
        plain_off[i] = plain_total;
        in_total += hd[i].padded_bytes;
        plain_total += (hd[i].plain_bytes + 15) / 16 * 16;
        for (int p = 0; p < hd[i].npass; ++p) njobs += hd[i].pass[p].h > 0;
    }
    const size_t meta_bytes = ((size_t)njobs * sizeof(UnfilterJob) + (size_t)n * sizeof(ColorJob) + (size_t)n * 768 + 63) / 64 * 64;
    uint8_t* hb = (uint8_t*)host_buf.get(in_total + meta_bytes);
    if (!hb) return h->fail(LOCR_ERR_CUDA, "pinned allocation for the PNG scanlines failed");
    // inflate: one host thread per image (a zlib stream is one serial bit stream)
    {
        unsigned hw = std::thread::hardware_concurrency();
        int nthreads = n < 8 ? n : 8;
        if (hw && (int)hw < nthreads) nthreads = (int)hw;
        std::atomic<int> next(0);
        std::vector<std::string> errs((size_t)n);
        std::vector<char> ok((size_t)n, 1);
        auto work = [&]() {
            for (;;) {
                const int i = next.fetch_add(1);
                if (i >= n) break;
                if (!inflate_idat(hd[i], hb + in_off[i], true, &errs[i])) ok[i] = 0;
            }
        };
        if (nthreads <= 1) {
            work();
        } else {
            std::vector<std::thread> pool;
            for (int t = 0; t < nthreads; ++t) pool.emplace_back(work);
            for (auto& t : pool) t.join();
        }
        for (int i = 0; i < n; ++i)
            if (!ok[i]) return h->fail(LOCR_ERR_INVALID, errs[i]);
    }
    cudaStream_t s = h->stream;
    uint8_t* d_in = (uint8_t*)engine_buffer(h, "png.in", in_total + meta_bytes);
    uint8_t* d_plain = (uint8_t*)engine_buffer(h, "png.plain", plain_total + 16);
    int* d_flag = (int*)engine_buffer(h, "png.flag", 16);
    if (!d_in || !d_plain || !d_flag) return h->fail(LOCR_ERR_CUDA, "allocation failed");
    // job tables and palettes ride behind the scanlines in the same pinned buffer / copy
    uint8_t* meta = hb + in_total;
    UnfilterJob* uj = reinterpret_cast<UnfilterJob*>(meta);
    ColorJob* cj = reinterpret_cast<ColorJob*>(meta + (size_t)njobs * sizeof(UnfilterJob));
    uint8_t* pals = reinterpret_cast<uint8_t*>(cj + n);
    uint8_t* d_meta = d_in + in_total;
    const uint8_t* d_pals = d_meta + (size_t)njobs * sizeof(UnfilterJob) + (size_t)n * sizeof(ColorJob);
    int j = 0, maxW = 0, maxH = 0;
    for (int i = 0; i < n; ++i) {
        const Header& H = hd[i];
        ColorJob& c = cj[i];
        memset(&c, 0, sizeof(c));
        c.plain = d_plain + plain_off[i];
        c.out = d_out[i];
        c.pal = H.color == 3 ? d_pals + (size_t)i * 768 : nullptr;
        c.npal = H.npal;
        c.W = H.W; c.H = H.H; c.depth = H.depth; c.color = H.color; c.channels = H.channels; c.interlace = H.interlace;
        memcpy(pals + (size_t)i * 768, H.pal, 768);
        for (int p = 0; p < H.npass; ++p) {
            const PassGeo& g = H.pass[p];
            c.pass_pitch[p] = (int)g.out_pitch;
            c.pass_off[p] = (long)g.out_off;
            if (g.h <= 0) continue;
            if (g.rowbytes > 0x7fffff00u) return h->fail(LOCR_ERR_CAPACITY, "PNG: scanline too long");
            uj[j].in = d_in + in_off[i] + g.in_off;
            uj[j].out = d_plain + plain_off[i] + g.out_off;
            uj[j].h = g.h;
            uj[j].rowbytes = (int)g.rowbytes;
            uj[j].in_pitch = (int)g.in_pitch;
            uj[j].out_pitch = (int)g.out_pitch;
            uj[j].bpp = H.bpp;
            ++j;
        }
        if (H.W > maxW) maxW = H.W;
        if (H.H > maxH) maxH = H.H;
    }
    LOCR_CUDA_OK(cudaMemcpyAsync(d_in, hb, in_total + meta_bytes, cudaMemcpyHostToDevice, s));
    LOCR_CUDA_OK(cudaMemsetAsync(d_flag, 0, 4, s));
    {
        ProfScope ps_(h, "png_unfilter", 0, false);
        png_unfilter_kernel<<<njobs, 1024, 0, s>>>(reinterpret_cast<const UnfilterJob*>(d_meta), d_flag);
    }
    {
        ProfScope ps_(h, "png_color", 0, false);
        dim3 blk(32, 8), grd((maxW + 31) / 32, (maxH + 7) / 8, n);
        png_color_kernel<<<grd, blk, 0, s>>>(
            reinterpret_cast<const ColorJob*>(d_meta + (size_t)njobs * sizeof(UnfilterJob)), d_flag);
    }
    h->launches += 2;
    LOCR_CUDA_OK(cudaGetLastError());
    int flag = 0;
    LOCR_CUDA_OK(cudaMemcpyAsync(&flag, d_flag, 4, cudaMemcpyDeviceToHost, s));
    // the pinned buffer is reused by the next call of this thread: wait until the copy has left it
    LOCR_CUDA_OK(cudaStreamSynchronize(s));
    if (in_total > ((size_t)1 << 30)) {
        host_buf.release();
        engine_release(h, "png.in");
        engine_release(h, "png.plain");
    }
    if (flag == 1) return h->fail(LOCR_ERR_INVALID, "PNG: bad filter type in the image data");
    if (flag == 2) return h->fail(LOCR_ERR_INVALID, "PNG: palette index out of range");
    return LOCR_OK;
}

}  // namespace locr
