// cv2.imread / cv2.imdecode(IMREAD_COLOR) for baseline JPEG files — the step in front of the detect-then-recognize path
// (reference ocr/pipeline.py:68).  OpenCV hands JPEG files to libjpeg(-turbo) with its defaults; the
// stages are restated from the published algorithms and are integer-exact:
//   host   : marker parsing, Huffman entropy decoding (ITU-T T.81 annex F: sequential scans; annex G: progressive scans
//            with spectral selection, successive approximation and EOB runs), one host thread per image -> quantised
//            coefficients, int16 [component][block row][block col][64] in pinned memory
//   device : jpeg_idct_kernel   dequantisation + jidctint.c "ISLOW" inverse DCT (CONST_BITS 13, PASS1_BITS 2)
//            jpeg_color_kernel  jdsample.c up-sampling (triangle filters for 2:1 ratios, replication otherwise) fused
//                               with jdcolor.c YCbCr -> RGB (16-bit fixed point), written as packed BGR
// EXIF orientations 2-8 are applied while the pixels are written, like cv2.imread does after decoding.
// The entropy decoder is inherently serial per restart interval, which is why it stays on host cores; everything that
// is data-parallel runs on the GPU and the decoded image never visits the host.
#include "jpeg.cuh"

#include <stdlib.h>

#include <atomic>
#include <thread>
#include <vector>

#include "engine.cuh"

namespace locr {

namespace {

constexpr int kLook = 10;   // Huffman look-ahead bits

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTab {
    bool present = false;
    uint8_t vals[256];
    uint16_t look[1 << kLook];   // (code length << 8) | symbol, 0 = longer than kLook bits
    // AC fast path: when a code AND its magnitude bits fit into the look-ahead window, the decoded coefficient itself:
    // (value << 16) | (zero run << 8) | (bits consumed), 0 = take the general path
    int32_t fast[1 << kLook];
    int32_t maxcode[18];
    int32_t valoffset[17];
};

struct Comp {
    int id, h, v, tq, td, ta;
    int blocks_h, blocks_v;      // padded to whole MCUs
    int dw, dh;                  // true down-sampled size (what the up-samplers see)
    size_t coef_off;             // int16 elements from the image's coefficient base
    size_t plane_off;            // bytes from the image's plane base
};

struct Header {
    int H = 0, W = 0, ncomp = 0;
    Comp comp[3];
    uint16_t qt[4][64];          // natural order
    bool qt_present[4] = {false, false, false, false};
    HuffTab dc[4], ac[4];
    int restart_interval = 0;
    const uint8_t* data = nullptr;   // the whole file
    size_t size = 0;
    const uint8_t* scan = nullptr;   // entropy-coded data of the first scan
    size_t scan_len = 0;
    int orientation = 1;
    bool progressive = false;    // SOF2
    bool fast = false;           // one interleaved sequential scan in frame order: decode_scan(); else decode_multiscan()
    size_t sos_pos = 0;          // offset of the first SOS marker (multi-scan files are walked again from here)
    int hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
    size_t coef_elems = 0, plane_bytes = 0;
};

// Returns false for a table whose code lengths over-subscribe the code space (a corrupt DHT segment).
bool build_table(const uint8_t* bits /* [16] */, const uint8_t* vals, int count, HuffTab* t) {
    {
        int32_t code = 0;
        for (int l = 1; l <= 16; ++l) {
            code += bits[l - 1];
            if (code > (1 << l)) return false;
            code <<= 1;
        }
    }
    t->present = true;
    memset(t->vals, 0, sizeof(t->vals));
    memcpy(t->vals, vals, (size_t)count);
    memset(t->look, 0, sizeof(t->look));
    int32_t code = 0;
    int p = 0;
    for (int l = 1; l <= 16; ++l) {
        const int nb = bits[l - 1];
        if (nb) {
            t->valoffset[l] = p - code;
            for (int i = 0; i < nb; ++i, ++p, ++code) {
                if (l <= kLook) {
                    const int lo = code << (kLook - l);
                    for (int k = 0; k < (1 << (kLook - l)); ++k) t->look[lo + k] = (uint16_t)((l << 8) | t->vals[p]);
                }
            }
            t->maxcode[l] = code - 1;
        } else {
            t->maxcode[l] = -1;
            t->valoffset[l] = 0;
        }
        code <<= 1;
    }
    t->maxcode[17] = 0x7fffffff;
    for (int i = 0; i < (1 << kLook); ++i) {
        t->fast[i] = 0;
        const uint16_t e = t->look[i];
        if (!e) continue;
        const int len = e >> 8, run = (e & 0xFF) >> 4, mag = e & 15;
        if (mag == 0 || len + mag > kLook) continue;
        const int v = (i >> (kLook - len - mag)) & ((1 << mag) - 1);
        const int val = v >= (1 << (mag - 1)) ? v : v - (1 << mag) + 1;
        t->fast[i] = (int32_t)(((uint32_t)(uint16_t)(int16_t)val << 16) | (uint32_t)(run << 8) | (uint32_t)(len + mag));
    }
    return true;
}

int exif_orientation(const uint8_t* t, size_t n) {
    if (n < 14) return 1;
    const bool le = t[0] == 'I';
    auto rd16 = [&](size_t o) -> uint32_t { return le ? (t[o] | (t[o + 1] << 8)) : ((t[o] << 8) | t[o + 1]); };
    auto rd32 = [&](size_t o) -> uint32_t {
        return le ? (t[o] | (t[o + 1] << 8) | (t[o + 2] << 16) | ((uint32_t)t[o + 3] << 24))
                  : (((uint32_t)t[o] << 24) | (t[o + 1] << 16) | (t[o + 2] << 8) | t[o + 3]);
    };
    const size_t off = rd32(4);
    if (off + 2 > n) return 1;
    const uint32_t cnt = rd16(off);
    for (uint32_t k = 0; k < cnt; ++k) {
        const size_t e = off + 2 + 12 * (size_t)k;
        if (e + 12 > n) break;
        if (rd16(e) == 0x0112) return (int)rd16(e + 8);
    }
    return 1;
}

size_t max_image_pixels() {
    static size_t v = 0;
    if (v == 0) {
        const char* e = getenv("LOCR_MAX_IMAGE_PIXELS");
        v = e ? (size_t)strtoull(e, nullptr, 10) : 0;
        if (v == 0) v = (size_t)1 << 30;     // cv2.imread refuses larger images (CV_IO_MAX_IMAGE_PIXELS)
    }
    return v;
}

// Marker segments up to the first SOS.  Returns false with a reason on anything this decoder does not cover.
bool parse(const uint8_t* d, size_t n, Header* hd, std::string* err) {
    auto bad = [&](const char* m) { *err = std::string("JPEG: ") + m; return false; };
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return bad("not a JPEG file (no SOI marker)");
    hd->data = d;
    hd->size = n;
    size_t pos = 2;
    struct Frame { int id, h, v, tq; } frame[3];
    int nframe = 0;
    bool have_frame = false;
    bool saw_jfif = false;
    int adobe_transform = -1;      // APP14 comes BEFORE the frame header: remembered here, judged at the first scan
    for (;;) {
        if (pos + 4 > n) return bad("truncated file");
        if (d[pos] != 0xFF) return bad("marker expected");
        while (pos + 1 < n && d[pos + 1] == 0xFF) ++pos;
        if (pos + 4 > n) return bad("truncated file");
        const int m = d[pos + 1];
        pos += 2;
        if (m == 0xD9) return bad("EOI before any scan");
        if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        const size_t len = ((size_t)d[pos] << 8) | d[pos + 1];
        if (len < 2 || pos + len > n) return bad("truncated segment");
        const uint8_t* s = d + pos + 2;
        const size_t sl = len - 2;
        pos += len;
        if (m == 0xDB) {
            size_t i = 0;
            while (i < sl) {
                const int pq = s[i] >> 4, tq = s[i] & 15;
                ++i;
                if (tq > 3 || i + (pq ? 128 : 64) > sl) return bad("bad DQT segment");
                for (int k = 0; k < 64; ++k) {
                    hd->qt[tq][kZigzag[k]] = pq ? (uint16_t)((s[i + 2 * k] << 8) | s[i + 2 * k + 1]) : s[i + k];
                }
                i += pq ? 128 : 64;
                hd->qt_present[tq] = true;
            }
        } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
            if (sl < 6 || s[0] != 8) return bad("only 8-bit samples are supported");
            hd->progressive = m == 0xC2;
            hd->H = (s[1] << 8) | s[2];
            hd->W = (s[3] << 8) | s[4];
            nframe = s[5];
            if (nframe != 1 && nframe != 3) return bad("only 1- or 3-component files are supported");
            if (sl < 6 + 3 * (size_t)nframe) return bad("bad SOF segment");
            for (int k = 0; k < nframe; ++k)
                frame[k] = {s[6 + 3 * k], s[7 + 3 * k] >> 4, s[7 + 3 * k] & 15, s[8 + 3 * k]};
            have_frame = true;
        } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return bad("unsupported JPEG process (lossless / hierarchical / arithmetic-coded): only Huffman-coded DCT files");
        } else if (m == 0xC4) {
            size_t i = 0;
            while (i < sl) {
                if (i + 17 > sl) return bad("bad DHT segment");
                const int tc = s[i] >> 4, th = s[i] & 15;
                int cnt = 0;
                for (int k = 0; k < 16; ++k) cnt += s[i + 1 + k];
                if (tc > 1 || th > 3 || cnt > 256 || i + 17 + cnt > sl) return bad("bad DHT segment");
                if (!build_table(s + i + 1, s + i + 17, cnt, tc ? &hd->ac[th] : &hd->dc[th])) return bad("bad Huffman table");
                i += 17 + (size_t)cnt;
            }
        } else if (m == 0xDD) {
            if (sl < 2) return bad("bad DRI segment");
            hd->restart_interval = (s[0] << 8) | s[1];
        } else if (m == 0xE1 && sl > 6 && memcmp(s, "Exif\0\0", 6) == 0) {
            hd->orientation = exif_orientation(s + 6, sl - 6);
        } else if (m == 0xE0 && sl >= 5 && memcmp(s, "JFIF\0", 5) == 0) {
            saw_jfif = true;
        } else if (m == 0xEE && sl >= 12 && memcmp(s, "Adobe", 5) == 0) {
            adobe_transform = s[11];
        } else if (m == 0xDA) {
            if (!have_frame) return bad("SOS before SOF");
            if (nframe == 3 && !saw_jfif) {
                // libjpeg's colour-space guess (jdapimin.c default_decompress_parms): JFIF implies YCbCr; otherwise an
                // Adobe marker with transform 0, or component ids 'R','G','B' without any marker, mean untransformed RGB
                const bool rgb_ids = frame[0].id == 'R' && frame[1].id == 'G' && frame[2].id == 'B';
                if (adobe_transform == 0 || (adobe_transform < 0 && rgb_ids))
                    return bad("untransformed RGB files (Adobe transform 0 / component ids R,G,B) are not supported");
            }
            const int ns = sl >= 1 ? s[0] : 0;
            if (ns < 1 || ns > nframe || sl < 4 + 2 * (size_t)ns) return bad("bad SOS segment");
            hd->ncomp = nframe;
            for (int k = 0; k < nframe; ++k) {
                Comp& c = hd->comp[k];
                c.id = frame[k].id; c.h = frame[k].h; c.v = frame[k].v; c.tq = frame[k].tq; c.td = c.ta = 0;
                if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return bad("bad component");
                if (!hd->qt_present[c.tq]) return bad("missing quantisation table");
            }
            // the common file: ONE interleaved sequential scan with the components in frame order
            hd->fast = !hd->progressive && ns == nframe;
            for (int k = 0; k < ns && hd->fast; ++k) {
                const int cid = s[1 + 2 * k], tabs = s[2 + 2 * k];
                Comp& c = hd->comp[k];
                if (cid != c.id) { hd->fast = false; break; }
                c.td = tabs >> 4; c.ta = tabs & 15;
                if (c.td > 3 || c.ta > 3) return bad("bad component");
                if (!hd->dc[c.td].present || !hd->ac[c.ta].present) return bad("missing table");
            }
            hd->sos_pos = pos - len - 2;
            hd->scan = d + pos;
            hd->scan_len = n - pos;
            break;
        }
    }
    if (hd->H <= 0 || hd->W <= 0) return bad("empty image");
    if ((size_t)hd->H * (size_t)hd->W > max_image_pixels())
        return bad("image larger than the pixel limit (OpenCV's CV_IO_MAX_IMAGE_PIXELS, 2^30; LOCR_MAX_IMAGE_PIXELS)");
    if (hd->orientation < 1 || hd->orientation > 8) hd->orientation = 1;   // OpenCV leaves other values alone
    hd->hmax = hd->vmax = 1;
    for (int k = 0; k < hd->ncomp; ++k) {
        if (hd->comp[k].h > hd->hmax) hd->hmax = hd->comp[k].h;
        if (hd->comp[k].v > hd->vmax) hd->vmax = hd->comp[k].v;
    }
    if (hd->ncomp == 1) { hd->comp[0].h = hd->comp[0].v = 1; hd->hmax = hd->vmax = 1; }   // T.81 A.2.2: single-component scans are not interleaved
    hd->mcux = (hd->W + 8 * hd->hmax - 1) / (8 * hd->hmax);
    hd->mcuy = (hd->H + 8 * hd->vmax - 1) / (8 * hd->vmax);
    size_t co = 0, po = 0;
    for (int k = 0; k < hd->ncomp; ++k) {
        Comp& c = hd->comp[k];
        if (hd->hmax % c.h || hd->vmax % c.v) return bad("fractional sampling ratios are not supported");
        c.blocks_h = hd->mcux * c.h;
        c.blocks_v = hd->mcuy * c.v;
        c.dw = (hd->W * c.h + hd->hmax - 1) / hd->hmax;
        c.dh = (hd->H * c.v + hd->vmax - 1) / hd->vmax;
        c.coef_off = co;
        c.plane_off = po;
        co += (size_t)c.blocks_h * c.blocks_v * 64;
        po += (size_t)c.blocks_h * c.blocks_v * 64;
    }
    hd->coef_elems = co;
    hd->plane_bytes = po;
    // Every coded block costs at least one bit in every scan that visits it, so a file whose entropy-coded part holds
    // fewer bits than it declares blocks is truncated: refuse it HERE, before any buffer is sized from the header
    // (a 200-byte file may declare 65535 x 65535 pixels).
    if ((size_t)hd->scan_len * 8 < co / 64) return bad("premature end of the entropy-coded data");
    return true;
}

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc = 0;
    int n = 0;
    int pad = 0;         // zero bits appended behind the data (a marker or the end of the file was reached)
    bool marker = false;
    inline void fill() {
        // fast path: four bytes at a time while none of them is 0xFF (no stuffing, no marker) - the common case
        while (n <= 32 && !marker && p + 4 <= end) {
            const uint32_t v = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
            const uint32_t inv = ~v;
            if ((inv - 0x01010101u) & ~inv & 0x80808080u) break;    // some byte of v is 0xFF
            acc = (acc << 32) | v;
            n += 32;
            p += 4;
        }
        while (n <= 56) {
            uint32_t b = 0;
            if (!marker && p < end) {
                b = *p;
                if (b == 0xFF) {
                    const uint32_t nb = (p + 1 < end) ? p[1] : 0xD9;
                    if (nb == 0) p += 2;                       // stuffed zero byte
                    else if (nb == 0xFF) { ++p; continue; }    // fill byte
                    else { marker = true; b = 0; pad += 8; }   // a real marker: feed zeros from here on
                } else {
                    ++p;
                }
            } else {
                pad += 8;
            }
            acc = (acc << 8) | b;
            n += 8;
        }
    }
    inline uint32_t peek(int k) const { return (uint32_t)(acc >> (n - k)) & ((1u << k) - 1u); }
    inline void skip(int k) { n -= k; }
    // true once a code or its extra bits reached into the padding: the file is truncated.  libjpeg carries on with a
    // warning and a partly grey image; this reader refuses instead of returning pixels it cannot vouch for.
    inline bool exhausted() const { return n < pad; }
};

inline int decode_symbol(BitReader& br, const HuffTab& t) {
    const uint16_t e = t.look[br.peek(kLook)];
    if (e) {
        br.skip(e >> 8);
        return e & 0xFF;
    }
    int l = kLook + 1;
    int32_t code = (int32_t)br.peek(l);
    while (l <= 16 && code > t.maxcode[l]) {
        ++l;
        code = (int32_t)br.peek(l);
    }
    if (l > 16) return -1;
    br.skip(l);
    return t.vals[(code + t.valoffset[l]) & 255];
}

// Entropy decoding of the whole scan into coef (int16, natural order, [component][block row][block col][64]).
bool decode_scan(const Header& hd, int16_t* coef, std::string* err) {
    BitReader br;
    br.p = hd.scan;
    br.end = hd.scan + hd.scan_len;
    int pred[3] = {0, 0, 0};
    const int total = hd.mcux * hd.mcuy;
    int until_restart = hd.restart_interval;
    for (int mcu = 0; mcu < total; ++mcu) {
        if (hd.restart_interval && until_restart == 0) {
            br.acc = 0; br.n = 0; br.pad = 0;                   // drop the padding bits
            const uint8_t* q = br.p;
            while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) ++q;
            if (q + 1 >= br.end) { *err = "JPEG: restart marker missing"; return false; }
            br.p = q + 2;
            br.marker = false;
            pred[0] = pred[1] = pred[2] = 0;
            until_restart = hd.restart_interval;
        }
        --until_restart;
        const int my = mcu / hd.mcux, mx = mcu - my * hd.mcux;
        for (int ci = 0; ci < hd.ncomp; ++ci) {
            const Comp& c = hd.comp[ci];
            const HuffTab& dct = hd.dc[c.td];
            const HuffTab& act = hd.ac[c.ta];
            for (int by = 0; by < c.v; ++by) {
                for (int bx = 0; bx < c.h; ++bx) {
                    int16_t* blk = coef + c.coef_off + ((size_t)(my * c.v + by) * c.blocks_h + (mx * c.h + bx)) * 64;
                    memset(blk, 0, 128);
                    br.fill();
                    int s = decode_symbol(br, dct);
                    if (s < 0 || s > 15) { *err = "JPEG: corrupt entropy-coded data"; return false; }
                    if (s) {
                        const int v = (int)br.peek(s);
                        br.skip(s);
                        pred[ci] += v >= (1 << (s - 1)) ? v : v - (1 << s) + 1;
                    }
                    blk[0] = (int16_t)pred[ci];
                    for (int k = 1; k < 64;) {
                        if (br.n < 32) br.fill();
                        const int32_t f = act.fast[br.peek(kLook)];
                        if (f) {                       // code + magnitude inside the look-ahead window
                            k += (f >> 8) & 15;
                            if (k > 63) { *err = "JPEG: corrupt entropy-coded data"; return false; }
                            blk[kZigzag[k]] = (int16_t)(f >> 16);
                            br.skip(f & 255);
                            ++k;
                            continue;
                        }
                        const int rs = decode_symbol(br, act);
                        if (rs < 0) { *err = "JPEG: corrupt entropy-coded data"; return false; }
                        const int r = rs >> 4;
                        s = rs & 15;
                        if (s == 0) {
                            if (r == 15) { k += 16; continue; }
                            break;
                        }
                        k += r;
                        if (k > 63) { *err = "JPEG: corrupt entropy-coded data"; return false; }
                        const int v = (int)br.peek(s);
                        br.skip(s);
                        blk[kZigzag[k]] = (int16_t)(v >= (1 << (s - 1)) ? v : v - (1 << s) + 1);
                        ++k;
                    }
                    if (br.exhausted()) { *err = "JPEG: premature end of the entropy-coded data"; return false; }
                }
            }
        }
    }
    return true;
}

// One block of a progressive or non-interleaved scan (T.81 annex G; the refinement pass follows jdphuff.c's
// decode_mcu_AC_refine flow).  ss / se = spectral selection, ah / al = successive approximation.
inline bool decode_block_general(BitReader& br, int16_t* blk, const HuffTab* dct, const HuffTab* act, int ss, int se,
                                 int ah, int al, bool progressive, int* pred, int* eobrun) {
    br.fill();
    if (ss == 0) {
        if (ah == 0) {
            const int s = decode_symbol(br, *dct);
            if (s < 0 || s > 15) return false;
            if (s) {
                const int v = (int)br.peek(s);
                br.skip(s);
                *pred += v >= (1 << (s - 1)) ? v : v - (1 << s) + 1;
            }
            blk[0] = (int16_t)(*pred * (1 << al));
        } else {
            if (br.peek(1)) blk[0] = (int16_t)(blk[0] | (1 << al));
            br.skip(1);
        }
        if (se == 0) return true;
    }
    int k = ss > 1 ? ss : 1;
    if (ah == 0) {
        if (*eobrun > 0) { --*eobrun; return true; }
        while (k <= se) {
            if (br.n < 32) br.fill();
            const int rs = decode_symbol(br, *act);
            if (rs < 0) return false;
            const int r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r == 15) { k += 16; continue; }
                if (progressive) {
                    *eobrun = (1 << r) - 1;
                    if (r) { *eobrun += (int)br.peek(r); br.skip(r); }
                }
                break;
            }
            k += r;
            if (k > 63) return false;
            const int v = (int)br.peek(s);
            br.skip(s);
            blk[kZigzag[k]] = (int16_t)((v >= (1 << (s - 1)) ? v : v - (1 << s) + 1) * (1 << al));
            ++k;
        }
        return true;
    }
    const int p1 = 1 << al, m1 = -(1 << al);
    if (*eobrun == 0) {
        while (k <= se) {
            if (br.n < 32) br.fill();
            const int rs = decode_symbol(br, *act);
            if (rs < 0) return false;
            int r = rs >> 4;
            const int s = rs & 15;
            int val = 0;
            if (s) {
                val = br.peek(1) ? p1 : m1;
                br.skip(1);
            } else if (r != 15) {
                *eobrun = 1 << r;
                if (r) { *eobrun += (int)br.peek(r); br.skip(r); }
                break;
            }
            while (k <= se) {
                int16_t* c = blk + kZigzag[k];
                if (*c != 0) {
                    if (br.n < 8) br.fill();
                    if (br.peek(1) && (*c & p1) == 0) *c = (int16_t)(*c + (*c >= 0 ? p1 : m1));
                    br.skip(1);
                } else if (--r < 0) {
                    break;
                }
                ++k;
            }
            if (val && k <= 63) blk[kZigzag[k]] = (int16_t)val;
            ++k;
        }
    }
    if (*eobrun > 0) {
        while (k <= se) {
            int16_t* c = blk + kZigzag[k];
            if (*c != 0) {
                if (br.n < 8) br.fill();
                if (br.peek(1) && (*c & p1) == 0) *c = (int16_t)(*c + (*c >= 0 ? p1 : m1));
                br.skip(1);
            }
            ++k;
        }
        --*eobrun;
    }
    return true;
}

// Progressive files and sequential files with more than one scan: walks the marker segments from the first SOS on
// (Huffman tables and the restart interval may change between scans) and decodes every scan into the coefficient planes.
bool decode_multiscan(const Header& hd, int16_t* coef, std::string* err) {
    auto bad = [&](const char* m) { *err = std::string("JPEG: ") + m; return false; };
    std::vector<HuffTab> dc(hd.dc, hd.dc + 4), ac(hd.ac, hd.ac + 4);
    int ri = hd.restart_interval;
    memset(coef, 0, hd.coef_elems * 2);
    const uint8_t* d = hd.data;
    const size_t n = hd.size;
    size_t pos = hd.sos_pos;
    int scans = 0;
    while (pos + 4 <= n) {
        if (d[pos] != 0xFF) return bad("marker expected between scans");
        while (pos + 1 < n && d[pos + 1] == 0xFF) ++pos;
        if (pos + 2 > n) break;
        const int m = d[pos + 1];
        pos += 2;
        if (m == 0xD9) break;
        if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (pos + 2 > n) return bad("truncated file");
        const size_t len = ((size_t)d[pos] << 8) | d[pos + 1];
        if (len < 2 || pos + len > n) return bad("truncated segment");
        const uint8_t* s = d + pos + 2;
        const size_t sl = len - 2;
        pos += len;
        if (m == 0xC4) {
            size_t i = 0;
            while (i < sl) {
                if (i + 17 > sl) return bad("bad DHT segment");
                const int tc = s[i] >> 4, th = s[i] & 15;
                int cnt = 0;
                for (int k = 0; k < 16; ++k) cnt += s[i + 1 + k];
                if (tc > 1 || th > 3 || cnt > 256 || i + 17 + cnt > sl) return bad("bad DHT segment");
                if (!build_table(s + i + 1, s + i + 17, cnt, tc ? &ac[th] : &dc[th])) return bad("bad Huffman table");
                i += 17 + (size_t)cnt;
            }
        } else if (m == 0xDD) {
            if (sl < 2) return bad("bad DRI segment");
            ri = (s[0] << 8) | s[1];
        } else if (m == 0xDB) {
            return bad("quantisation tables redefined between scans are not supported");
        } else if (m == 0xDA) {
            const int ns = sl >= 1 ? s[0] : 0;
            if (ns < 1 || ns > hd.ncomp || sl < 4 + 2 * (size_t)ns) return bad("bad SOS segment");
            int ci[3], td[3], ta[3];
            for (int k = 0; k < ns; ++k) {
                ci[k] = -1;
                for (int j = 0; j < hd.ncomp; ++j)
                    if (hd.comp[j].id == s[1 + 2 * k]) ci[k] = j;
                td[k] = s[2 + 2 * k] >> 4;
                ta[k] = s[2 + 2 * k] & 15;
                if (ci[k] < 0 || td[k] > 3 || ta[k] > 3) return bad("bad scan component");
            }
            const int ss = s[1 + 2 * ns], se = s[2 + 2 * ns], ah = s[3 + 2 * ns] >> 4, al = s[3 + 2 * ns] & 15;
            if (ss > se || se > 63 || al > 13 || ah > 13) return bad("bad scan parameters");
            if (!hd.progressive && (ss != 0 || se != 63 || ah || al)) return bad("bad sequential scan header");
            if (hd.progressive && ss == 0 && se != 0) return bad("bad progressive scan header");
            if (ss > 0 && ns != 1) return bad("AC scans must hold one component");
            for (int k = 0; k < ns; ++k) {
                if (ss == 0 && ah == 0 && !dc[td[k]].present) return bad("missing DC table");
                if (se > 0 && !ac[ta[k]].present) return bad("missing AC table");
            }
            BitReader br;
            br.p = d + pos;
            br.end = d + n;
            int pred[3] = {0, 0, 0};
            int eobrun = 0;
            // units of the scan: whole MCUs when interleaved, else the blocks of the component's true extent
            const Comp& c0 = hd.comp[ci[0]];
            const int ux = ns > 1 ? hd.mcux : (c0.dw + 7) / 8;
            const int uy = ns > 1 ? hd.mcuy : (c0.dh + 7) / 8;
            int until_restart = ri;
            for (int u = 0; u < ux * uy; ++u) {
                if (ri && until_restart == 0) {
                    br.acc = 0; br.n = 0; br.pad = 0;
                    const uint8_t* q = br.p;
                    while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) ++q;
                    if (q + 1 >= br.end) return bad("restart marker missing");
                    br.p = q + 2;
                    br.marker = false;
                    pred[0] = pred[1] = pred[2] = 0;
                    eobrun = 0;
                    until_restart = ri;
                }
                --until_restart;
                const int uy_i = u / ux, ux_i = u - uy_i * ux;
                for (int k = 0; k < ns; ++k) {
                    const Comp& c = hd.comp[ci[k]];
                    const int nh = ns > 1 ? c.h : 1, nv = ns > 1 ? c.v : 1;
                    for (int by = 0; by < nv; ++by)
                        for (int bx = 0; bx < nh; ++bx) {
                            int16_t* blk = coef + c.coef_off +
                                           ((size_t)(uy_i * nv + by) * c.blocks_h + (ux_i * nh + bx)) * 64;
                            if (!decode_block_general(br, blk, &dc[td[k]], &ac[ta[k]], ss, se, ah, al, hd.progressive,
                                                      &pred[ci[k]], &eobrun))
                                return bad("corrupt entropy-coded data");
                            if (br.exhausted()) return bad("premature end of the entropy-coded data");
                        }
                }
            }
            // next marker: the first 0xFF that is followed by neither a stuffed zero, a restart marker nor a fill byte
            const uint8_t* q = br.p;
            while (q + 1 < d + n && !(q[0] == 0xFF && q[1] != 0x00 && q[1] != 0xFF && !(q[1] >= 0xD0 && q[1] <= 0xD7))) ++q;
            pos = (size_t)(q - d);
            ++scans;
        }
    }
    if (scans == 0) return bad("no scan");
    return true;
}

bool decode_image(const Header& hd, int16_t* coef, std::string* err) {
    return hd.fast ? decode_scan(hd, coef, err) : decode_multiscan(hd, coef, err);
}

// ---------------------------------------------------------------------------------------------- device kernels

struct IdctParams {
    const int16_t* coef[3];
    uint8_t* plane[3];
    int nblocks[3], blocks_h[3];
    int ncomp;
    uint16_t q[3][64];
};

// jidctint.c jpeg_idct_islow, one pass over eight values; DESCALE(x, shift) = (x + (1 << (shift - 1))) >> shift
template <int SHIFT>
__device__ __forceinline__ void idct8(int x0, int x1, int x2, int x3, int x4, int x5, int x6, int x7, int* o) {
    int z1 = (x2 + x6) * 4433;
    const int tmp2 = z1 + x6 * (-15137);
    const int tmp3 = z1 + x2 * 6270;
    const int tmp0 = (x0 + x4) << 13;
    const int tmp1 = (x0 - x4) << 13;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    int t0 = x7, t1 = x5, t2 = x3, t3 = x1;
    z1 = t0 + t3;
    int z2 = t1 + t2, z3 = t0 + t2, z4 = t1 + t3;
    const int z5 = (z3 + z4) * 9633;
    t0 *= 2446; t1 *= 16819; t2 *= 25172; t3 *= 12299;
    z1 *= -7373; z2 *= -20995;
    z3 = z3 * (-16069) + z5;
    z4 = z4 * (-3196) + z5;
    t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
    constexpr int R = 1 << (SHIFT - 1);
    o[0] = (tmp10 + t3 + R) >> SHIFT; o[7] = (tmp10 - t3 + R) >> SHIFT;
    o[1] = (tmp11 + t2 + R) >> SHIFT; o[6] = (tmp11 - t2 + R) >> SHIFT;
    o[2] = (tmp12 + t1 + R) >> SHIFT; o[5] = (tmp12 - t1 + R) >> SHIFT;
    o[3] = (tmp13 + t0 + R) >> SHIFT; o[4] = (tmp13 - t0 + R) >> SHIFT;
}

__device__ __forceinline__ uint32_t range_limit(int v) {
    v &= 1023;                       // range_limit[x & RANGE_MASK]: a 10-bit signed value, re-centred and clamped
    if (v >= 512) v -= 1024;
    v += 128;
    return (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// One thread = one 8x8 block: 64 dequantised coefficients -> 64 samples of the component plane
// (plane row pitch = blocks_h * 8 bytes).
__global__ void jpeg_idct_kernel(const IdctParams p) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0;
    while (c < p.ncomp && b >= p.nblocks[c]) { b -= p.nblocks[c]; ++c; }
    if (c >= p.ncomp) return;
    const uint4* src = reinterpret_cast<const uint4*>(p.coef[c] + (size_t)b * 64);
    const uint16_t* q = p.q[c];
    int ws[64];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint4 u = __ldg(src + r);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ws[r * 8 + 2 * k] = (int)(int16_t)(w[k] & 0xFFFF) * (int)q[r * 8 + 2 * k];
            ws[r * 8 + 2 * k + 1] = (int)(int16_t)(w[k] >> 16) * (int)q[r * 8 + 2 * k + 1];
        }
    }
    // pass 1: columns, results scaled up by 2^PASS1_BITS
#pragma unroll
    for (int col = 0; col < 8; ++col) {
        int o[8];
        idct8<11>(ws[col], ws[8 + col], ws[16 + col], ws[24 + col], ws[32 + col], ws[40 + col], ws[48 + col], ws[56 + col], o);
#pragma unroll
        for (int r = 0; r < 8; ++r) ws[r * 8 + col] = o[r];
    }
    // pass 2: rows, descale by CONST_BITS + PASS1_BITS + 3 and range-limit
    const int bh = p.blocks_h[c];
    const int by = b / bh, bx = b - by * bh;
    uint8_t* dst = p.plane[c] + ((size_t)by * 8) * ((size_t)bh * 8) + (size_t)bx * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int o[8];
        idct8<18>(ws[r * 8], ws[r * 8 + 1], ws[r * 8 + 2], ws[r * 8 + 3], ws[r * 8 + 4], ws[r * 8 + 5], ws[r * 8 + 6], ws[r * 8 + 7], o);
        uint2 v;
        v.x = range_limit(o[0]) | (range_limit(o[1]) << 8) | (range_limit(o[2]) << 16) | (range_limit(o[3]) << 24);
        v.y = range_limit(o[4]) | (range_limit(o[5]) << 8) | (range_limit(o[6]) << 16) | (range_limit(o[7]) << 24);
        *reinterpret_cast<uint2*>(dst + (size_t)r * bh * 8) = v;
    }
}

enum { UP_NONE = 0, UP_H2V1 = 1, UP_H2V2 = 2, UP_H1V2 = 3, UP_REPL = 4 };

struct ColorParams {
    const uint8_t* plane[3];
    int pitch[3], dw[3], dh[3], hs[3], vs[3], mode[3];
    int ncomp, H, W;
    int orientation;   // EXIF orientation 1..8 (OpenCV's ExifTransform applied while writing)
    uint8_t* out;      // packed BGR, [H][W][3] for orientations 1-4, [W][H][3] for 5-8
};

// jdsample.c: h2v1_fancy_upsample / h2v2_fancy_upsample / h1v2_fancy_upsample / int_upsample, evaluated per output sample
__device__ __forceinline__ int upsampled(const uint8_t* __restrict__ p, int pitch, int dw, int dh, int mode, int hs, int vs,
                                         int y, int x) {
    if (mode == UP_NONE) return p[(size_t)y * pitch + x];
    if (mode == UP_H2V1) {
        const int i = x >> 1;
        const uint8_t* row = p + (size_t)y * pitch;
        const int v = row[i];
        if (x & 1) return i == dw - 1 ? v : (3 * v + row[i + 1] + 2) >> 2;
        return i == 0 ? v : (3 * v + row[i - 1] + 1) >> 2;
    }
    if (mode == UP_H2V2) {
        const int r = y >> 1, i = x >> 1;
        int o = (y & 1) ? r + 1 : r - 1;     // the nearer neighbouring row; the image edge replicates its last row
        o = o < 0 ? 0 : (o > dh - 1 ? dh - 1 : o);
        const uint8_t* r0 = p + (size_t)r * pitch;
        const uint8_t* r1 = p + (size_t)o * pitch;
        const int cs = 3 * r0[i] + r1[i];
        if (x & 1) {
            if (i == dw - 1) return (4 * cs + 7) >> 4;
            return (3 * cs + 3 * r0[i + 1] + r1[i + 1] + 7) >> 4;
        }
        if (i == 0) return (4 * cs + 8) >> 4;
        return (3 * cs + 3 * r0[i - 1] + r1[i - 1] + 8) >> 4;
    }
    if (mode == UP_H1V2) {
        const int r = y >> 1;
        int o = (y & 1) ? r + 1 : r - 1;
        o = o < 0 ? 0 : (o > dh - 1 ? dh - 1 : o);
        const int a = p[(size_t)r * pitch + x], b = p[(size_t)o * pitch + x];
        return (3 * a + b + ((y & 1) ? 2 : 1)) >> 2;
    }
    return p[(size_t)(y / vs) * pitch + x / hs];
}

__device__ __forceinline__ int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// One thread = one output pixel.  jdcolor.c ycc_rgb_convert with SCALEBITS = 16:
// R = y + ((FIX(1.40200) * cr + HALF) >> 16), B = y + ((FIX(1.77200) * cb + HALF) >> 16),
// G = y + ((-FIX(0.34414) * cb + HALF - FIX(0.71414) * cr) >> 16), cb / cr centred on 128.
__global__ void jpeg_color_kernel(const ColorParams p) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= p.W || y >= p.H) return;
    const int Y = upsampled(p.plane[0], p.pitch[0], p.dw[0], p.dh[0], p.mode[0], p.hs[0], p.vs[0], y, x);
    // cv2.imread applies the EXIF orientation (loadsave.cpp ExifTransform: flips for 2-4, a transpose followed by a flip
    // for 5-8): source pixel (y, x) lands at (yd, xd) of the rotated image
    int yd = y, xd = x, wd = p.W;
    switch (p.orientation) {
        case 2: xd = p.W - 1 - x; break;
        case 3: yd = p.H - 1 - y; xd = p.W - 1 - x; break;
        case 4: yd = p.H - 1 - y; break;
        case 5: yd = x; xd = y; wd = p.H; break;
        case 6: yd = x; xd = p.H - 1 - y; wd = p.H; break;
        case 7: yd = p.W - 1 - x; xd = p.H - 1 - y; wd = p.H; break;
        case 8: yd = p.W - 1 - x; xd = y; wd = p.H; break;
        default: break;
    }
    uint8_t* o = p.out + ((size_t)yd * wd + xd) * 3;
    if (p.ncomp == 1) {
        o[0] = o[1] = o[2] = (uint8_t)Y;
        return;
    }
    const int cb = upsampled(p.plane[1], p.pitch[1], p.dw[1], p.dh[1], p.mode[1], p.hs[1], p.vs[1], y, x) - 128;
    const int cr = upsampled(p.plane[2], p.pitch[2], p.dw[2], p.dh[2], p.mode[2], p.hs[2], p.vs[2], y, x) - 128;
    o[0] = (uint8_t)clamp255(Y + ((116130 * cb + 32768) >> 16));
    o[1] = (uint8_t)clamp255(Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
    o[2] = (uint8_t)clamp255(Y + ((91881 * cr + 32768) >> 16));
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

int jpeg_probe(const uint8_t* data, size_t nbytes, int* height, int* width, int* components, std::string* err) {
    std::vector<Header> hd(1);
    if (!parse(data, nbytes, &hd[0], err)) return LOCR_ERR_INVALID;
    const bool swap = hd[0].orientation >= 5;     // size of the image as cv2.imread returns it
    *height = swap ? hd[0].W : hd[0].H;
    *width = swap ? hd[0].H : hd[0].W;
    *components = hd[0].ncomp;
    return LOCR_OK;
}

int jpeg_host_coefficients(const uint8_t* data, size_t nbytes, int16_t* out, size_t capacity, int* info,
                           std::string* err) {
    std::vector<Header> hd(1);
    if (!parse(data, nbytes, &hd[0], err)) return LOCR_ERR_INVALID;
    const Header& H = hd[0];
    info[0] = H.H; info[1] = H.W; info[2] = H.ncomp; info[3] = H.hmax; info[4] = H.vmax; info[5] = H.mcux; info[6] = H.mcuy;
    for (int k = 0; k < 3; ++k) {
        info[7 + 4 * k] = k < H.ncomp ? H.comp[k].h : 0;
        info[8 + 4 * k] = k < H.ncomp ? H.comp[k].v : 0;
        info[9 + 4 * k] = k < H.ncomp ? H.comp[k].blocks_h : 0;
        info[10 + 4 * k] = k < H.ncomp ? H.comp[k].blocks_v : 0;
    }
    if (out == nullptr) return LOCR_OK;
    if (capacity < H.coef_elems) { *err = "JPEG: coefficient buffer too small"; return LOCR_ERR_CAPACITY; }
    if (!decode_image(H, out, err)) return LOCR_ERR_INVALID;
    return LOCR_OK;
}

int jpeg_decode_to_device(locr_handle* h, const uint8_t* const* blobs, const int64_t* nbytes, int n,
                          uint8_t* const* d_out) {
    static thread_local PinnedBuf host_coef;
    std::vector<Header> hd((size_t)n);
    std::vector<size_t> coef_off((size_t)n), plane_off((size_t)n);
    size_t coef_total = 0, plane_total = 0;
    std::string err;
    for (int i = 0; i < n; ++i) {
        if (blobs[i] == nullptr || nbytes[i] <= 0) return h->fail(LOCR_ERR_INVALID, "JPEG: empty input");
        if (!parse(blobs[i], (size_t)nbytes[i], &hd[i], &err)) return h->fail(LOCR_ERR_INVALID, err);
        coef_off[i] = coef_total;
        plane_off[i] = plane_total;
        coef_total += hd[i].coef_elems;
        plane_total += (hd[i].plane_bytes + 255) / 256 * 256;
    }
    int16_t* hc = (int16_t*)host_coef.get(coef_total * 2);
    if (!hc) return h->fail(LOCR_ERR_CUDA, "pinned allocation for the JPEG coefficients failed");
    // entropy decoding: one host thread per image (a baseline scan without restart markers is one serial bit stream)
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
                if (!decode_image(hd[i], hc + coef_off[i], &errs[i])) ok[i] = 0;
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
    int16_t* dc = (int16_t*)engine_buffer(h, "jpeg.coef", coef_total * 2);
    uint8_t* dp = (uint8_t*)engine_buffer(h, "jpeg.planes", plane_total);
    if (!dc || !dp) return h->fail(LOCR_ERR_CUDA, "allocation failed");
    LOCR_CUDA_OK(cudaMemcpyAsync(dc, hc, coef_total * 2, cudaMemcpyHostToDevice, s));
    for (int i = 0; i < n; ++i) {
        const Header& H = hd[i];
        IdctParams ip;
        ColorParams cp;
        memset(&ip, 0, sizeof(ip));
        memset(&cp, 0, sizeof(cp));
        ip.ncomp = cp.ncomp = H.ncomp;
        int blocks = 0;
        for (int k = 0; k < H.ncomp; ++k) {
            const Comp& c = H.comp[k];
            ip.coef[k] = dc + coef_off[i] + c.coef_off;
            ip.plane[k] = dp + plane_off[i] + c.plane_off;
            ip.nblocks[k] = c.blocks_h * c.blocks_v;
            ip.blocks_h[k] = c.blocks_h;
            memcpy(ip.q[k], H.qt[c.tq], 128);
            blocks += ip.nblocks[k];
            cp.plane[k] = ip.plane[k];
            cp.pitch[k] = c.blocks_h * 8;
            cp.dw[k] = c.dw;
            cp.dh[k] = c.dh;
            cp.hs[k] = H.hmax / c.h;
            cp.vs[k] = H.vmax / c.v;
            // jdsample.c jinit_upsampler: triangle filters for the 2:1 ratios when the component is more than two
            // samples wide, replication for every other integral ratio
            if (cp.hs[k] == 1 && cp.vs[k] == 1) cp.mode[k] = UP_NONE;
            else if (cp.hs[k] == 2 && cp.vs[k] == 1 && c.dw > 2) cp.mode[k] = UP_H2V1;
            else if (cp.hs[k] == 2 && cp.vs[k] == 2 && c.dw > 2) cp.mode[k] = UP_H2V2;
            else if (cp.hs[k] == 1 && cp.vs[k] == 2) cp.mode[k] = UP_H1V2;
            else cp.mode[k] = UP_REPL;
        }
        cp.H = H.H;
        cp.W = H.W;
        cp.orientation = H.orientation;
        cp.out = d_out[i];
        {
            ProfScope ps_(h, "jpeg_idct", 0, false);
            jpeg_idct_kernel<<<(blocks + 127) / 128, 128, 0, s>>>(ip);
        }
        {
            ProfScope ps_(h, "jpeg_color", 0, false);
            dim3 blk(32, 8), grd((H.W + 31) / 32, (H.H + 7) / 8);
            jpeg_color_kernel<<<grd, blk, 0, s>>>(cp);
        }
        h->launches += 2;
    }
    LOCR_CUDA_OK(cudaGetLastError());
    // the pinned coefficient buffer is reused by the next call of this thread: wait until the copy has left it
    LOCR_CUDA_OK(cudaStreamSynchronize(s));
    // the scratch buffers only ever grow: hand the memory of an unusually large request back instead of keeping
    // gigabytes pinned in a serving process
    if (coef_total * 2 > ((size_t)1 << 30)) {
        host_coef.release();
        engine_release(h, "jpeg.coef");
        engine_release(h, "jpeg.planes");
    }
    return LOCR_OK;
}

}  // namespace locr
