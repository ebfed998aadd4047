// AddressSanitizer / UBSan harness for the host half of the PNG reader (chunk walk + CRC + zlib inflate into the
// contiguous and the slot-padded scanline layouts - the part that touches untrusted bytes): mutated files (byte flips,
// truncation, IHDR corruption with the CRC repaired, insertions, deletions) through parse / inflate_idat with exact-size
// heap buffers.  CPU only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 --expt-relaxed-constexpr -I include \
//        -I lightly_ocr_b200/csrc -Xcompiler -fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer \
//        -o /tmp/fuzz_png tools/fuzz_png.cu -lasan -lubsan -lz
//   ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0 /tmp/fuzz_png 300000 seed0.png seed1.png ...
// Seeds: one file per colour type x bit depth x interlace flag with random filter bytes, split IDAT chunks, an
// ancillary chunk in front of the image data, and one cv2.imencode file.
#include "../lightly_ocr_b200/csrc/png.cu"
#include <random>
#include <fstream>
namespace locr {
std::string& tls_error() { static thread_local std::string e; return e; }
int fail(int code, const std::string& m) { tls_error() = m; return code; }
void* engine_buffer(locr_handle*, const std::string&, size_t) { return nullptr; }
}
static void put32(uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }
int main(int argc, char** argv) {
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 2; i < argc; ++i) {
        std::ifstream f(argv[i], std::ios::binary);
        seeds.emplace_back((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    }
    const int N = atoi(argv[1]);
    std::mt19937 rng(argc * 7919 + atoi(argv[1]));
    long ok = 0, bad = 0, hdr_ok = 0;
    for (int it = 0; it < N; ++it) {
        std::vector<uint8_t> d = seeds[it % seeds.size()];
        const int kind = rng() % 6;
        if (kind == 0) { for (int k = 0, n = 1 + rng() % 5; k < n; ++k) d[rng() % d.size()] = rng() & 255; }
        else if (kind == 1) { d.resize(8 + rng() % (d.size() - 8)); }
        else if (kind == 2 && d.size() >= 33) {
            // a changed IHDR (size, depth, colour type, interlace) with its CRC repaired: the declared geometry no longer
            // matches the compressed data, which is what the buffer sizing must survive
            const int f = rng() % 5;
            if (f == 0) put32(&d[16], 1 + rng() % 70);
            else if (f == 1) put32(&d[20], 1 + rng() % 70);
            else if (f == 2) { const int dd[5] = {1, 2, 4, 8, 16}; d[24] = dd[rng() % 5]; }
            else if (f == 3) { const int cc[5] = {0, 2, 3, 4, 6}; d[25] = cc[rng() % 5]; }
            else d[28] ^= 1;
            put32(&d[29], (uint32_t)crc32(0L, &d[12], 17));
        }
        else if (kind == 3) { size_t p = 8 + rng() % (d.size() - 8); int n = 1 + rng() % 20; std::vector<uint8_t> j(n); for (auto& b : j) b = rng() & 255; d.insert(d.begin() + p, j.begin(), j.end()); }
        else if (kind == 4) { size_t p = 8 + rng() % (d.size() - 9); size_t q = std::min(d.size(), p + 1 + rng() % 40); d.erase(d.begin() + p, d.begin() + q); }
        else {
            // damage inside an IDAT body with the chunk CRC repaired: corrupt deflate data reaches inflate
            size_t pos = 8;
            while (pos + 12 <= d.size()) {
                const size_t len = locr::be32(&d[pos]);
                if (pos + 12 + len > d.size()) break;
                if (memcmp(&d[pos + 4], "IDAT", 4) == 0 && len > 0) {
                    for (int k = 0, n = 1 + rng() % 3; k < n; ++k) d[pos + 8 + rng() % len] = rng() & 255;
                    put32(&d[pos + 8 + len], (uint32_t)crc32(0L, &d[pos + 4], (uInt)(len + 4)));
                    break;
                }
                pos += 12 + len;
            }
        }
        // exact-size heap copy so that ASan sees any read past the end of the file
        uint8_t* buf = new uint8_t[d.size()];
        memcpy(buf, d.data(), d.size());
        std::string err;
        locr::Header hd;
        if (locr::parse(buf, d.size(), &hd, &err)) {
            ++hdr_ok;
            bool good = true;
            if (hd.filtered_bytes <= (64u << 20)) {
                uint8_t* a = new uint8_t[hd.filtered_bytes];      // exact sizes: any write past the end is caught
                good = locr::inflate_idat(hd, a, false, &err);
                delete[] a;
                uint8_t* b = new uint8_t[hd.padded_bytes];
                const bool good2 = locr::inflate_idat(hd, b, true, &err);
                delete[] b;
                if (good != good2) { printf("layouts disagree on input %d\n", it); return 1; }
            }
            (good ? ok : bad)++;
        } else {
            ++bad;
        }
        size_t need = 0;
        int hh, ww, cc;
        locr::png_probe(buf, d.size(), &hh, &ww, &cc, &err);
        locr::png_host_scanlines(buf, d.size(), nullptr, 0, &need, &err);
        delete[] buf;
    }
    printf("asan fuzz: %d inputs, %ld headers accepted, %ld decoded, %ld refused\n", N, hdr_ok, ok, bad);
    return 0;
}
