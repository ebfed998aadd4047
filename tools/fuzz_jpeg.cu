// AddressSanitizer / UBSan harness for the host half of the JPEG reader (marker parsing + entropy decoding, the part that
// touches untrusted bytes): mutated files (byte flips, truncation, header corruption, insertions, deletions) through
// jpeg_host_coefficients with exact-size heap buffers.  CPU only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 --expt-relaxed-constexpr -I include \
//        -I lightly_ocr_b200/csrc -Xcompiler -fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer \
//        -o /tmp/fuzz_jpeg tools/fuzz_jpeg.cu -lasan -lubsan
//   ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0 /tmp/fuzz_jpeg 300000 seed0.jpg seed1.jpg ...
// Round 1: 500 000 inputs from 7 seed files (baseline, progressive, restart intervals, 4:4:4 / 4:1:1, gray, EXIF) clean
// after the over-subscribed-Huffman-table check went in (the first run found that heap overflow in build_table).
#include "../lightly_ocr_b200/csrc/jpeg.cu"
#include <random>
#include <fstream>
namespace locr {
std::string& tls_error() { static thread_local std::string e; return e; }
int fail(int code, const std::string& m) { tls_error() = m; return code; }
void* engine_buffer(locr_handle*, const std::string&, size_t) { return nullptr; }
}
int main(int argc, char** argv) {
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 2; i < argc; ++i) {
        std::ifstream f(argv[i], std::ios::binary);
        seeds.emplace_back((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    }
    const int N = atoi(argv[1]);
    std::mt19937 rng(argc * 7919 + atoi(argv[1]));
    long ok = 0, bad = 0;
    for (int it = 0; it < N; ++it) {
        std::vector<uint8_t> d = seeds[it % seeds.size()];
        const int kind = rng() % 5;
        if (kind == 0) { for (int k = 0, n = 1 + rng() % 5; k < n; ++k) d[rng() % d.size()] = rng() & 255; }
        else if (kind == 1) { d.resize(2 + rng() % (d.size() - 2)); }
        else if (kind == 2) { for (int k = 0, n = 1 + rng() % 3; k < n; ++k) d[2 + rng() % std::min<size_t>(d.size() - 2, 700)] = rng() & 255; }
        else if (kind == 3) { size_t p = 2 + rng() % (d.size() - 2); int n = 1 + rng() % 20; std::vector<uint8_t> j(n); for (auto& b : j) b = rng() & 255; d.insert(d.begin() + p, j.begin(), j.end()); }
        else { size_t p = 2 + rng() % (d.size() - 3); size_t q = std::min(d.size(), p + 1 + rng() % 40); d.erase(d.begin() + p, d.begin() + q); }
        // exact-size heap copy so that ASan sees any read past the end of the file
        uint8_t* buf = new uint8_t[d.size()];
        memcpy(buf, d.data(), d.size());
        int info[19];
        std::string err;
        int rc = locr::jpeg_host_coefficients(buf, d.size(), nullptr, 0, info, &err);
        if (rc == 0 && (long)info[0] * info[1] <= 4000000) {
            size_t elems = 0;
            for (int k = 0; k < info[2]; ++k) elems += (size_t)info[9 + 4 * k] * info[10 + 4 * k] * 64;
            int16_t* out = new int16_t[elems];      // exact size: any write past the planes is caught
            rc = locr::jpeg_host_coefficients(buf, d.size(), out, elems, info, &err);
            delete[] out;
        }
        (rc == 0 ? ok : bad)++;
        delete[] buf;
    }
    printf("asan fuzz: %d inputs, %ld decoded, %ld refused\n", N, ok, bad);
    return 0;
}
