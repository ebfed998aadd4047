// The polygon path of getDetBoxes (reference ocr/tools/det_utils.py:97-245 `poly_core`, reached with poly=True /
// CRAFT.enablePoly): for every kept box a 14-point polygon that follows a curved word, or "none".
//
// One CTA per (image, box):
//   thread 0   : box edge lengths in float32 (np.linalg.norm of float32 points), the perspective matrix box -> (w x h)
//                rectangle (cv2.getPerspectiveTransform: the 8x8 system solved by Gaussian elimination with partial
//                pivoting, double precision) and its inverse (cofactors);
//   all threads: one column of the warped word each - warpPerspective(INTER_NEAREST) evaluated on the fly
//                (source pixel = cvRound of the projected coordinate, round half to even; 0 outside the map), compared
//                with the box's label id, first / last set row per column;
//   thread 0   : the reference's sequential pivot search over the column list, the end-point probing with 8-connected
//                lines (cv::clipLine + cv::LineIterator restated on integers) and the back-projection of the 14 points.
// Everything the reference computes in Python floats is double precision here, in the reference's order of operations.
// oracle/poly_ref.py is the blueprint (pinned against the live reference and the live cv2); tests/test_poly_gpu.py
// compares this kernel with it: the same boxes get polygons, coordinates agree to ~1e-9.
// Compiled with --fmad=false like postproc.cu: the float32 edge lengths decide w and h.
#include "polys.cuh"

#include <math.h>

namespace locr {

namespace {

constexpr int kMaxCols = 4096;      // widest warped word (score maps are at most 1024 x 1024: the diagonal is 1449)
constexpr int kNumCp = 5;
constexpr int kTotSeg = 2 * kNumCp + 1;

struct Warp {
    double m[9];     // inverse of the perspective matrix: destination (x, y) -> source coordinates
    const int32_t* labels;
    int H, W, id;
};

__device__ __forceinline__ bool word_at(const Warp& wp, int x, int y) {
    const double X0 = wp.m[0] * x + wp.m[1] * y + wp.m[2];
    const double Y0 = wp.m[3] * x + wp.m[4] * y + wp.m[5];
    double Wd = wp.m[6] * x + wp.m[7] * y + wp.m[8];
    Wd = Wd != 0.0 ? 1.0 / Wd : 0.0;
    const double fx = fmax(-2147483648.0, fmin(2147483647.0, X0 * Wd));
    const double fy = fmax(-2147483648.0, fmin(2147483647.0, Y0 * Wd));
    const long long xi = __double2ll_rn(fx), yi = __double2ll_rn(fy);      // cvRound: round half to even
    if (xi < 0 || xi >= wp.W || yi < 0 || yi >= wp.H) return false;
    return wp.labels[yi * wp.W + xi] == wp.id;
}

// cv2.getPerspectiveTransform(src, dst): false when the system is singular.
__device__ bool perspective(const float* src, const double* dst, double* m) {
    double a[8][8], b[8];
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) a[i][j] = 0.0;
    for (int i = 0; i < 4; ++i) {
        const double sx = src[2 * i], sy = src[2 * i + 1], dx = dst[2 * i], dy = dst[2 * i + 1];
        a[i][0] = a[i + 4][3] = sx;
        a[i][1] = a[i + 4][4] = sy;
        a[i][2] = a[i + 4][5] = 1.0;
        a[i][6] = -sx * dx;
        a[i][7] = -sy * dx;
        a[i + 4][6] = -sx * dy;
        a[i + 4][7] = -sy * dy;
        b[i] = dx;
        b[i + 4] = dy;
    }
    for (int i = 0; i < 8; ++i) {
        int k = i;
        for (int j = i + 1; j < 8; ++j)
            if (fabs(a[j][i]) > fabs(a[k][i])) k = j;
        if (fabs(a[k][i]) < 2.220446049250313e-14) return false;
        if (k != i) {
            for (int j = 0; j < 8; ++j) { const double t = a[i][j]; a[i][j] = a[k][j]; a[k][j] = t; }
            const double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        const double d = -1.0 / a[i][i];
        for (int j = i + 1; j < 8; ++j) {
            const double alpha = a[j][i] * d;
            for (int kk = i + 1; kk < 8; ++kk) a[j][kk] = a[j][kk] + alpha * a[i][kk];
            b[j] = b[j] + alpha * b[i];
        }
    }
    for (int i = 7; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < 8; ++k) s = s - a[i][k] * m[k];
        m[i] = s / a[i][i];
    }
    m[8] = 1.0;
    return true;
}

__device__ bool invert3(const double* m, double* t) {
    const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                       m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (det == 0.0) return false;
    const double d = 1.0 / det;
    t[0] = (m[4] * m[8] - m[5] * m[7]) * d;
    t[1] = (m[2] * m[7] - m[1] * m[8]) * d;
    t[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    t[3] = (m[5] * m[6] - m[3] * m[8]) * d;
    t[4] = (m[0] * m[8] - m[2] * m[6]) * d;
    t[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    t[6] = (m[3] * m[7] - m[4] * m[6]) * d;
    t[7] = (m[1] * m[6] - m[0] * m[7]) * d;
    t[8] = (m[0] * m[4] - m[1] * m[3]) * d;
    return true;
}

__device__ __forceinline__ long long cdiv(long long a, long long b) { return a / b; }   // C++: truncation toward zero

// cv::clipLine(Size(w, h), pt1, pt2)
__device__ bool clip_line(int w, int h, long long& x1, long long& y1, long long& x2, long long& y2) {
    const long long right = w - 1, bottom = h - 1;
    if (w <= 0 || h <= 0) return false;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += cdiv((a - y1) * (x2 - x1), (y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += cdiv((a - y2) * (x2 - x1), (y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += cdiv((a - x1) * (y2 - y1), (x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += cdiv((a - x2) * (y2 - y1), (x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// np.sum(np.logical_and(word_label, line_img)) != 0 for cv2.line(line_img, p1, p2, 1, thickness=1)
__device__ bool line_hits_word(const Warp& wp, int w, int h, long long x1, long long y1, long long x2, long long y2) {
    if (!clip_line(w, h, x1, y1, x2, y2)) return false;
    long long dx = x2 - x1, dy = y2 - y1;
    if (dx < 0) {       // cv::Line walks left to right
        x1 = x2; y1 = y2;
        dx = -dx; dy = -dy;
    }
    const int sy = dy >= 0 ? 1 : -1;
    if (dy < 0) dy = -dy;
    long long x = x1, y = y1;
    if (dx >= dy) {
        long long err = dx - 2 * dy;
        for (long long i = 0; i <= dx; ++i) {
            if (word_at(wp, (int)x, (int)y)) return true;
            if (err < 0) { y += sy; err += 2 * dx; }
            err -= 2 * dy;
            x += 1;
        }
    } else {
        long long err = dy - 2 * dx;
        for (long long i = 0; i <= dy; ++i) {
            if (word_at(wp, (int)x, (int)y)) return true;
            if (err < 0) { x += 1; err += 2 * dy; }
            err -= 2 * dx;
            y += sy;
        }
    }
    return false;
}

// int(float64): truncation toward zero (Python int())
__device__ __forceinline__ long long pyint(double v) { return (long long)v; }

__global__ void __launch_bounds__(128)
pp_polys_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ box_label,
                const int32_t* __restrict__ counts, const int32_t* __restrict__ labels, int H, int W, int max_boxes,
                double* __restrict__ polys, int32_t* __restrict__ valid) {
    __shared__ Warp wp;
    __shared__ int s_w, s_h, s_ok;
    __shared__ short first_row[kMaxCols], last_row[kMaxCols];
    const int b = blockIdx.y, k = blockIdx.x;
    if (k >= counts[2 * b] || k >= max_boxes) return;
    const float* box = boxes + ((size_t)b * max_boxes + k) * 8;
    double* out = polys + ((size_t)b * max_boxes + k) * 28;
    int32_t* ok_out = valid + (size_t)b * max_boxes + k;
    if (threadIdx.x == 0) {
        s_ok = 0;
        // w, h = int(np.linalg.norm(box[0] - box[1]) + 1), int(np.linalg.norm(box[1] - box[2]) + 1), all float32
        const float d0x = box[0] - box[2], d0y = box[1] - box[3];
        const float d1x = box[2] - box[4], d1y = box[3] - box[5];
        const int w = (int)(sqrtf(d0x * d0x + d0y * d0y) + 1.0f);
        const int h = (int)(sqrtf(d1x * d1x + d1y * d1y) + 1.0f);
        s_w = w; s_h = h;
        if (w >= 10 && h >= 10 && w <= kMaxCols && h < 32768) {
            const double tar[8] = {0.0, 0.0, (double)w, 0.0, (double)w, (double)h, 0.0, (double)h};
            double m[9];
            if (perspective(box, tar, m) && invert3(m, wp.m)) {
                wp.labels = labels + (size_t)b * H * W;
                wp.H = H; wp.W = W;
                wp.id = box_label[(size_t)b * max_boxes + k];
                s_ok = 1;
            }
        }
        *ok_out = 0;
    }
    __syncthreads();
    if (!s_ok) return;
    const int w = s_w, h = s_h;
    // per column: first and last row of the word (the reference's cp list needs at least two set pixels)
    for (int x = threadIdx.x; x < w; x += blockDim.x) {
        int f = -1, l = -1, n = 0;
        for (int y = 0; y < h; ++y)
            if (word_at(wp, x, y)) {
                if (f < 0) f = y;
                l = y;
                ++n;
            }
        first_row[x] = (short)(n >= 2 ? f : -1);
        last_row[x] = (short)l;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;

    int max_len = -1;
    for (int x = 0; x < w; ++x)
        if (first_row[x] >= 0) max_len = max(max_len, last_row[x] - first_row[x] + 1);
    if ((double)h * 0.7 < (double)max_len) return;       // the word fills the box: no polygon needed

    // pivot points with fixed length (det_utils.py:139-176)
    const double seg_w = (double)w / kTotSeg;
    double pp_x[kNumCp], pp_y[kNumCp];
    bool pp_set[kNumCp];
    double sec_x[kTotSeg], sec_y[kTotSeg];
    int seg_height[kNumCp];
    for (int i = 0; i < kNumCp; ++i) { pp_set[i] = false; seg_height[i] = 0; pp_x[i] = pp_y[i] = 0.0; }
    for (int i = 0; i < kTotSeg; ++i) sec_x[i] = sec_y[i] = 0.0;
    int seg_num = 0, num_sec = 0, prev_h = -1;
    for (int x = 0; x < w; ++x) {
        if (first_row[x] < 0) continue;
        const int sy = first_row[x], ey = last_row[x];
        if ((double)(seg_num + 1) * seg_w <= (double)x && seg_num <= kTotSeg) {
            if (num_sec == 0) break;
            sec_x[seg_num] = sec_x[seg_num] / num_sec;
            sec_y[seg_num] = sec_y[seg_num] / num_sec;
            num_sec = 0;
            seg_num += 1;
            prev_h = -1;
        }
        if (seg_num >= kTotSeg) return;     // the reference would raise IndexError here; cannot happen: x < w = 11 seg_w
        const double cy = (double)(sy + ey) * 0.5;
        const int cur_h = ey - sy + 1;
        sec_x[seg_num] = sec_x[seg_num] + (double)x;
        sec_y[seg_num] = sec_y[seg_num] + cy;
        num_sec += 1;
        if (seg_num % 2 == 0) continue;
        if (prev_h < cur_h) {
            const int j = (seg_num - 1) / 2;
            pp_x[j] = (double)x; pp_y[j] = cy; pp_set[j] = true;
            seg_height[j] = cur_h;
            prev_h = cur_h;
        }
    }
    if (num_sec != 0) {
        sec_x[kTotSeg - 1] = sec_x[kTotSeg - 1] / num_sec;
        sec_y[kTotSeg - 1] = sec_y[kTotSeg - 1] / num_sec;
    }
    int hmax = 0;
    for (int i = 0; i < kNumCp; ++i) {
        if (!pp_set[i]) return;
        hmax = max(hmax, seg_height[i]);
    }
    if (seg_w < (double)hmax * 0.25) return;
    // np.median of five integers
    int srt[kNumCp];
    for (int i = 0; i < kNumCp; ++i) srt[i] = seg_height[i];
    for (int i = 1; i < kNumCp; ++i) {
        const int v = srt[i];
        int j = i - 1;
        while (j >= 0 && srt[j] > v) { srt[j + 1] = srt[j]; --j; }
        srt[j + 1] = v;
    }
    const double half_char_h = (double)srt[kNumCp / 2] * 1.45 / 2;

    double np_[kNumCp][4];
    for (int i = 0; i < kNumCp; ++i) {
        const double dx = sec_x[i * 2 + 2] - sec_x[i * 2];
        const double dy = sec_y[i * 2 + 2] - sec_y[i * 2];
        if (dx == 0.0) {
            np_[i][0] = pp_x[i]; np_[i][1] = pp_y[i] - half_char_h; np_[i][2] = pp_x[i]; np_[i][3] = pp_y[i] + half_char_h;
            continue;
        }
        const double rad = -atan2(dy, dx);
        const double c = half_char_h * cos(rad), s = half_char_h * sin(rad);
        np_[i][0] = pp_x[i] - s; np_[i][1] = pp_y[i] - c; np_[i][2] = pp_x[i] + s; np_[i][3] = pp_y[i] + c;
    }
    // edge points that clear the character heat (det_utils.py:199-222)
    const double grad_s = (pp_y[1] - pp_y[0]) / (pp_x[1] - pp_x[0]) + (pp_y[2] - pp_y[1]) / (pp_x[2] - pp_x[1]);
    const double grad_e = (pp_y[3] - pp_y[4]) / (pp_x[3] - pp_x[4]) + (pp_y[2] - pp_y[3]) / (pp_x[2] - pp_x[3]);
    bool found_s = false, found_e = false;
    double spp[4], epp[4];
    for (int it = 0; it < 8; ++it) {                          // np.arange(0.5, 2.0, 0.2)
        const double r = 0.5 + it * 0.2;
        const double dx = 2 * half_char_h * r;
        if (!found_s) {
            const double dy = grad_s * dx;
            const double p[4] = {np_[0][0] - dx, np_[0][1] - dy, np_[0][2] - dx, np_[0][3] - dy};
            if (!line_hits_word(wp, w, h, pyint(p[0]), pyint(p[1]), pyint(p[2]), pyint(p[3])) || r + 2 * 0.2 >= 2.0) {
                for (int q = 0; q < 4; ++q) spp[q] = p[q];
                found_s = true;
            }
        }
        if (!found_e) {
            const double dy = grad_e * dx;
            const double p[4] = {np_[kNumCp - 1][0] + dx, np_[kNumCp - 1][1] + dy, np_[kNumCp - 1][2] + dx,
                                 np_[kNumCp - 1][3] + dy};
            if (!line_hits_word(wp, w, h, pyint(p[0]), pyint(p[1]), pyint(p[2]), pyint(p[3])) || r + 2 * 0.2 >= 2.0) {
                for (int q = 0; q < 4; ++q) epp[q] = p[q];
                found_e = true;
            }
        }
        if (found_s && found_e) break;
    }
    if (!(found_s && found_e)) return;

    // back-projection: warp_coord(Minv, pt) with Minv = inverse of the perspective matrix (= wp.m)
    auto unwarp = [&](double px, double py, double* o) {
        const double ox = wp.m[0] * px + wp.m[1] * py + wp.m[2];
        const double oy = wp.m[3] * px + wp.m[4] * py + wp.m[5];
        const double oz = wp.m[6] * px + wp.m[7] * py + wp.m[8];
        o[0] = ox / oz;
        o[1] = oy / oz;
    };
    int n = 0;
    unwarp(spp[0], spp[1], out + 2 * n++);
    for (int i = 0; i < kNumCp; ++i) unwarp(np_[i][0], np_[i][1], out + 2 * n++);
    unwarp(epp[0], epp[1], out + 2 * n++);
    unwarp(epp[2], epp[3], out + 2 * n++);
    for (int i = kNumCp - 1; i >= 0; --i) unwarp(np_[i][2], np_[i][3], out + 2 * n++);
    unwarp(spp[2], spp[3], out + 2 * n++);
    *ok_out = 1;
}

}  // namespace

void launch_polys(const float* boxes, const int32_t* box_label, const int32_t* counts, const int32_t* labels, int B,
                  int H, int W, int max_boxes, double* polys, int32_t* valid, cudaStream_t s) {
    if (B <= 0 || max_boxes <= 0) return;
    pp_polys_kernel<<<dim3(max_boxes, B), 128, 0, s>>>(boxes, box_label, counts, labels, H, W, max_boxes, polys, valid);
}

}  // namespace locr
