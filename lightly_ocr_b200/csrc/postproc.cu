#include "postproc.cuh"

#include <float.h>
#include <math.h>

namespace locr {

namespace {

constexpr int kCap = 65536;      // components per image the statistics tables can hold
constexpr int kScanBlock = 1024; // pixels per scan block
constexpr int kMaxRows = 1024;   // tallest score map supported by the box kernel

struct Work {
    int* parent;      // [npix] union-find parent (pixel index within the batch), -1 for background
    uint8_t* flags;   // [npix] bit0 = text_score, bit1 = link_score
    int* rootid;      // [npix] component id (1-based, per image) stored at root pixels
    int* blocksum;    // [B][bpi]
    int* ncomp;       // [B]
    int* area;        // [B][kCap] ...
    int* minx;
    int* miny;
    int* maxx;
    int* maxy;
    int* maxtext;     // orderable-int view of the float maximum
    int* root;        // root pixel index (within the batch)
    int* valid;       // [B][kCap]
    float* cbox;      // [B][kCap][8]
    int* crect;       // [B][kCap][4]
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

Work carve(void* ws, int B, int H, int W) {
    Work w;
    uint8_t* p = reinterpret_cast<uint8_t*>(ws);
    const size_t npix = (size_t)B * H * W;
    const int bpi = (H * W + kScanBlock - 1) / kScanBlock;
    auto take = [&](size_t bytes) {
        void* r = p;
        p += align256(bytes);
        return r;
    };
    w.parent = (int*)take(npix * 4);
    w.flags = (uint8_t*)take(npix);
    w.rootid = (int*)take(npix * 4);
    w.blocksum = (int*)take((size_t)B * bpi * 4);
    w.ncomp = (int*)take((size_t)B * 4);
    const size_t c = (size_t)B * kCap;
    w.area = (int*)take(c * 4);
    w.minx = (int*)take(c * 4);
    w.miny = (int*)take(c * 4);
    w.maxx = (int*)take(c * 4);
    w.maxy = (int*)take(c * 4);
    w.maxtext = (int*)take(c * 4);
    w.root = (int*)take(c * 4);
    w.valid = (int*)take(c * 4);
    w.cbox = (float*)take(c * 8 * 4);
    w.crect = (int*)take(c * 4 * 4);
    return w;
}

__device__ __forceinline__ int float_orderable(float f) {
    const int b = __float_as_int(f);
    return b >= 0 ? b : (b ^ 0x7fffffff);
}
__device__ __forceinline__ float orderable_float(int o) { return __int_as_float(o >= 0 ? o : (o ^ 0x7fffffff)); }

__device__ __forceinline__ int uf_find(const int* parent, int i) {
    int p = parent[i];
    while (p != i) {
        i = p;
        p = parent[i];
    }
    return i;
}
__device__ __forceinline__ void uf_unite(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(&parent[a], b);  // the smaller raster index becomes the root
        if (old == a) return;
        a = old;
    }
}

// Thresholds + tile-local labelling with shared-memory staging.
//   text_score = text > low_text, link_score = link > link_threshold (cv2.threshold THRESH_BINARY is strict);
//   foreground = clip(text_score + link_score, 0, 1)                                   (det_utils.py:37-44)
// One CTA = one 32 x 32 tile of one score map: the flags and a union-find over LOCAL indices live in shared memory
// (4-connectivity, atomicMin so that the smallest raster index becomes the root, which is also the smallest global
// index of the tile's part of the component), then every pixel is written to the global parent array already
// flattened to its tile root.  pp_merge_borders afterwards unites across tile edges only (1/16 of the pixels).
constexpr int kTile = 32;

__device__ __forceinline__ int uf_find_s(const int* parent, int i) {
    int p = parent[i];
    while (p != i) {
        i = p;
        p = parent[i];
    }
    return i;
}
__device__ __forceinline__ void uf_unite_s(int* parent, int a, int b) {
    while (true) {
        a = uf_find_s(parent, a);
        b = uf_find_s(parent, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

__global__ void __launch_bounds__(256)
pp_label_tiles(const float2* __restrict__ score, int H, int W, float low_text, float link_thr,
               int* __restrict__ parent, uint8_t* __restrict__ flags) {
    __shared__ int lab[kTile * kTile];
    __shared__ uint8_t fl[kTile * kTile];
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const int img = blockIdx.z * H * W;
    for (int l = threadIdx.x; l < kTile * kTile; l += 256) {
        const int x = x0 + (l & (kTile - 1)), y = y0 + (l >> 5);
        int f = 0;
        if (x < W && y < H) {
            const float2 v = __ldg(&score[img + y * W + x]);
            f = (v.x > low_text ? 1 : 0) | (v.y > link_thr ? 2 : 0);
        }
        fl[l] = (uint8_t)f;
        lab[l] = f ? l : -1;
    }
    __syncthreads();
    for (int l = threadIdx.x; l < kTile * kTile; l += 256) {
        if (!fl[l]) continue;
        if ((l & (kTile - 1)) > 0 && fl[l - 1]) uf_unite_s(lab, l, l - 1);
        if (l >= kTile && fl[l - kTile]) uf_unite_s(lab, l, l - kTile);
    }
    __syncthreads();
    for (int l = threadIdx.x; l < kTile * kTile; l += 256) {
        const int x = x0 + (l & (kTile - 1)), y = y0 + (l >> 5);
        if (x >= W || y >= H) continue;
        const int g = img + y * W + x;
        flags[g] = fl[l];
        if (fl[l]) {
            const int r = uf_find_s(lab, l);
            parent[g] = img + (y0 + (r >> 5)) * W + x0 + (r & (kTile - 1));
        } else {
            parent[g] = -1;
        }
    }
}

// Unites the tile-local components across tile edges: only pixels of a tile's first column / first row have a
// neighbour in another tile.  One thread per edge pixel.
__global__ void __launch_bounds__(256)
pp_merge_borders(int* __restrict__ parent, const uint8_t* __restrict__ flags, int B, int H, int W) {
    const int tx = (W + kTile - 1) / kTile, ty = (H + kTile - 1) / kTile;
    const int per_img = (tx - 1) * H + (ty - 1) * W;      // vertical seams, then horizontal seams
    const int total = B * per_img;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int b = i / per_img;
        int r = i - b * per_img;
        int x, y, other;
        if (r < (tx - 1) * H) {               // pixel (x = k * 32, y) against its left neighbour
            x = (r / H + 1) * kTile;
            y = r % H;
            other = -1;
        } else {                              // pixel (x, y = k * 32) against its upper neighbour
            r -= (tx - 1) * H;
            y = (r / W + 1) * kTile;
            x = r % W;
            other = -W;
        }
        const int g = b * H * W + y * W + x;
        if (flags[g] && flags[g + other]) uf_unite(parent, g, g + other);
    }
}

// Flatten every pixel to its root and count the roots of each 1024-pixel block (per image).
__global__ void __launch_bounds__(256) pp_flatten_count(int* __restrict__ parent, int HW, int bpi,
                                                         int* __restrict__ blocksum) {
    const int b = blockIdx.y, blk = blockIdx.x;
    const long base = (long)b * HW;
    int cnt = 0;
    for (int k = threadIdx.x; k < kScanBlock; k += 256) {
        const int li = blk * kScanBlock + k;
        if (li < HW) {
            const long i = base + li;
            const int p = parent[i];
            if (p >= 0) {
                const int r = uf_find(parent, (int)i);
                parent[i] = r;
                cnt += (r == (int)i);
            }
        }
    }
    __shared__ int red[8];
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int q = 0; q < 8; ++q) s += red[q];
        blocksum[b * bpi + blk] = s;
    }
}

// Exclusive scan of the block counts of one image (bpi <= 1024) -> offsets; total -> ncomp.
__global__ void __launch_bounds__(1024) pp_scan(int* __restrict__ blocksum, int bpi, int* __restrict__ ncomp) {
    __shared__ int s[1024];
    const int b = blockIdx.x, t = threadIdx.x;
    const int v = t < bpi ? blocksum[b * bpi + t] : 0;
    s[t] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int add = t >= o ? s[t - o] : 0;
        __syncthreads();
        s[t] += add;
        __syncthreads();
    }
    if (t < bpi) blocksum[b * bpi + t] = s[t] - v;
    if (t == 1023) ncomp[b] = s[t];
}

// Raster-ordered component ids: id = 1 + number of roots before this root in the image (cv2's label numbering).
__global__ void __launch_bounds__(1024) pp_assign(const int* __restrict__ parent, int HW, int bpi,
                                                   const int* __restrict__ blocksum, int* __restrict__ rootid,
                                                   int* area, int* minx, int* miny, int* maxx, int* maxy, int* maxtext,
                                                   int* root) {
    __shared__ int s[1024];
    const int b = blockIdx.y, blk = blockIdx.x, t = threadIdx.x;
    const int li = blk * kScanBlock + t;
    const long i = (long)b * HW + li;
    const int isroot = (li < HW && parent[i] == (int)i) ? 1 : 0;
    s[t] = isroot;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int add = t >= o ? s[t - o] : 0;
        __syncthreads();
        s[t] += add;
        __syncthreads();
    }
    if (isroot) {
        const int id = blocksum[b * bpi + blk] + s[t];  // inclusive scan -> 1-based id
        rootid[i] = id;
        if (id < kCap) {
            const int c = b * kCap + id;
            area[c] = 0;
            minx[c] = 0x7fffffff;
            miny[c] = 0x7fffffff;
            maxx[c] = -1;
            maxy[c] = -1;
            maxtext[c] = (int)0x80000000;
            root[c] = (int)i;
        }
    }
}

__global__ void pp_stats(const float2* __restrict__ score, const int* __restrict__ parent,
                         const int* __restrict__ rootid, int B, int H, int W, int* area, int* minx, int* miny,
                         int* maxx, int* maxy, int* maxtext, int32_t* __restrict__ labels_out) {
    // Neighbouring pixels mostly belong to the same component: the lanes of a warp that share a component combine their
    // contribution with warp reductions and ONE lane issues the six atomics (sum / min / max commute, so the tables come
    // out exactly as with one atomic per pixel).
    const long npix = (long)B * H * W;
    const int lane = threadIdx.x & 31;
    for (long base = (long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < npix; base += (long)gridDim.x * blockDim.x) {
        const long i = base + lane;
        int id = 0, c = -1, x = 0, y = 0, t = (int)0x80000000;
        if (i < npix) {
            const int r = parent[i];
            if (r >= 0) {
                id = rootid[r];
                if (id < kCap) {
                    x = (int)(i % W);
                    y = (int)((i / W) % H);
                    c = (int)(i / ((long)W * H)) * kCap + id;
                    t = float_orderable(__ldg(&score[i]).x);
                }
            }
            if (labels_out != nullptr) labels_out[i] = id;
        }
        const unsigned fg = __ballot_sync(0xffffffffu, c >= 0);
        if (c >= 0) {
            const unsigned peers = __match_any_sync(fg, c);
            const int cnt = __popc(peers);
            const int mnx = __reduce_min_sync(peers, x), mxx = __reduce_max_sync(peers, x);
            const int mny = __reduce_min_sync(peers, y), mxy = __reduce_max_sync(peers, y);
            const int mt = __reduce_max_sync(peers, t);
            if (lane == __ffs(peers) - 1) {
                atomicAdd(&area[c], cnt);
                atomicMin(&minx[c], mnx);
                atomicMax(&maxx[c], mxx);
                atomicMin(&miny[c], mny);
                atomicMax(&maxy[c], mxy);
                atomicMax(&maxtext[c], mt);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One warp per component: det_boxes_core's per-label body (reference det_utils.py:49-92).
struct Pt {
    short x, y;
};

__device__ __forceinline__ long cross_uv(Pt o, Pt a, Pt b) {
    // coordinates swapped (u = y, v = x): sweeping rows top to bottom
    return (long)(a.y - o.y) * (b.x - o.x) - (long)(a.x - o.x) * (b.y - o.y);
}

// OpenCV rotatingCalipers(CALIPERS_MINAREARECT) + minAreaRect + boxPoints, float32 statement for statement
// (imgproc/src/rotcalipers.cpp); `ring` holds the hull in cv2.convexHull(clockwise=false) order.
__device__ void min_area_box(const Pt* ring, int n, float* vxs, float* vys, float* invs, float (&box)[8]) {
    float cx, cy, w, h;
    double ang;
    if (n > 2) {
        int left = 0, bottom = 0, right = 0, top = 0;
        float left_x, right_x, top_y, bottom_y;
        left_x = right_x = (float)ring[0].x;
        top_y = bottom_y = (float)ring[0].y;
        for (int i = 0; i < n; ++i) {
            const float px = (float)ring[i].x, py = (float)ring[i].y;
            if (px < left_x) { left_x = px; left = i; }
            if (px > right_x) { right_x = px; right = i; }
            if (py > top_y) { top_y = py; top = i; }
            if (py < bottom_y) { bottom_y = py; bottom = i; }
            const int j = (i + 1 < n) ? i + 1 : 0;
            const double dx = (double)ring[j].x - (double)ring[i].x;
            const double dy = (double)ring[j].y - (double)ring[i].y;
            vxs[i] = (float)dx;
            vys[i] = (float)dy;
            invs[i] = (float)(1. / sqrt(dx * dx + dy * dy));
        }
        float orientation = 0.f;
        {
            double ax = vxs[n - 1], ay = vys[n - 1];
            for (int i = 0; i < n; ++i) {
                const double bx = vxs[i], by = vys[i];
                const double convexity = ax * by - ay * bx;
                if (convexity != 0) {
                    orientation = (convexity > 0) ? 1.f : -1.f;
                    break;
                }
                ax = bx;
                ay = by;
            }
        }
        float base_a = orientation, base_b = 0.f;
        int seq[4] = {bottom, right, top, left};
        float minarea = FLT_MAX;
        int b_left = 0, b_bottom = 0;
        float b_a = 0.f, b_width = 0.f, b_b = 0.f, b_height = 0.f;
        for (int k = 0; k < n; ++k) {
            float dp[4];
            dp[0] = +base_a * vxs[seq[0]] + base_b * vys[seq[0]];
            dp[1] = -base_b * vxs[seq[1]] + base_a * vys[seq[1]];
            dp[2] = -base_a * vxs[seq[2]] - base_b * vys[seq[2]];
            dp[3] = +base_b * vxs[seq[3]] - base_a * vys[seq[3]];
            float maxcos = dp[0] * invs[seq[0]];
            int main_element = 0;
            for (int i = 1; i < 4; ++i) {
                const float cosalpha = dp[i] * invs[seq[i]];
                if (cosalpha > maxcos) {
                    main_element = i;
                    maxcos = cosalpha;
                }
            }
            {
                const int pindex = seq[main_element];
                const float lead_x = vxs[pindex] * invs[pindex];
                const float lead_y = vys[pindex] * invs[pindex];
                switch (main_element) {
                    case 0: base_a = lead_x; base_b = lead_y; break;
                    case 1: base_a = lead_y; base_b = -lead_x; break;
                    case 2: base_a = -lead_x; base_b = -lead_y; break;
                    default: base_a = -lead_y; base_b = lead_x; break;
                }
            }
            seq[main_element] += 1;
            seq[main_element] = (seq[main_element] == n) ? 0 : seq[main_element];
            float dx = (float)ring[seq[1]].x - (float)ring[seq[3]].x;
            float dy = (float)ring[seq[1]].y - (float)ring[seq[3]].y;
            const float width = dx * base_a + dy * base_b;
            dx = (float)ring[seq[2]].x - (float)ring[seq[0]].x;
            dy = (float)ring[seq[2]].y - (float)ring[seq[0]].y;
            const float height = -dx * base_b + dy * base_a;
            const float area = width * height;
            if (area <= minarea) {
                minarea = area;
                b_left = seq[3];
                b_a = base_a;
                b_width = width;
                b_b = base_b;
                b_height = height;
                b_bottom = seq[0];
            }
        }
        const float A1 = b_a, B1 = b_b, A2 = -b_b, B2 = b_a;
        const float C1 = A1 * (float)ring[b_left].x + (float)ring[b_left].y * B1;
        const float C2 = A2 * (float)ring[b_bottom].x + (float)ring[b_bottom].y * B2;
        const float idet = 1.f / (A1 * B2 - A2 * B1);
        const float px = (C1 * B2 - C2 * B1) * idet;
        const float py = (A1 * C2 - A2 * C1) * idet;
        const float o2 = A1 * b_width, o3 = B1 * b_width, o4 = A2 * b_height, o5 = B2 * b_height;
        cx = px + (o2 + o4) * 0.5f;
        cy = py + (o3 + o5) * 0.5f;
        w = (float)sqrt((double)o2 * o2 + (double)o3 * o3);
        h = (float)sqrt((double)o4 * o4 + (double)o5 * o5);
        ang = atan2((double)o3, (double)o2);
    } else if (n == 2) {
        cx = ((float)ring[0].x + (float)ring[1].x) * 0.5f;
        cy = ((float)ring[0].y + (float)ring[1].y) * 0.5f;
        const double dx = (double)ring[1].x - (double)ring[0].x, dy = (double)ring[1].y - (double)ring[0].y;
        w = (float)sqrt(dx * dx + dy * dy);
        h = 0.f;
        ang = atan2(dy, dx);
    } else {
        cx = n == 1 ? (float)ring[0].x : 0.f;
        cy = n == 1 ? (float)ring[0].y : 0.f;
        w = h = 0.f;
        ang = 0.0;
    }
    // OpenCV 4.13: the angle stays in double and is folded into [-90, 0) with width/height swaps
    ang = ang * 180.0 / 3.1415926535897932384626433832795;
    while (ang >= 0.0) {
        ang -= 90.0;
        const float t = w; w = h; h = t;
    }
    while (ang < -90.0) {
        ang += 90.0;
        const float t = w; w = h; h = t;
    }
    const float angle = (float)ang;
    // RotatedRect::points
    const double a_ = (double)angle * 3.1415926535897932384626433832795 / 180.;
    const float b = (float)cos(a_) * 0.5f;
    const float a = (float)sin(a_) * 0.5f;
    box[0] = cx - a * h - b * w;
    box[1] = cy + b * h - a * w;
    box[2] = cx + a * h - b * w;
    box[3] = cy - b * h - a * w;
    box[4] = 2 * cx - box[0];
    box[5] = 2 * cy - box[1];
    box[6] = 2 * cx - box[2];
    box[7] = 2 * cy - box[3];
}

constexpr int kBoxWarps = 2;

__global__ void __launch_bounds__(kBoxWarps * 32)
pp_boxes(const int* __restrict__ parent, const uint8_t* __restrict__ flags, const int* __restrict__ ncomp,
         const int* __restrict__ area_, const int* __restrict__ minx_, const int* __restrict__ miny_,
         const int* __restrict__ maxx_, const int* __restrict__ maxy_, const int* __restrict__ maxtext_,
         const int* __restrict__ root_, int B, int H, int W, float text_threshold, double scale_x, double scale_y,
         int* __restrict__ valid, float* __restrict__ cbox, int* __restrict__ crect) {
    __shared__ short s_rmin[kBoxWarps][kMaxRows], s_rmax[kBoxWarps][kMaxRows];
    __shared__ short s_dmin[kBoxWarps][kMaxRows], s_dmax[kBoxWarps][kMaxRows];
    __shared__ Pt s_ring[kBoxWarps][2 * kMaxRows];
    __shared__ float s_vx[kBoxWarps][256], s_vy[kBoxWarps][256], s_inv[kBoxWarps][256];
    const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    int n = ncomp[b];
    if (n >= kCap) n = kCap - 1;
    short* rmin = s_rmin[wp];
    short* rmax = s_rmax[wp];
    short* dmin = s_dmin[wp];
    short* dmax = s_dmax[wp];
    Pt* ring = s_ring[wp];
    for (int id = 1 + blockIdx.x * kBoxWarps + wp; id <= n; id += gridDim.x * kBoxWarps) {
        const int c = b * kCap + id;
        const int area = area_[c];
        bool ok = area >= 10;
        if (ok) ok = !(orderable_float(maxtext_[c]) < text_threshold);
        if (!ok) {
            if (lane == 0) valid[c] = 0;
            continue;
        }
        const int x0 = minx_[c], y0 = miny_[c];
        const int w = maxx_[c] - x0 + 1, h = maxy_[c] - y0 + 1;
        const int rt = root_[c];
        // niter = int(math.sqrt(size * min(w, h) / (w * h)) * 2): int32 product, float64 quotient and root
        const int mn = w < h ? w : h;
        const int niter = (int)(sqrt((double)(area * mn) / (double)(w * h)) * 2.0);
        int sx = x0 - niter, ex = x0 + w + niter + 1, sy = y0 - niter, ey = y0 + h + niter + 1;
        if (sx < 0) sx = 0;
        if (sy < 0) sy = 0;
        if (ex >= W) ex = W;
        if (ey >= H) ey = H;
        // segmap rows: pixels of this component that are not (link_score == 1 and text_score == 0)
        const long img0 = (long)b * H * W;
        for (int j = 0; j < h; ++j) {
            int lo = 0x7fff, hi = -1;
            const long rowbase = img0 + (long)(y0 + j) * W;
            for (int x = x0 + lane; x < x0 + w; x += 32) {
                const long i = rowbase + x;
                if (parent[i] == rt && flags[i] != 2) {
                    lo = lo < x ? lo : x;
                    hi = hi > x ? hi : x;
                }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const int l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
                lo = lo < l2 ? lo : l2;
                hi = hi > h2 ? hi : h2;
            }
            if (lane == 0) {
                rmin[j] = (short)lo;
                rmax[j] = (short)hi;
            }
        }
        __syncwarp();
        // dilation by the (1+niter)^2 rectangle, OpenCV anchor k/2: a pixel p reaches [p - lo_, p + hi_]
        const int k = 1 + niter, anchor = k / 2;
        const int lo_ = k - 1 - anchor, hi_ = anchor;
        const int Y0 = (y0 - lo_) > sy ? (y0 - lo_) : sy;
        int Y1 = y0 + h + hi_;
        if (Y1 > ey) Y1 = ey;
        const int nrows = Y1 - Y0;
        for (int r = lane; r < nrows; r += 32) {
            const int Y = Y0 + r;
            int lo = 0x7fff, hi = -1;
            for (int py = Y - hi_; py <= Y + lo_; ++py) {
                const int j = py - y0;
                if (j >= 0 && j < h && rmin[j] <= rmax[j]) {
                    const int a = rmin[j] - lo_, bb = rmax[j] + hi_;
                    lo = lo < a ? lo : a;
                    hi = hi > bb ? hi : bb;
                }
            }
            if (hi >= 0) {
                if (lo < sx) lo = sx;
                if (hi > ex - 1) hi = ex - 1;
                if (lo > hi) { lo = 0x7fff; hi = -1; }
            }
            dmin[r] = (short)lo;
            dmax[r] = (short)hi;
        }
        __syncwarp();
        if (lane == 0) {
            // --- convex hull of the row extremes (Andrew's monotone chain with rows as the sweep axis):
            //     `lo_chain` = small-x side walked top -> bottom, `up_chain` = large-x side walked bottom -> top.
            Pt* lo_chain = ring;
            Pt* up_chain = ring + kMaxRows;
            int m = 0, m2 = 0;
            int l = 0x7fff, rr = -1, t = 0x7fff, bt = -1;
            for (int r = 0; r < nrows; ++r) {
                if (dmax[r] < 0) continue;
                const int Y = Y0 + r;
                l = l < dmin[r] ? l : dmin[r];
                rr = rr > dmax[r] ? rr : dmax[r];
                t = t < Y ? t : Y;
                bt = bt > Y ? bt : Y;
                for (int e = 0; e < 2; ++e) {
                    if (e == 1 && dmax[r] == dmin[r]) break;
                    Pt p;
                    p.x = e == 0 ? dmin[r] : dmax[r];
                    p.y = (short)Y;
                    while (m >= 2 && cross_uv(lo_chain[m - 2], lo_chain[m - 1], p) <= 0) --m;
                    if (m < kMaxRows - 1) lo_chain[m++] = p;
                }
            }
            for (int r = nrows - 1; r >= 0; --r) {
                if (dmax[r] < 0) continue;
                const int Y = Y0 + r;
                for (int e = 1; e >= 0; --e) {
                    if (e == 0 && dmax[r] == dmin[r]) break;
                    Pt p;
                    p.x = e == 0 ? dmin[r] : dmax[r];
                    p.y = (short)Y;
                    while (m2 >= 2 && cross_uv(up_chain[m2 - 2], up_chain[m2 - 1], p) <= 0) --m2;
                    if (m2 < kMaxRows - 1) up_chain[m2++] = p;
                }
            }
            // ring = lo_chain[:-1] + up_chain[:-1]  (a single point yields a ring of one)
            int nring = m > 1 ? m - 1 : m;
            for (int i = 0; i + 1 < m2 && nring < kMaxRows; ++i) lo_chain[nring++] = up_chain[i];
            // cv2.convexHull(clockwise=false) walks the opposite way round and starts at the vertex with the largest
            // x (largest y among those).
            int start = 0;
            for (int i = 1; i < nring; ++i)
                if (ring[i].x > ring[start].x || (ring[i].x == ring[start].x && ring[i].y > ring[start].y)) start = i;
            Pt* cvr = ring + kMaxRows;
            const int nh = nring < 256 ? nring : 256;
            for (int i = 0; i < nh; ++i) {
                int src = start - i;
                if (src < 0) src += nring;
                cvr[i] = ring[src];
            }
            // cv2's final cyclic shift: make the original (raster) indices ascending or descending if possible
            if (nh >= 3) {
                int min_idx = 0, max_idx = 0, lt = 0;
                auto key = [&](int i) { return (int)cvr[i].y * 65536 + (int)cvr[i].x; };
                int i;
                for (i = 1; i < nh; ++i) {
                    const int idx = key(i);
                    lt += key(i - 1) < idx;
                    if (lt > 1 && lt <= i - 2) break;
                    if (idx < key(min_idx)) min_idx = i;
                    if (idx > key(max_idx)) max_idx = i;
                }
                const int mmdist = max_idx > min_idx ? max_idx - min_idx : min_idx - max_idx;
                if ((mmdist == 1 || mmdist == nh - 1) && (lt <= 1 || lt >= nh - 2)) {
                    const int ascending = (max_idx + 1) % nh == min_idx;
                    const int i0 = ascending ? min_idx : max_idx;
                    if (i0 > 0) {
                        int j = i0;
                        for (i = 0; i < nh; ++i) {
                            ring[i] = cvr[j];
                            const int next_j = j + 1 < nh ? j + 1 : 0;
                            if (i < nh - 1 && (ascending != (key(j) < key(next_j)))) break;
                            j = next_j;
                        }
                        if (i == nh)
                            for (int q = 0; q < nh; ++q) cvr[q] = ring[q];
                    }
                }
            }
            float box[8];
            min_area_box(cvr, nh, s_vx[wp], s_vy[wp], s_inv[wp], box);
            // align diamond-shape (det_utils.py:79-84), float32 like numpy
            {
                const float dx0 = box[0] - box[2], dy0 = box[1] - box[3];
                const float dx1 = box[2] - box[4], dy1 = box[3] - box[5];
                const float ew = sqrtf(dx0 * dx0 + dy0 * dy0), eh = sqrtf(dx1 * dx1 + dy1 * dy1);
                const float mx = ew > eh ? ew : eh, mnv = ew < eh ? ew : eh;
                const float ratio = mx / (mnv + 1e-5f);
                if (fabsf(1.f - ratio) <= 0.1f) {
                    box[0] = (float)l; box[1] = (float)t;
                    box[2] = (float)rr; box[3] = (float)t;
                    box[4] = (float)rr; box[5] = (float)bt;
                    box[6] = (float)l; box[7] = (float)bt;
                }
            }
            // clock-wise order starting at the corner with the smallest x + y (first minimum)
            int st = 0;
            float best = box[0] + box[1];
            for (int i = 1; i < 4; ++i) {
                const float sxy = box[2 * i] + box[2 * i + 1];
                if (sxy < best) {
                    best = sxy;
                    st = i;
                }
            }
            int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = (int)0x80000000, mxy = (int)0x80000000;
            for (int i = 0; i < 4; ++i) {
                const int src = (i + st) & 3;
                const float bx = box[2 * src], by = box[2 * src + 1];
                cbox[(size_t)c * 8 + 2 * i] = bx;
                cbox[(size_t)c * 8 + 2 * i + 1] = by;
                // adjustResultCoordinates: float64 product rounded to float32, then int32 truncation
                const int ix = (int)(float)((double)bx * scale_x);
                const int iy = (int)(float)((double)by * scale_y);
                mnx = mnx < ix ? mnx : ix;
                mxx = mxx > ix ? mxx : ix;
                mny = mny < iy ? mny : iy;
                mxy = mxy > iy ? mxy : iy;
            }
            crect[(size_t)c * 4 + 0] = mny;
            crect[(size_t)c * 4 + 1] = mnx;
            crect[(size_t)c * 4 + 2] = mxy;
            crect[(size_t)c * 4 + 3] = mxx;
            valid[c] = 1;
        }
        __syncwarp();
    }
}

// Ordered compaction of the kept components of one image (label order = the reference's `det` list order).
__global__ void __launch_bounds__(256)
pp_compact(const int* __restrict__ ncomp, const int* __restrict__ valid, const float* __restrict__ cbox,
           const int* __restrict__ crect, int max_boxes, float* __restrict__ boxes, int32_t* __restrict__ rects,
           int32_t* __restrict__ box_label, int32_t* __restrict__ counts) {
    __shared__ int wsum[8];
    __shared__ int base;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, wp = t >> 5;
    int n = ncomp[b];
    const int total = n;
    if (n >= kCap) n = kCap - 1;
    if (t == 0) base = 0;
    __syncthreads();
    for (int id0 = 1; id0 <= n; id0 += 256) {
        const int id = id0 + t;
        const int v = (id <= n) ? valid[b * kCap + id] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, v);
        const int before = __popc(m & ((1u << lane) - 1));
        if (lane == 0) wsum[wp] = __popc(m);
        __syncthreads();
        int off = base;
        for (int q = 0; q < wp; ++q) off += wsum[q];
        const int slot = off + before;
        if (v && slot < max_boxes) {
            const size_t c = (size_t)b * kCap + id;
            const size_t o = (size_t)b * max_boxes + slot;
            for (int q = 0; q < 8; ++q) boxes[o * 8 + q] = cbox[c * 8 + q];
            for (int q = 0; q < 4; ++q) rects[o * 4 + q] = crect[c * 4 + q];
            box_label[o] = id;
        }
        __syncthreads();
        if (t == 0) {
            int s = 0;
            for (int q = 0; q < 8; ++q) s += wsum[q];
            base += s;
        }
        __syncthreads();
    }
    if (t == 0) {
        counts[b * 2 + 0] = base;
        counts[b * 2 + 1] = total;
    }
}

inline int grid_for(long total, int block) {
    long g = (total + block - 1) / block;
    const long cap = 148L * 8;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

size_t postproc_workspace_bytes(int B, int H, int W) {
    const size_t npix = (size_t)B * H * W;
    const int bpi = (H * W + kScanBlock - 1) / kScanBlock;
    const size_t c = (size_t)B * kCap;
    return align256(npix * 4) * 2 + align256(npix) + align256((size_t)B * bpi * 4) + align256((size_t)B * 4) +
           8 * align256(c * 4) + align256(c * 32) + align256(c * 16) + 4096;
}

int launch_postproc(const float* score, const PostprocParams& p, void* workspace, float* boxes, int32_t* rects,
                    int32_t* box_label, int32_t* counts, int32_t* labels_out, cudaStream_t s) {
    if (p.H > kMaxRows || p.W > 32000 || p.H * p.W > 1024 * kScanBlock) return -1;
    Work w = carve(workspace, p.B, p.H, p.W);
    const long npix = (long)p.B * p.H * p.W;
    const int HW = p.H * p.W;
    const int bpi = (HW + kScanBlock - 1) / kScanBlock;
    const float2* sc = reinterpret_cast<const float2*>(score);
    {
        const dim3 tiles((p.W + kTile - 1) / kTile, (p.H + kTile - 1) / kTile, p.B);
        pp_label_tiles<<<tiles, 256, 0, s>>>(sc, p.H, p.W, p.low_text, p.link_threshold, w.parent, w.flags);
        const long seams = (long)p.B * ((long)(tiles.x - 1) * p.H + (long)(tiles.y - 1) * p.W);
        pp_merge_borders<<<grid_for(seams > 0 ? seams : 1, 256), 256, 0, s>>>(w.parent, w.flags, p.B, p.H, p.W);
    }
    pp_flatten_count<<<dim3(bpi, p.B), 256, 0, s>>>(w.parent, HW, bpi, w.blocksum);
    pp_scan<<<p.B, 1024, 0, s>>>(w.blocksum, bpi, w.ncomp);
    pp_assign<<<dim3(bpi, p.B), 1024, 0, s>>>(w.parent, HW, bpi, w.blocksum, w.rootid, w.area, w.minx, w.miny, w.maxx,
                                             w.maxy, w.maxtext, w.root);
    pp_stats<<<grid_for(npix, 256), 256, 0, s>>>(sc, w.parent, w.rootid, p.B, p.H, p.W, w.area, w.minx, w.miny, w.maxx,
                                                w.maxy, w.maxtext, labels_out);
    pp_boxes<<<dim3(148, p.B), kBoxWarps * 32, 0, s>>>(w.parent, w.flags, w.ncomp, w.area, w.minx, w.miny, w.maxx,
                                                       w.maxy, w.maxtext, w.root, p.B, p.H, p.W, p.text_threshold,
                                                       p.scale_x, p.scale_y, w.valid, w.cbox, w.crect);
    pp_compact<<<p.B, 256, 0, s>>>(w.ncomp, w.valid, w.cbox, w.crect, p.max_boxes, boxes, rects, box_label, counts);
    return 8;
}

}  // namespace locr
