// ds_mask_kernels.h — the global stage's mask preparation, included by ds_kernels.h.
//
// stitchInterStripsCustom builds two masks per strip around its warp (all file:line in /root/reference):
//   buildWarpedContentMask  src/stitch_global.cpp:353-383   gray > 3 -> float 0/1 -> warpAffine LINEAR/CONSTANT -> > 0.999
//   buildSoftBlendMask      src/stitch_global.cpp:332-351   seam AND content -> >1 -> float 0/1 -> GaussianBlur(sigma 10,
//                                                            REPLICATE) -> * binary -> * 255 -> 8U
// with the seam mask brought to the strip's size by resize(INTER_NEAREST) + threshold(> 1) (:649-655).
// Arithmetic pinned against cv2 4.13.0 (tests/test_oracle_vs_cv2.py):
//   BGR2GRAY 8U        (B*3735 + G*19235 + R*9798 + 2^14) >> 15
//   float warpAffine   of a 0/1 image on warpAffine's 5-bit coordinates: the bilinear weights (32-ax)(32-ay)/1024 ...
//                      are dyadic, every partial sum is exact, so "> 0.999" is "sum of the weights of the valid content
//                      taps >= 1023/1024" in integers
//   GaussianBlur f32   kernel n = cvRound(8 sigma + 1) | 1, separable; row pass: sum over k = 0..n-1 in that order;
//                      column pass: centre tap, then (t[-k] + t[+k]) * ky[k] for k = 1..n/2, fused multiply-add in the
//                      columns the declared oracle build filters 8 at a time (x < (W & ~7)), separately rounded in the tail
//   convertTo(8U, 255) saturate(rint(float(v) * 255.f))
#pragma once

// ---------------------------------------------------------------------------------------------
// Content mask + seam mask -> the frame's mask plane over its warped bbox.

struct MaskPrepParams {
    FrameDev F;                           // coordinates + BGRX source
    int want_content;                     // build the content mask (needs F.src resident)
    const uint8_t* seam; int seam_pitch;  // full-resolution seam mask over the bbox, or null
    const uint8_t* low; int low_pitch;    // low-resolution seam mask, resized with INTER_NEAREST through ix / iy, or null
    const int* ix; const int* iy;
    int binarize;                         // ensureBinaryMask on the seam value: > 1 -> 255, else 0
    int and_nearest;                      // AND the nearest-warped all-255 mask (read-back of the mask the blender sees)
    uint8_t* content_out;                 // content mask plane (dense), or null
    uint8_t* out;                         // seam AND content (dense)
};

// gray > 3 at source pixel (x, y); outside the image the BORDER_CONSTANT value 0
DS_D int content_tap(const FrameDev& F, int x, int y) {
    if ((unsigned)x >= (unsigned)F.src_w || (unsigned)y >= (unsigned)F.src_h) return 0;
    const uint32_t p = ld_ro(F.src + (size_t)y * F.src_pitch + x);
    const int gray = (int)(((p & 255u) * 3735u + ((p >> 8) & 255u) * 19235u + ((p >> 16) & 255u) * 9798u + 16384u) >> 15);
    return gray > 3;
}

DS_D int content_mask_value(const FrameDev& F, int u, int v) {
    const Coord c = eval_coord(F, u, v);
    const int wx1 = c.ax, wx0 = 32 - c.ax, wy1 = c.ay, wy0 = 32 - c.ay;
    int s = 0;
    if (content_tap(F, c.sx, c.sy)) s += wx0 * wy0;
    if ((wx1 * wy0) != 0 && content_tap(F, c.sx + 1, c.sy)) s += wx1 * wy0;
    if ((wx0 * wy1) != 0 && content_tap(F, c.sx, c.sy + 1)) s += wx0 * wy1;
    if ((wx1 * wy1) != 0 && content_tap(F, c.sx + 1, c.sy + 1)) s += wx1 * wy1;
    return s >= 1023 ? 255 : 0;
}

struct MaskPrepBody {
    static constexpr int PER_BLOCK = 1024;
    static int smem_bytes() { return 0; }
    template <int NT>
    DS_DM void run(const MaskPrepParams& p, int block, int tid, unsigned char*) {
        const FrameDev& F = p.F;
        const long long n = (long long)F.w * F.h;
        for (int it = tid; it < PER_BLOCK; it += NT) {
            const long long idx = (long long)block * PER_BLOCK + it;
            if (idx >= n) break;
            const int v = (int)(idx / F.w), u = (int)(idx - (long long)v * F.w);
            int m = 255;
            if (p.low) m = (int)ld_ro(p.low + (size_t)p.iy[v] * p.low_pitch + p.ix[u]);
            else if (p.seam) m = (int)p.seam[(size_t)v * p.seam_pitch + u];   // may alias `out`
            if (p.binarize) m = m > 1 ? 255 : 0;
            if (p.and_nearest) m &= eval_coord(F, u, v).m;
            if (p.want_content) {
                const int cm = content_mask_value(F, u, v);
                if (p.content_out) p.content_out[idx] = (uint8_t)cm;
                m &= cm;
            }
            p.out[idx] = (uint8_t)m;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// buildSoftBlendMask: one CTA per 64x64 tile of the mask plane. The tile's window (tile + R on every side, REPLICATE
// at the plane's border) is staged as bytes; a uniform window (the bulk of a strip: far from any seam or hole)
// short-cuts to 255 / 0 — the blurred value of an all-ones window is within a few ulp of 1 and rounds to 255.

#define DS_SOFT_MAXR 40
struct SoftMaskParams {
    const uint8_t* bin; int bin_pitch;   // seam AND content; a pixel counts as 1 when > 1 (ensureBinaryMask)
    uint8_t* out; int out_pitch;
    int w, h;
    int R;                               // kernel radius, n = 2R + 1 <= 81
    float k[2 * DS_SOFT_MAXR + 1];       // cv::getGaussianKernel(n, sigma, CV_32F)
};
struct SoftMaskBody {
    static constexpr int T = 64;
    static constexpr int WIN = T + 2 * DS_SOFT_MAXR;   // 144
    static int smem_bytes() { return WIN * WIN + WIN * T * (int)sizeof(float); }
    template <int NT>
    DS_DM void run(const SoftMaskParams& p, int block, int tid, unsigned char* smem) {
        const int tiles_x = (p.w + T - 1) / T;
        const int ty = block / tiles_x, tx = block - ty * tiles_x;
        const int x0 = tx * T, y0 = ty * T;
        const int R = p.R, win_w = T + 2 * R, win_h = T + 2 * R;
        unsigned char* in = smem;                         // [win_h][WIN]
        float* t = (float*)(smem + WIN * WIN);            // [win_h][T]
        int all0 = 1, all1 = 1;
        for (int i = tid; i < win_w * win_h; i += NT) {
            const int r = i / win_w, cidx = i - r * win_w;
            const int y = imin(imax(y0 - R + r, 0), p.h - 1), x = imin(imax(x0 - R + cidx, 0), p.w - 1);
            const int b = ld_ro(p.bin + (size_t)y * p.bin_pitch + x) > 1;
            in[r * WIN + cidx] = (unsigned char)b;
            all0 &= !b; all1 &= b;
        }
        all0 = block_and(all0);
        all1 = block_and(all1);
        if (all0 || all1) {
            const unsigned char val = all1 ? 255 : 0;
            for (int i = tid; i < T * T; i += NT) {
                const int r = i / T, cidx = i - r * T;
                if (y0 + r < p.h && x0 + cidx < p.w) p.out[(size_t)(y0 + r) * p.out_pitch + x0 + cidx] = val;
            }
            return;
        }
        // row pass: plain order over the kernel; the input is 0 / 1, so every product is exact
        for (int i = tid; i < win_h * T; i += NT) {
            const int r = i / T, cidx = i - r * T;
            const unsigned char* row = in + r * WIN + cidx;
            float s = 0.f;
            for (int j = 0; j <= 2 * R; j++)
                if (row[j]) s = f_add(s, p.k[j]);
            t[r * T + cidx] = s;
        }
        DS_SYNC();
        // column pass, * binary, * 255 -> 8U
        const int simd_w = p.w & ~7;
        for (int i = tid; i < T * T; i += NT) {
            const int r = i / T, cidx = i - r * T;
            const int x = x0 + cidx, y = y0 + r;
            if (x >= p.w || y >= p.h) continue;
            int o = 0;
            if (in[(r + R) * WIN + cidx + R]) {
                const float* tc = t + (r + R) * T + cidx;
                float s = f_mul(tc[0], p.k[R]);
                if (x < simd_w) {
                    for (int j = 1; j <= R; j++) s = f_fma(f_add(tc[-j * T], tc[j * T]), p.k[R + j], s);
                } else {
                    for (int j = 1; j <= R; j++) s = f_add(s, f_mul(f_add(tc[-j * T], tc[j * T]), p.k[R + j]));
                }
                o = sat8i(f2i_rn(f_mul(s, 255.f)));
            }
            p.out[(size_t)y * p.out_pitch + x] = (unsigned char)o;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// autoCropBlackBorder (src/stitch_common.cpp:4-27), device half: BGR2GRAY > 1 of the composited canvas, reduced to
// the runs of foreground pixels of every row. A block owns one row; run starts and ends are appended to the row's
// event list (x << 1 | is_end) through an atomic cursor — a handful per row for a mosaic. The host pairs them, labels
// the 8-connected components of the runs and picks the contour the reference picks (ds_runtime.cu: auto_crop).

struct RowRunsParams {
    const uint8_t* out; size_t out_pitch; int bpp;   // composited canvas, BGR8 / BGRA8
    int w, h;
    int cap;                                         // events per row
    int* count;                                      // [h] events appended (may exceed cap: overflow)
    int* events;                                     // [h][cap]
};
DS_D int crop_fg(const RowRunsParams& p, const uint8_t* row, int x) {
    if ((unsigned)x >= (unsigned)p.w) return 0;
    const uint8_t* q = row + (size_t)x * p.bpp;
    return (int)((q[0] * 3735u + q[1] * 19235u + q[2] * 9798u + 16384u) >> 15) > 1;
}
struct RowRunsBody {
    static int smem_bytes() { return 0; }
    template <int NT>
    DS_DM void run(const RowRunsParams& p, int block, int tid, unsigned char*) {
        const int y = block;
        const uint8_t* row = p.out + (size_t)y * p.out_pitch;
        for (int x = tid; x < p.w; x += NT) {
            if (!crop_fg(p, row, x)) continue;
            const int st = !crop_fg(p, row, x - 1), en = !crop_fg(p, row, x + 1);
            if (st) { const int k = ds_atomic_add(p.count + y, 1); if (k < p.cap) p.events[(size_t)y * p.cap + k] = x << 1; }
            if (en) { const int k = ds_atomic_add(p.count + y, 1); if (k < p.cap) p.events[(size_t)y * p.cap + k] = (x << 1) | 1; }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// BlocksGainCompensator::apply's gain-map upsizing (src/stitch_robust.cpp:209-211): cv::resize(CV_32FC1,
// INTER_LINEAR) of the compensator's block gain map to the warped bbox, as the declared OpenCV build computes it
// (oracle: orc_resize_linear_f32, pinned against cv2.resize): per-column / per-row source index and fraction tabulated on
// the host in double, h = fma(S[x1] - S[x0], a, S[x0]) on both source rows, out = fma(h1 - h0, b, h0).
struct GainResizeParams {
    const float* src; int sw, sh, spitch;      // block gain map (device copy), pitch in elements
    const int* ix; const float* ax;            // per output column
    const int* iy; const float* ay;            // per output row
    float* dst; int dw, dh;                    // per-pixel gain plane over the warped bbox
};
struct GainResizeBody {
    static constexpr int PER_BLOCK = 1024;
    static int smem_bytes() { return 0; }
    template <int NT>
    DS_DM void run(const GainResizeParams& p, int block, int tid, unsigned char*) {
        const long long n = (long long)p.dw * p.dh;
        for (int it = tid; it < PER_BLOCK; it += NT) {
            const long long idx = (long long)block * PER_BLOCK + it;
            if (idx >= n) break;
            const int y = (int)(idx / p.dw), x = (int)(idx - (long long)y * p.dw);
            const int x0 = ld_ro(p.ix + x), x1 = imin(x0 + 1, p.sw - 1);
            const int y0 = ld_ro(p.iy + y), y1 = imin(y0 + 1, p.sh - 1);
            const float a = ld_ro(p.ax + x), b = ld_ro(p.ay + y);
            const float* r0 = p.src + (size_t)y0 * p.spitch;
            const float* r1 = p.src + (size_t)y1 * p.spitch;
            const float s00 = ld_ro(r0 + x0), s01 = ld_ro(r0 + x1), s10 = ld_ro(r1 + x0), s11 = ld_ro(r1 + x1);
            const float h0 = f_fma(f_sub(s01, s00), a, s00);
            const float h1 = f_fma(f_sub(s11, s10), a, s10);
            p.dst[idx] = f_fma(f_sub(h1, h0), b, h0);
        }
    }
};
