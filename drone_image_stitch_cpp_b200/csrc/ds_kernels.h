// ds_kernels.h — kernel bodies of libdronestitch_cuda.
//
// Every body is `template <int NT> run(P, block, tid, smem)`: phases are strided loops
// `for (i = tid; i < n; i += NT)` separated by DS_SYNC(). Compiled by nvcc for sm_100a they are the
// product kernels; compiled with DS_EMU and NT = 1 they are the CPU emulation the non-GPU tests use
// to check tile / halo / border logic against the oracle (tests/emu, never shipped).
//
// Arithmetic follows SURVEY.md Appendix A (OpenCV 4.13 semantics the reference relies on):
//   A2/A3 plane backward map + INTER_BITS tables, A4 15-bit bilinear, A5 nearest mask,
//   A6/A7 warpAffine / warpPerspective coordinates, A8 pyrDown 16S, A9 pyrDown f32 (op order),
//   A10 pyrUp 16S, A11 MultiBandBlender feed / blend, A12 FeatherBlender.
#pragma once
#include "ds_types.h"
#include <type_traits>
#if !DS_CUDA
#include <stdio.h>
#include <stdlib.h>
#endif

// ---------------------------------------------------------------------------------------------
// exact helpers

// cv::borderInterpolate for REFLECT (type 1) / REFLECT_101 (type 2); p may be any int.
DS_D int refl(int p, int len, int type) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    const int delta = (type == BORDER_REFL101);
    do {
        if (p < 0) p = -p - 1 + delta;
        else p = len - 1 - (p - len) - delta;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}
DS_D int refl101(int p, int len) { return refl(p, len, BORDER_REFL101); }

// pyrUp neighbour rules (A10): index -1 -> 1 (or 0 when n == 1); index n -> n-1.
DS_D int up_l(int i, int n) { return i > 0 ? i - 1 : (n > 1 ? 1 : 0); }
DS_D int up_r(int i, int n) { return i < n - 1 ? i + 1 : n - 1; }

// Source coordinate of bbox pixel (u, v) in [0,w) x [0,h) — fixed-point tables + nearest mask.
DS_D Coord eval_coord(const FrameDev& F, int u, int v) {
    Coord c;
    if (F.kind == XF_PLANE) {
        float U = (float)(F.tlx + u), V = (float)(F.tly + v);
        if (F.scale != 1.f) { U = f_div(U, F.scale); V = f_div(V, F.scale); }
        const float up = f_sub(U, F.t0), vp = f_sub(V, F.t1);
        float x = f_add(f_add(f_mul(F.k[0], up), f_mul(F.k[1], vp)), F.k2one);
        float y = f_add(f_add(f_mul(F.k[3], up), f_mul(F.k[4], vp)), F.k5one);
        const float z = f_add(f_add(f_mul(F.k[6], up), f_mul(F.k[7], vp)), F.k8one);
        if (z != 1.f) { x = f_div(x, z); y = f_div(y, z); }
        const int ix = f2i_rn(f_mul(x, 32.f)), iy = f2i_rn(f_mul(y, 32.f));
        c.sx = sat16i(ix >> 5); c.sy = sat16i(iy >> 5);
        c.ax = ix & 31; c.ay = iy & 31;
        const int nx = sat16i(f2i_rn(x)), ny = sat16i(f2i_rn(y));
        c.m = ((unsigned)nx < (unsigned)F.src_w && (unsigned)ny < (unsigned)F.src_h) ? 255 : 0;
    } else if (F.kind == XF_AFFINE) {
        const double m0 = F.inv[0], m1 = F.inv[1], b1 = F.inv[2], m3 = F.inv[3], m4 = F.inv[4], b2 = F.inv[5];
        const int ad = d2i_rn(d_mul(d_mul(m0, (double)u), 1024.0));
        const int bd = d2i_rn(d_mul(d_mul(m3, (double)u), 1024.0));
        const int xr = d2i_rn(d_mul(d_add(d_mul(m1, (double)v), b1), 1024.0));
        const int yr = d2i_rn(d_mul(d_add(d_mul(m4, (double)v), b2), 1024.0));
        const int X = (xr + 16 + ad) >> 5, Y = (yr + 16 + bd) >> 5;
        c.sx = sat16i(X >> 5); c.sy = sat16i(Y >> 5);
        c.ax = X & 31; c.ay = Y & 31;
        const int nx = sat16i((xr + 512 + ad) >> 10), ny = sat16i((yr + 512 + bd) >> 10);
        c.m = ((unsigned)nx < (unsigned)F.src_w && (unsigned)ny < (unsigned)F.src_h) ? 255 : 0;
    } else {
        const double* I = F.inv;
        const double X0 = d_add(d_add(d_mul(I[0], (double)u), d_mul(I[1], (double)v)), I[2]);
        const double Y0 = d_add(d_add(d_mul(I[3], (double)u), d_mul(I[4], (double)v)), I[5]);
        double Wd = d_add(d_add(d_mul(I[6], (double)u), d_mul(I[7], (double)v)), I[8]);
        const double W32 = Wd != 0.0 ? d_div(32.0, Wd) : 0.0;
        const double W1 = Wd != 0.0 ? d_div(1.0, Wd) : 0.0;
        const double lo = -2147483648.0, hi = 2147483647.0;
        double fx = d_mul(X0, W32), fy = d_mul(Y0, W32);
        fx = fx < lo ? lo : (fx > hi ? hi : fx);
        fy = fy < lo ? lo : (fy > hi ? hi : fy);
        const int X = d2i_rn(fx), Y = d2i_rn(fy);
        c.sx = sat16i(X >> 5); c.sy = sat16i(Y >> 5);
        c.ax = X & 31; c.ay = Y & 31;
        double gx = d_mul(X0, W1), gy = d_mul(Y0, W1);
        gx = gx < lo ? lo : (gx > hi ? hi : gx);
        gy = gy < lo ? lo : (gy > hi ? hi : gy);
        const int nx = sat16i(d2i_rn(gx)), ny = sat16i(d2i_rn(gy));
        c.m = ((unsigned)nx < (unsigned)F.src_w && (unsigned)ny < (unsigned)F.src_h) ? 255 : 0;
    }
    return c;
}

DS_D uint32_t src_tap(const FrameDev& F, int x, int y) {
    // x, y already border-resolved; negative = outside with BORDER_CONSTANT -> 0
    if ((x | y) < 0) return 0u;
    return ld_ro(F.src + (size_t)y * F.src_pitch + x);
}

// Radiometric gains applied to a warped pixel, in the reference's order: per-strip channel gain (float32),
// exposure-compensator scalar gains (float64), block gain map (float32 per pixel at bbox position ur, vr).
DS_D void apply_gains(const FrameDev& F, int& b, int& g, int& r, int ur, int vr) {
    if (F.has_gain) {
        b = sat8i(f2i_rn(f_mul((float)b, F.gain[0]))); g = sat8i(f2i_rn(f_mul((float)g, F.gain[1]))); r = sat8i(f2i_rn(f_mul((float)r, F.gain[2])));
    }
    if (F.has_cgain) {
        b = sat8i(d2i_rn(d_mul((double)b, F.cgain[0]))); g = sat8i(d2i_rn(d_mul((double)g, F.cgain[1]))); r = sat8i(d2i_rn(d_mul((double)r, F.cgain[2])));
    }
    if (F.gainmap) {
        const float gm = ld_ro(F.gainmap + (size_t)vr * F.gainmap_pitch + ur);
        b = sat8i(f2i_rn(f_mul((float)b, gm))); g = sat8i(f2i_rn(f_mul((float)g, gm))); r = sat8i(f2i_rn(f_mul((float)r, gm)));
    }
}
// the same with the gain-map value already loaded (gm is ignored without a gain map)
DS_D void apply_gains_with(const FrameDev& F, int& b, int& g, int& r, float gm) {
    if (F.has_gain) {
        b = sat8i(f2i_rn(f_mul((float)b, F.gain[0]))); g = sat8i(f2i_rn(f_mul((float)g, F.gain[1]))); r = sat8i(f2i_rn(f_mul((float)r, F.gain[2])));
    }
    if (F.has_cgain) {
        b = sat8i(d2i_rn(d_mul((double)b, F.cgain[0]))); g = sat8i(d2i_rn(d_mul((double)g, F.cgain[1]))); r = sat8i(d2i_rn(d_mul((double)r, F.cgain[2])));
    }
    if (F.gainmap) {
        b = sat8i(f2i_rn(f_mul((float)b, gm))); g = sat8i(f2i_rn(f_mul((float)g, gm))); r = sat8i(f2i_rn(f_mul((float)r, gm)));
    }
}

// A4: cv::remap INTER_LINEAR 8UC3 on the fixed-point coordinate of bbox pixel (ur, vr), then the gains.
DS_D px8 sample_bilinear(const FrameDev& F, const Coord& c, int ur, int vr) {
    uint32_t p00, p01, p10, p11;
    if ((unsigned)c.sx < (unsigned)(F.src_w - 1) && (unsigned)c.sy < (unsigned)(F.src_h - 1)) {
        const uint32_t* r0 = F.src + (size_t)c.sy * F.src_pitch + c.sx;
        p00 = ld_ro(r0); p01 = ld_ro(r0 + 1);
        p10 = ld_ro(r0 + F.src_pitch); p11 = ld_ro(r0 + F.src_pitch + 1);
    } else if (F.border == BORDER_CONST) {
        const int x0 = (unsigned)c.sx < (unsigned)F.src_w ? c.sx : -1;
        const int x1 = (unsigned)(c.sx + 1) < (unsigned)F.src_w ? c.sx + 1 : -1;
        const int y0 = (unsigned)c.sy < (unsigned)F.src_h ? c.sy : -1;
        const int y1 = (unsigned)(c.sy + 1) < (unsigned)F.src_h ? c.sy + 1 : -1;
        p00 = src_tap(F, x0, y0); p01 = src_tap(F, x1, y0);
        p10 = src_tap(F, x0, y1); p11 = src_tap(F, x1, y1);
    } else {
        const int x0 = refl(c.sx, F.src_w, BORDER_REFL), x1 = refl(c.sx + 1, F.src_w, BORDER_REFL);
        const int y0 = refl(c.sy, F.src_h, BORDER_REFL), y1 = refl(c.sy + 1, F.src_h, BORDER_REFL);
        p00 = src_tap(F, x0, y0); p01 = src_tap(F, x1, y0);
        p10 = src_tap(F, x0, y1); p11 = src_tap(F, x1, y1);
    }
    // weights (32-ax)(32-ay)*32 ... sum to 2^15; (sum + 2^14) >> 15 == (hv + 512) >> 10
    const int wx1 = c.ax, wx0 = 32 - c.ax, wy1 = c.ay, wy0 = 32 - c.ay;
    int out[3];
    DS_UNROLL
    for (int ch = 0; ch < 3; ch++) {
        const int sh = 8 * ch;
        const int a = (int)((p00 >> sh) & 255u), b = (int)((p01 >> sh) & 255u);
        const int d = (int)((p10 >> sh) & 255u), e = (int)((p11 >> sh) & 255u);
        const int hv = (a * wx0 + b * wx1) * wy0 + (d * wx0 + e * wx1) * wy1;
        out[ch] = (hv + 512) >> 10;
    }
    if (F.any_gain) apply_gains(F, out[0], out[1], out[2], ur, vr);
    px8 r;
    r.b = (unsigned char)out[0]; r.g = (unsigned char)out[1]; r.r = (unsigned char)out[2]; r.a = 0;
    return r;
}

// Warped mask value of bbox pixel (u, v): nearest-warped 255s ANDed with the optional seam mask.
DS_D int mask_value(const FrameDev& F, const Coord& c, int u, int v) {
    int m = c.m;
    if (F.seam) m &= (int)ld_ro(F.seam + (size_t)v * F.seam_pitch + u);
    return m;
}

// A9: float pyrDown taps in the op order of the declared oracle build.
DS_D float pd_scalar(float s0, float s1, float s2, float s3, float s4) {
    // s2*6 + (s1+s3)*4 + s0 + s4, left to right
    return f_add(f_add(f_add(f_mul(s2, 6.f), f_mul(f_add(s1, s3), 4.f)), s0), s4);
}
DS_D float pd_h_simd(float r0, float r1, float r2, float r3, float r4) {
    // r2*6 + ((r1+r3)*4 + (r0+r4))
    return f_add(f_mul(r2, 6.f), f_add(f_mul(f_add(r1, r3), 4.f), f_add(r0, r4)));
}
DS_D float pd_v_simd(float r0, float r1, float r2, float r3, float r4) {
    // ((r1+r3)+r2)*4 + ((r0+r4)+(r2+r2))
    return f_add(f_mul(f_add(f_add(r1, r3), r2), 4.f), f_add(f_add(r0, r4), f_add(r2, r2)));
}
// Horizontal class of output column j for a source row of width w (dw = (w+1)/2).
DS_D bool pd_h_is_simd(int j, int w, int dw) {
    int width0 = (w - 3) / 2 + 1;  // C division (truncates toward zero), as cv::pyrDown_
    if (width0 > dw) width0 = dw;
    const int n = width0 - 1;       // columns handed to the vector loop, starting at x = 1
    return j >= 1 && j < 1 + (n > 0 ? (n & ~3) : 0);
}
DS_D bool pd_v_is_simd(int j, int dw) { return j < (dw & ~3); }

#if !DS_CUDA
// emulation only: how many tile-frames took which level-0 warp path (DS_EMU_STATS=1 prints them at exit)
static long long g_emu_paths[6];
static void ds_emu_report() { fprintf(stderr, "[ds emu] level-0 tile-frames: interior %lld, gap %lld, edge %lld, general-fast %lld, per-pixel %lld; taps from the source box %lld\n", g_emu_paths[0], g_emu_paths[1], g_emu_paths[2], g_emu_paths[3], g_emu_paths[4], g_emu_paths[5]); }
static inline void ds_emu_count(int k) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("DS_EMU_STATS"); on = (e && atoi(e) > 0) ? 1 : 0; if (on) atexit(ds_emu_report); }
    if (on) { _Pragma("omp atomic") g_emu_paths[k]++; }
}
#endif
constexpr int ds_al16(int v) { return (v + 15) & ~15; }  // every smem section starts 16-B aligned
constexpr int ds_al128(int v) { return (v + 127) & ~127; }  // TMA destinations need 128 B

// ---------------------------------------------------------------------------------------------
// generic launch plumbing

#if DS_CUDA
// One named __global__ per body (so profiles show ds_mb_feed_l0 etc. instead of one template name).
template <class Body, int NT> struct KernelOf;
// Programmatic dependent launch: every kernel first lets the next launch of its stream start being scheduled (its CTAs
// become resident as this grid's last wave drains) and then waits until the previous grid has completed and its memory
// is visible. Nothing is read or written before the wait, so the only overlap is launch latency and the wave tail.
__device__ __forceinline__ void ds_grid_dependency_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#define DS_DEFINE_KERNEL(kname, Body, NT, P, MINB)                                         \
    __global__ void __launch_bounds__(NT, MINB) kname(const P p) {                         \
        extern __shared__ __align__(128) unsigned char ds_smem[];                           \
        ds_grid_dependency_sync();                                                          \
        Body::template run<NT>(p, (int)blockIdx.x, (int)threadIdx.x, ds_smem);             \
    }                                                                                      \
    template <> struct KernelOf<Body, NT> { static constexpr void (*fn)(const P) = kname; };
#endif

// ---------------------------------------------------------------------------------------------
// BGR (3 B/px, arbitrary stride) -> BGRX (4 B/px) expansion done once at upload.

struct ExpandParams {
    const uint8_t* src; size_t src_stride;
    uint32_t* dst; int dst_pitch;
    int w, h;
};
struct ExpandBody {
    static constexpr int PER_BLOCK = 1024;   // items per block; one item = 4 consecutive pixels of a row
    static int smem_bytes() { return 0; }
    static long long blocks(const ExpandParams& p) { return ((long long)((p.w + 3) / 4) * p.h + PER_BLOCK - 1) / PER_BLOCK; }
    template <int NT>
    DS_DM void run(const ExpandParams& p, int block, int tid, unsigned char*) {
        const int qw = (p.w + 3) / 4;
        const long long n = (long long)qw * p.h;
        for (int i = tid; i < PER_BLOCK; i += NT) {
            const long long idx = (long long)block * PER_BLOCK + i;
            if (idx >= n) break;
            int y, xq;   // 32-bit arithmetic when the frame allows it: a 64-bit division would dominate this copy kernel
            if (n <= 0x7fffffffLL) { const unsigned u = (unsigned)idx; y = (int)(u / (unsigned)qw); xq = (int)(u - (unsigned)y * (unsigned)qw); }
            else { y = (int)(idx / qw); xq = (int)(idx - (long long)y * qw); }
            const int x = xq * 4;
            const uint8_t* s = p.src + (size_t)y * p.src_stride + (size_t)x * 3;
            uint32_t* d = p.dst + (size_t)y * p.dst_pitch + x;   // dst_pitch is a multiple of 4 pixels: 16-byte aligned
            if (x + 4 <= p.w && (((size_t)s) & 3) == 0) {
                // 12 source bytes as three words -> four BGRX words, one 128-bit store
                const uint32_t w0 = ld_ro((const uint32_t*)s), w1 = ld_ro((const uint32_t*)s + 1), w2 = ld_ro((const uint32_t*)s + 2);
                *(uint4*)d = make_u4(w0 & 0xffffffu, (w0 >> 24) | ((w1 & 0xffffu) << 8), (w1 >> 16) | ((w2 & 0xffu) << 16), w2 >> 8);
            } else {
                for (int k = 0; k < 4 && x + k < p.w; k++)
                    d[k] = (uint32_t)s[3 * k] | ((uint32_t)s[3 * k + 1] << 8) | ((uint32_t)s[3 * k + 2] << 16);
            }
        }
    }
};

// NVLink P2P halo pull: every segment is a run of whole rows of a straddling frame's level-1 Gaussian or weight
// level in the neighbouring handle's memory (peer-mapped), copied to the same rows of the local arrays. Loads
// bypass the local caches (the data was written by another GPU).
struct PullSeg { const uint4* src; uint4* dst; long long n16; };
struct PullParams { const PullSeg* segs; int nsegs; };
struct PullBody {
    static constexpr int BLOCKS_PER_SEG = 48;
    static int smem_bytes() { return 0; }
    template <int NT>
    DS_DM void run(const PullParams& p, int block, int tid, unsigned char*) {
        const int si = block / BLOCKS_PER_SEG, part = block - si * BLOCKS_PER_SEG;
        if (si >= p.nsegs) return;
        const PullSeg sg = p.segs[si];
        // four independent 16-byte peer loads in flight per thread: an NVLink round trip is a few microseconds
        const long long stride = (long long)BLOCKS_PER_SEG * NT;
        long long i = (long long)part * NT + tid;
        for (; i + 3 * stride < sg.n16; i += 4 * stride) {
            const uint4 v0 = ld_peer(sg.src + i), v1 = ld_peer(sg.src + i + stride), v2 = ld_peer(sg.src + i + 2 * stride), v3 = ld_peer(sg.src + i + 3 * stride);
            sg.dst[i] = v0; sg.dst[i + stride] = v1; sg.dst[i + 2 * stride] = v2; sg.dst[i + 3 * stride] = v3;
        }
        for (; i < sg.n16; i += stride) sg.dst[i] = ld_peer(sg.src + i);
    }
};

// Hand-over counters between neighbouring handles: one thread publishes `value` into up to two words of the
// neighbours' memory after making this stream's earlier writes visible system-wide.
struct SignalParams { int* a; int* b; int value; };
struct SignalBody {
    static int smem_bytes() { return 0; }
    template <int NT>
    DS_DM void run(const SignalParams& p, int block, int tid, unsigned char*) {
        if (block != 0 || tid != 0) return;
        fence_system();
        if (p.a) st_flag(p.a, p.value);
        if (p.b) st_flag(p.b, p.value);
    }
};

// Copies the host-built launch metadata (tile lists, frame descriptors, tensor maps) from mapped pinned host
// memory into its device arena: a kernel instead of a DMA so that it never queues behind frame uploads.
struct MetaCopyParams { const uint4* src; uint4* dst; long long n16; };
struct MetaCopyBody {
    static constexpr int PER_BLOCK = 1024;
    static int smem_bytes() { return 0; }
    template <int NT>
    DS_DM void run(const MetaCopyParams& p, int block, int tid, unsigned char*) {
        for (int i = tid; i < PER_BLOCK; i += NT) {
            const long long idx = (long long)block * PER_BLOCK + i;
            if (idx >= p.n16) break;
            p.dst[idx] = p.src[idx];
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Seam-mask upsizing of composePanorama (A13): dilate 3x3 of the low-res seam mask, then
// resize(INTER_LINEAR_EXACT) to the warped bbox. The 8.8 fixed-point coefficients are tabulated on the
// host in double (as OpenCV does, softdouble); the kernel does the dilation on the fly.

struct SeamUpParams {
    const uint8_t* src; int sw, sh, sstride;   // low-res mask (device copy)
    const int* ix; const int* cx;              // per output column: left sample, c1 (0..256)
    const int* iy; const int* cy;              // per output row
    uint8_t* dst; int dw, dh;                  // full-res seam mask (bbox size)
};
struct SeamUpBody {
    static constexpr int PER_BLOCK = 1024;
    static int smem_bytes() { return 0; }
    DS_DM int dil(const SeamUpParams& p, int x, int y) {
        int m = 0;
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                const int yy = y + dy, xx = x + dx;
                if ((unsigned)yy < (unsigned)p.sh && (unsigned)xx < (unsigned)p.sw) m = imax(m, (int)ld_ro(p.src + (size_t)yy * p.sstride + xx));
            }
        return m;
    }
    template <int NT>
    DS_DM void run(const SeamUpParams& p, int block, int tid, unsigned char*) {
        const long long n = (long long)p.dw * p.dh;
        for (int it = tid; it < PER_BLOCK; it += NT) {
            const long long idx = (long long)block * PER_BLOCK + it;
            if (idx >= n) break;
            const int y = (int)(idx / p.dw), x = (int)(idx - (long long)y * p.dw);
            const int x0 = p.ix[x], x1 = imin(x0 + 1, p.sw - 1), c1 = p.cx[x];
            const int y0 = p.iy[y], y1 = imin(y0 + 1, p.sh - 1), d1 = p.cy[y];
            const int h0 = (256 - c1) * dil(p, x0, y0) + c1 * dil(p, x1, y0);
            const int h1 = (256 - c1) * dil(p, x0, y1) + c1 * dil(p, x1, y1);
            p.dst[idx] = (uint8_t)(((256 - d1) * h0 + d1 * h1 + 32768) >> 16);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Debug taps: fixed-point tables and the warped bbox of one frame.

struct TapParams {
    const FrameDev* frames; int frame;
    int16_t* xy; uint16_t* a;      // tables (either both or none)
    uint8_t* bgr; uint8_t* mask;   // warped image / mask (either both or none)
};
struct TapBody {
    static constexpr int PER_BLOCK = 1024;
    static int smem_bytes() { return 0; }
    template <int NT>
    DS_DM void run(const TapParams& p, int block, int tid, unsigned char*) {
        const FrameDev& F = p.frames[p.frame];
        const long long n = (long long)F.w * F.h;
        for (int i = tid; i < PER_BLOCK; i += NT) {
            const long long idx = (long long)block * PER_BLOCK + i;
            if (idx >= n) break;
            const int v = (int)(idx / F.w), u = (int)(idx - (long long)v * F.w);
            const Coord c = eval_coord(F, u, v);
            if (p.xy) {
                p.xy[2 * idx] = (int16_t)c.sx; p.xy[2 * idx + 1] = (int16_t)c.sy;
                p.a[idx] = (uint16_t)(c.ay * 32 + c.ax);
            }
            if (p.bgr) {
                const px8 s = sample_bilinear(F, c, u, v);
                p.bgr[3 * idx] = s.b; p.bgr[3 * idx + 1] = s.g; p.bgr[3 * idx + 2] = s.r;
                p.mask[idx] = (uint8_t)mask_value(F, c, u, v);
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// FEATHER step 1: nearest-warped mask of a frame's bbox as a bit plane (1 = non-zero mask).
// One warp packs 32 pixels with a ballot; tail bits (u >= w) are set so they never act as zeros.

struct MaskBitsParams {
    const FrameDev* frames; int frame;
    uint32_t* bits;  // == frames[frame].mbits (non-const view)
};
struct MaskBitsBody {
    static constexpr int WORDS_PER_BLOCK = 8;  // 256 threads = 8 warps = 8 words
    static int smem_bytes() { return 0; }
    template <int NT>
    DS_DM void run(const MaskBitsParams& p, int block, int tid, unsigned char*) {
        const FrameDev& F = p.frames[p.frame];
        const long long nwords = (long long)F.mbits_pitch * F.h;
#if DS_CUDA
        const long long word = (long long)block * WORDS_PER_BLOCK + (tid >> 5);
        const int lane = tid & 31;
        bool bit = true;
        int v = 0, k = 0;
        if (word < nwords) {
            v = (int)(word / F.mbits_pitch); k = (int)(word - (long long)v * F.mbits_pitch);
            const int u = k * 32 + lane;
            if (u < F.w) {
                const Coord c = eval_coord(F, u, v);
                bit = mask_value(F, c, u, v) != 0;
            }
        }
        const unsigned b = __ballot_sync(0xffffffffu, bit);
        if (word < nwords && lane == 0) p.bits[word] = b;
#else
        (void)tid;
        for (int wi = 0; wi < WORDS_PER_BLOCK; wi++) {
            const long long word = (long long)block * WORDS_PER_BLOCK + wi;
            if (word >= nwords) break;
            const int v = (int)(word / F.mbits_pitch), k = (int)(word - (long long)v * F.mbits_pitch);
            uint32_t b = 0;
            for (int lane = 0; lane < 32; lane++) {
                const int u = k * 32 + lane;
                bool bit = true;
                if (u < F.w) {
                    const Coord c = eval_coord(F, u, v);
                    bit = mask_value(F, c, u, v) != 0;
                }
                b |= (bit ? 1u : 0u) << lane;
            }
            p.bits[word] = b;
        }
#endif
    }
};

// The same bit plane for plane-mapped frames without a seam mask, one THREAD per 32-pixel word: along a bbox row the
// backward map is monotone in each source coordinate (also in float arithmetic: a chain of monotone roundings), so when
// both end pixels of the word map into the source every pixel between them does - two coordinate evaluations settle
// 32 pixels. Only the words that straddle the edge of the warped source evaluate every pixel.
struct MaskBitsRowParams {
    const FrameDev* frames; int frame;
    uint32_t* bits;
};
struct MaskBitsRowBody {
    static constexpr int PER_BLOCK = 256;
    static int smem_bytes() { return 0; }
    static bool eligible(const FrameDev& F) {   // host side
        return F.kind == XF_PLANE && !F.seam && F.k[6] == 0.f && F.k[7] == 0.f && F.k8one == 1.f;
    }
    template <int NT>
    DS_DM void run(const MaskBitsRowParams& p, int block, int tid, unsigned char*) {
        const FrameDev& F = p.frames[p.frame];
        const long long nwords = (long long)F.mbits_pitch * F.h;
        for (int it = tid; it < PER_BLOCK; it += NT) {
            const long long word = (long long)block * PER_BLOCK + it;
            if (word >= nwords) break;
            const int v = (int)(word / F.mbits_pitch), k = (int)(word - (long long)v * F.mbits_pitch);
            const int u0 = k * 32, u1 = imin(u0 + 31, F.w - 1);
            // out codes of the two end pixels: left / right / above / below the source (nearest-neighbour coordinates)
            auto code = [&](int u) {
                float U = (float)(F.tlx + u), V = (float)(F.tly + v);
                if (F.scale != 1.f) { U = f_div(U, F.scale); V = f_div(V, F.scale); }
                const float up = f_sub(U, F.t0), vp = f_sub(V, F.t1);
                const float x = f_add(f_add(f_mul(F.k[0], up), f_mul(F.k[1], vp)), F.k2one);
                const float y = f_add(f_add(f_mul(F.k[3], up), f_mul(F.k[4], vp)), F.k5one);
                const int nx = sat16i(f2i_rn(x)), ny = sat16i(f2i_rn(y));   // as eval_coord
                return (nx < 0 ? 1 : 0) | (nx >= F.src_w ? 2 : 0) | (ny < 0 ? 4 : 0) | (ny >= F.src_h ? 8 : 0);
            };
            const int c0 = code(u0), c1 = code(u1);
            uint32_t b = 0xffffffffu;   // tail bits (u >= w) stay set: they never act as zeros
            if (c0 & c1) {
                // both ends beyond the same side of the source: so is everything between them
                b = (u1 - u0 == 31) ? 0u : ~((1u << (u1 - u0 + 1)) - 1u);
            } else if (c0 | c1) {
                for (int u = u0; u <= u1; u++)
                    if (code(u) != 0) b &= ~(1u << (u - u0));
            }
            p.bits[word] = b;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// FEATHER step 1b: L1 distance of every bbox pixel to the nearest zero of the frame's warped mask, capped at 255 and
// exact below the feather radius R (R * sharpness >= 1, so farther pixels have weight 1 anyway) - cv::distanceTransform
// (DIST_L1, 3x3) as createWeightMap uses it (A12): zeros are sources only inside the image. Once per frame instead of
// once per canvas tile. One block = a strip of 256 columns x 96 rows:
//   (0) the bit window of the strip +- R is staged in shared memory; a window without a zero writes 255s and is done;
//   (a) horizontal distance of every window pixel from the row's bits alone: the nearest zero bit at or left of it
//       (count leading zeros of the inverted 64-bit window ending at the pixel) and at or right of it (trailing zeros);
//   (b) separable min-plus: one thread per column sweeps down and up over the window rows.
struct FeatherDistParams {
    const FrameDev* frames; int frame;
    uint8_t* dist;    // == frames[frame].dist (non-const view)
    int R;
};
struct FeatherDistBody {
    static constexpr int SW = 256, SH = 96, RMAX = 64;
    static constexpr int WIN_ROWS = SH + 2 * RMAX;
    static constexpr int WORDS = SW / 32 + 2 + 2 + 1;   // bit words per window row: two words (64 px >= RMAX) either side + spill
    static constexpr int H_BYTES = WIN_ROWS * SW, B_BYTES = WIN_ROWS * WORDS * 4;
    static int smem_bytes() { return H_BYTES + B_BYTES + 16; }
    static long long blocks(const FrameDev& F) { return (long long)((F.w + SW - 1) / SW) * ((F.h + SH - 1) / SH); }
    template <int NT>
    DS_DM void run(const FeatherDistParams& p, int block, int tid, unsigned char* smem) {
        const FrameDev& F = p.frames[p.frame];
        unsigned char* s_h = smem;
        uint32_t* s_b = (uint32_t*)(smem + H_BYTES);
        int* s_flag = (int*)(smem + H_BYTES + B_BYTES);
        const int nbx = (F.w + SW - 1) / SW;
        const int by = block / nbx, bx = block - by * nbx;
        const int x0 = bx * SW, y0 = by * SH, R = p.R;
        const int wy0 = y0 - R, wrows = SH + 2 * R;
        const int k0 = (x0 >> 5) - 2;                      // first bit word of the window rows (may be negative)
        if (tid == 0) *s_flag = 0;
        DS_SYNC();
        // (0) stage the bits: outside the bbox there are no zeros
        int found = 0;
        for (int i = tid; i < wrows * WORDS; i += NT) {
            const int r = i / WORDS, kk = i - r * WORDS;
            const int v = wy0 + r, k = k0 + kk;
            uint32_t wd = 0xffffffffu;
            if ((unsigned)v < (unsigned)F.h && (unsigned)k < (unsigned)F.mbits_pitch) wd = ld_ro(F.mbits + (size_t)v * F.mbits_pitch + k);
            s_b[i] = wd;
            if (wd != 0xffffffffu) found = 1;
        }
        if (found) *s_flag = 1;   // benign race: all writers store 1
        DS_SYNC();
        const int ow = imin(SW, F.w - x0), oh = imin(SH, F.h - y0);
        if (!*s_flag) {
            for (int i = tid; i < oh * (SW / 4); i += NT) {
                const int r = i / (SW / 4), c4 = (i - r * (SW / 4)) * 4;
                if (c4 >= ow) continue;
                uint8_t* q = p.dist + (size_t)(y0 + r) * F.dist_pitch + x0 + c4;   // dist_pitch and x0 are multiples of 4
                *(uint32_t*)q = 0xffffffffu;
            }
            return;
        }
        // (a) horizontal distances. Pixel u sits in word (u >> 5) - k0 of its staged row, bit u & 31.
        for (int i = tid; i < wrows * SW; i += NT) {
            const int r = i / SW, c = i - r * SW;
            const int u = x0 + c;
            const uint32_t* row = s_b + r * WORDS;
            const int kk = (u >> 5) - k0, b = u & 31;   // kk >= 2
            const uint32_t wm2 = row[kk - 2], wm1 = row[kk - 1], w0 = row[kk], wp1 = row[kk + 1], wp2 = row[kk + 2];
            // pixels u-63 .. u as a 64-bit window with pixel u in bit 63; u .. u+63 with pixel u in bit 0
            const uint32_t l_lo = ~funnel_r(wm2, wm1, b + 1), l_hi = ~funnel_r(wm1, w0, b + 1);
            const uint32_t r_lo = ~funnel_r(w0, wp1, b), r_hi = ~funnel_r(wp1, wp2, b);
            int dl = l_hi ? clz32(l_hi) : (l_lo ? 32 + clz32(l_lo) : 255);
            int dr = r_lo ? ctz32(r_lo) : (r_hi ? 32 + ctz32(r_hi) : 255);
            s_h[i] = (unsigned char)imin(dl, dr);
        }
        DS_SYNC();
        // (b) vertical min-plus sweeps, one column per thread
        // (eight rows per trip: the loads of a trip are issued before its stores, so only the two-instruction min-plus
        // recurrence is serial)
        for (int c = tid; c < SW; c += NT) {
            int d = 255;
            for (int r0 = 0; r0 < wrows; r0 += 8) {
                int v[8];
                DS_UNROLL
                for (int j = 0; j < 8; j++) v[j] = (int)s_h[imin(r0 + j, wrows - 1) * SW + c];
                DS_UNROLL
                for (int j = 0; j < 8; j++) { d = imin(imin(d + 1, 255), v[j]); v[j] = d; }
                DS_UNROLL
                for (int j = 0; j < 8; j++) if (r0 + j < wrows) s_h[(r0 + j) * SW + c] = (unsigned char)v[j];
            }
            d = 255;
            for (int r0 = wrows - 1; r0 >= 0; r0 -= 8) {
                int v[8];
                DS_UNROLL
                for (int j = 0; j < 8; j++) v[j] = (int)s_h[imax(r0 - j, 0) * SW + c];
                DS_UNROLL
                for (int j = 0; j < 8; j++) { d = imin(imin(d + 1, 255), v[j]); v[j] = d; }
                DS_UNROLL
                for (int j = 0; j < 8; j++) if (r0 - j >= 0) s_h[(r0 - j) * SW + c] = (unsigned char)v[j];
            }
        }
        DS_SYNC();
        for (int i = tid; i < oh * (SW / 4); i += NT) {
            const int r = i / (SW / 4), c4 = (i - r * (SW / 4)) * 4;
            if (c4 >= ow) continue;
            const unsigned char* hrow = s_h + (r + R) * SW + c4;
            uint8_t* q = p.dist + (size_t)(y0 + r) * F.dist_pitch + x0 + c4;
            *(uint32_t*)q = (uint32_t)hrow[0] | ((uint32_t)hrow[1] << 8) | ((uint32_t)hrow[2] << 16) | ((uint32_t)hrow[3] << 24);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// FEATHER step 2: one canvas tile gathers every frame that covers it, in feed order:
//   w = min(L1dist(mask) * sharpness, 1); acc += trunc(pix * w); wsum += w
// then out = trunc(acc / (wsum + 1e-5)), mask = wsum > 1e-5 (A12). The L1 distance is exact for
// d < R (R*sharpness >= 1) and only evaluated when the tile's R-window holds a zero mask pixel.

struct OutParams {
    uint8_t* out; size_t out_pitch;    // BGR8 or BGRA8 rows
    uint8_t* mask; size_t mask_pitch;  // BGR8 only (BGRA8 carries it in alpha)
    int fmt;                            // 0 BGR8, 1 BGRA8
    int w, h;                           // unpadded canvas
};

struct FeatherParams {
    const FrameDev* frames;
    const int* tile_off; const int* tile_frames;
    const int* tile_ids;  // tiles this launch processes (NULL: block == tile)
    int tiles_x;
    float sharpness; int R;
    int row0, row1;       // canvas rows this handle's band owns
    OutParams o;
};

struct FeatherBody {
    static constexpr int TW = 64, TH = 32, RMAX = 64;
    static int smem_bytes() { return (TW + TH) * 8 + TW * TH * 4; }
    template <int NT>
    DS_DM void run(const FeatherParams& p, int block, int tid, unsigned char* smem) {
        constexpr int PPT = TW * TH / NT;
        const int tile = p.tile_ids ? p.tile_ids[block] : block;
        float2* s_cx = (float2*)smem;               // per tile column: k0*u', k3*u'
        float2* s_ry = s_cx + TW;                   // per tile row:    k1*v', k4*v'
        const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
        const int X0 = tx * TW, Y0 = ty * TH;
        const int R = p.R;
        int acc[PPT][3];
        float ws[PPT];
        DS_UNROLL
        for (int k = 0; k < PPT; k++) { acc[k][0] = acc[k][1] = acc[k][2] = 0; ws[k] = 0.f; }

        for (int fi = p.tile_off[tile]; fi < p.tile_off[tile + 1]; fi++) {
            const FrameDev& F = p.frames[p.tile_frames[fi]];
            const uint8_t* const dist = F.dist;
            const int dpitch = F.dist_pitch;
            // 3) sample + accumulate
            // Interior tile-frames (tile inside the bbox, its four corners at least a pixel inside the source;
            // the plane map is monotone in u and in v, so the corners bound every pixel): border-free loop
            // with all taps of the thread's PPT pixels requested before any is used.
            bool interior = false;
            if (F.kind == XF_PLANE && !F.seam && !F.gainmap &&
                F.k[6] == 0.f && F.k[7] == 0.f && F.k8one == 1.f) {
                const int u0 = X0 - F.cx, v0 = Y0 - F.cy;
                if (u0 >= 0 && u0 + TW <= F.w && v0 >= 0 && v0 + TH <= F.h) {
                    float xs[4], ys[4];
                    for (int cidx = 0; cidx < 4; cidx++) {
                        float U = (float)(F.tlx + u0 + ((cidx & 1) ? TW - 1 : 0)), V = (float)(F.tly + v0 + ((cidx & 2) ? TH - 1 : 0));
                        if (F.scale != 1.f) { U = f_div(U, F.scale); V = f_div(V, F.scale); }
                        const float up = f_sub(U, F.t0), vp = f_sub(V, F.t1);
                        xs[cidx] = f_add(f_add(f_mul(F.k[0], up), f_mul(F.k[1], vp)), F.k2one);
                        ys[cidx] = f_add(f_add(f_mul(F.k[3], up), f_mul(F.k[4], vp)), F.k5one);
                    }
                    const float xmn = fminf(fminf(xs[0], xs[1]), fminf(xs[2], xs[3])), xmx = fmaxf(fmaxf(xs[0], xs[1]), fmaxf(xs[2], xs[3]));
                    const float ymn = fminf(fminf(ys[0], ys[1]), fminf(ys[2], ys[3])), ymx = fmaxf(fmaxf(ys[0], ys[1]), fmaxf(ys[2], ys[3]));
                    interior = xmn >= 1.f && xmx <= (float)(F.src_w - 3) && ymn >= 1.f && ymx <= (float)(F.src_h - 3);
                }
            }
            if (interior) {
                const uint32_t* const src = F.src;
                const int pitch = F.src_pitch;
                const float k2 = F.k2one, k5 = F.k5one;
                for (int i = tid; i < TW + TH; i += NT) {
                    if (i < TW) {
                        float U = (float)(F.tlx + X0 + i - F.cx);
                        if (F.scale != 1.f) U = f_div(U, F.scale);
                        const float up = f_sub(U, F.t0);
                        float2 c; c.x = f_mul(F.k[0], up); c.y = f_mul(F.k[3], up);
                        s_cx[i] = c;
                    } else {
                        float V = (float)(F.tly + Y0 + (i - TW) - F.cy);
                        if (F.scale != 1.f) V = f_div(V, F.scale);
                        const float vp = f_sub(V, F.t1);
                        float2 r; r.x = f_mul(F.k[1], vp); r.y = f_mul(F.k[4], vp);
                        s_ry[i - TW] = r;
                    }
                }
                DS_SYNC();
                constexpr int UB = PPT < 4 ? PPT : 4;   // pixels in flight per thread
                DS_UNROLL
                for (int k0 = 0; k0 < PPT; k0 += UB) {
                    int ix[UB], iy[UB], dd[UB];
                    uint32_t p00[UB], p01[UB], p10[UB], p11[UB];
                    DS_UNROLL
                    for (int b = 0; b < UB; b++) {
                        const int pidx = tid + (k0 + b) * NT;
                        const int yy = pidx / TW, xx = pidx - yy * TW;
                        dd[b] = (int)ld_ro(dist + (size_t)(Y0 + yy - F.cy) * dpitch + (X0 + xx - F.cx));
                        const float2 c = s_cx[xx], r = s_ry[yy];
                        const float x = f_add(f_add(c.x, r.x), k2);
                        const float y = f_add(f_add(c.y, r.y), k5);
#if DS_CUDA
                        ix[b] = __float2int_rn(f_mul(x, 32.f)); iy[b] = __float2int_rn(f_mul(y, 32.f));
#else
                        ix[b] = f2i_rn(f_mul(x, 32.f)); iy[b] = f2i_rn(f_mul(y, 32.f));
#endif
                        const uint32_t* r0 = src + ((iy[b] >> 5) * pitch + (ix[b] >> 5));
                        p00[b] = ld_ro(r0); p01[b] = ld_ro(r0 + 1); p10[b] = ld_ro(r0 + pitch); p11[b] = ld_ro(r0 + pitch + 1);
                    }
                    DS_UNROLL
                    for (int b = 0; b < UB; b++) {
                        const int k = k0 + b;
                        const int pidx = tid + k * NT;
                        const int yy = pidx / TW, xx = pidx - yy * TW;
                        float wgt = 1.f;
                        {
                            const int d = dd[b];
                            if (d == 0) continue;
                            if (d < R) { wgt = f_mul((float)d, p.sharpness); if (wgt > 1.f) wgt = 1.f; }
                        }
                        (void)yy; (void)xx;
                        const int ax = ix[b] & 31, ay = iy[b] & 31;
                        const uint32_t wb = (uint32_t)(32 - ax) | ((uint32_t)ax << 8), wg = wb << 16;
                        const uint32_t t0 = byte_perm(p00[b], p01[b], 0x5140), t0r = byte_perm(p00[b], p01[b], 0x6262);
                        const uint32_t t1 = byte_perm(p10[b], p11[b], 0x5140), t1r = byte_perm(p10[b], p11[b], 0x6262);
                        const int wy1 = ay, wy0 = 32 - ay;
                        int ob = (dot4u(t0, wb, 0) * wy0 + dot4u(t1, wb, 0) * wy1 + 512) >> 10;
                        int og = (dot4u(t0, wg, 0) * wy0 + dot4u(t1, wg, 0) * wy1 + 512) >> 10;
                        int orr = (dot4u(t0r, wb, 0) * wy0 + dot4u(t1r, wb, 0) * wy1 + 512) >> 10;
                        if (F.any_gain) apply_gains(F, ob, og, orr, 0, 0);
                        if (wgt == 1.f) {   // trunc(p * 1) == p
                            acc[k][0] += ob; acc[k][1] += og; acc[k][2] += orr;
                        } else {
                            acc[k][0] += (int)(short)f2i_rz(f_mul((float)ob, wgt));
                            acc[k][1] += (int)(short)f2i_rz(f_mul((float)og, wgt));
                            acc[k][2] += (int)(short)f2i_rz(f_mul((float)orr, wgt));
                        }
                        ws[k] = f_add(ws[k], wgt);
                    }
                }
            } else {
            DS_UNROLL
            for (int k = 0; k < PPT; k++) {
                const int pidx = tid + k * NT;
                const int yy = pidx / TW, xx = pidx - yy * TW;
                const int u = X0 + xx - F.cx, v = Y0 + yy - F.cy;
                if ((unsigned)u >= (unsigned)F.w || (unsigned)v >= (unsigned)F.h) continue;
                float wgt = 1.f;
                {
                    const int d = (int)ld_ro(dist + (size_t)v * dpitch + u);
                    if (d == 0) continue;
                    if (d < R) { wgt = f_mul((float)d, p.sharpness); if (wgt > 1.f) wgt = 1.f; }
                }
                const Coord c = eval_coord(F, u, v);
                const px8 s = sample_bilinear(F, c, u, v);
                acc[k][0] += (int)(short)f2i_rz(f_mul((float)s.b, wgt));
                acc[k][1] += (int)(short)f2i_rz(f_mul((float)s.g, wgt));
                acc[k][2] += (int)(short)f2i_rz(f_mul((float)s.r, wgt));
                ws[k] = f_add(ws[k], wgt);
            }
            }
            DS_SYNC();
        }
        // normalise into a shared tile of packed pixels (b | g << 8 | r << 16 | mask << 24), then store it with 4-pixel
        // accesses: 12 BGR bytes as three words + one mask word, or one 16-byte BGRA store
        uint32_t* s_out = (uint32_t*)(smem + (TW + TH) * 8);
        DS_UNROLL
        for (int k = 0; k < PPT; k++) {
            const int pidx = tid + k * NT;
            const float den = f_add(ws[k], 1e-5f);
            const int m = ws[k] > 1e-5f;
            const float rden = rcp_refined(den);   // the three quotients share the reciprocal refinement (ds_device.h)
            int o[3];
            DS_UNROLL
            for (int ch = 0; ch < 3; ch++) {
                const short a16 = (short)acc[k][ch];
                const int v16 = (int)(short)f2i_rz(div_by_rcp((float)a16, den, rden));
                o[ch] = m ? sat8i(v16) : 0;
            }
            s_out[pidx] = (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16) | (m ? 0xff000000u : 0u);
        }
        DS_SYNC();
        for (int i = tid; i < TW * TH / 4; i += NT) {
            const int yy = i / (TW / 4), x4 = (i - yy * (TW / 4)) * 4;
            const int X = X0 + x4, Y = Y0 + yy;
            if (X >= p.o.w || Y >= p.o.h || Y < p.row0 || Y >= p.row1) continue;
            const uint4 px = *(const uint4*)(s_out + yy * TW + x4);
            const int nout = imin(4, p.o.w - X);
            if (p.o.fmt == 1) {
                uint32_t* q = (uint32_t*)(p.o.out + (size_t)Y * p.o.out_pitch) + X;
                if (nout == 4) *(uint4*)q = px;
                else { const uint32_t v[4] = {px.x, px.y, px.z, px.w}; for (int k = 0; k < nout; k++) q[k] = v[k]; }
            } else if (nout == 4) {
                // X is a multiple of 4 and the pitches are multiples of 256: aligned words
                uint32_t* q = (uint32_t*)(p.o.out + (size_t)Y * p.o.out_pitch + (size_t)X * 3);
                q[0] = (px.x & 0xffffffu) | (px.y << 24);
                q[1] = ((px.y >> 8) & 0xffffu) | (px.z << 16);
                q[2] = ((px.z >> 16) & 0xffu) | (px.w << 8);
                *(uint32_t*)(p.o.mask + (size_t)Y * p.o.mask_pitch + X) = (px.x >> 24) | ((px.y >> 24) << 8) | ((px.z >> 24) << 16) | ((px.w >> 24) << 24);
            } else {
                const uint32_t v[4] = {px.x, px.y, px.z, px.w};
                for (int k = 0; k < nout; k++) {
                    uint8_t* q = p.o.out + (size_t)Y * p.o.out_pitch + (size_t)(X + k) * 3;
                    q[0] = (uint8_t)v[k]; q[1] = (uint8_t)(v[k] >> 8); q[2] = (uint8_t)(v[k] >> 16);
                    p.o.mask[(size_t)Y * p.o.mask_pitch + X + k] = (uint8_t)(v[k] >> 24);
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// MULTIBAND feed, gather formulation (A11): the straightforward tile kernel. Instantiated for level 0 only
// (ds_mb_feed_l0_generic: tiles with more frames than the fast kernel stages, and the zero-band blender whose level 0 is
// the top level); the hot paths are MBFastBody (level 0) and PyrDownBody / AccumBody (levels >= 1) below.
// A canvas tile of level l walks the frames whose feed ROI touches it, in feed order. Per frame:
//   phase 1  G_l over the tile + halo into shared memory (level 0: inverse warp of the source,
//            copyMakeBorder(REFLECT) and the warped mask folded in; l >= 1: read the frame's G_l, W_l)
//   phase 2  G_{l+1} = pyrDown16S (tile/2 + 1 ring, own part written to the frame's pyramid),
//            W_{l+1} = pyrDownF32 (own part written)
//   phase 3  lap = sat(G_l - pyrUp(G_{l+1})); acc += trunc(lap * W_l) (int16 wrap); wsum += W_l
// After the last frame: dst_l = trunc(acc / (wsum + 1e-5)), flag = wsum > 1e-5.
// The top level (l == L) accumulates G_L itself.

struct MBParams {
    const FrameDev* frames;
    const int* tile_off; const int* tile_frames;  // CSR: frames per tile of this level
    const int* tile_ids;                           // tiles this launch processes (NULL: block == tile)
    const int4* tile_rec;                          // per launched tile {tile, list begin, list end, first frame}: one load instead of three dependent ones
    int tiles_x;
    int level, L;
    px16* dst; int dst_w, dst_h;  // normalised Laplacian level of the padded canvas
    int acc_y0, acc_y1;           // rows of this level whose dst the band's collapse reads (dst written only there)
    int own_y0, own_y1;           // even-aligned rows of this level the handle processes at all
    const void* tmaps;            // CUtensorMap[frame][2] over the frame sources (level 0: L2 prefetch box, shared-memory box), or NULL
    int flags;                    // bit 0: fast warp loop also for in-bounds tile-frames in the gap of the feed ROI; bit 1: level-0 source boxes; bit 2: ds_mb_accum ring; bit 3: no specialised loop for gap tile-frames (A/B)
    const void* lmaps; int lstride;   // CUtensorMap[frame][lstride][AccumBody::LM_N] over the per-frame G_l / W_l planes (levels 1 .. L), or NULL
    int rnd_bias;                 // DS_RND_BIAS (a parameter on purpose: see the level-0 box addressing)
    uint32_t m_tiles_x;           // floor(2^32 / tiles_x): tile index decode without a division (ds_mb_accum)
};

template <int T, bool LEVEL0>
struct MBBody {
    static constexpr int PW = T + 7, GW = T / 2 + 2, JW = T / 2;
    static constexpr int G_BYTES = ds_al16(PW * PW * 4);
    static constexpr int W_BYTES = LEVEL0 ? 0 : ds_al16(PW * PW * 4);
    static constexpr int G1_BYTES = ds_al16(GW * GW * 8);
    static constexpr int H_BYTES = ds_al16(PW * JW * 4);
    static constexpr int ACC_BYTES = T * T * 8, WS_BYTES = T * T * 4;
    static int smem_bytes() { return G_BYTES + W_BYTES + G1_BYTES + H_BYTES + ACC_BYTES + WS_BYTES; }

    template <int NT>
    DS_DM void run(const MBParams& p, int block, int tid, unsigned char* smem) {
        px8* s_g8 = (px8*)smem;
        float* s_w = (float*)(smem + G_BYTES);
        px16* s_g1 = (px16*)(smem + G_BYTES + W_BYTES);
        float* s_h = (float*)(smem + G_BYTES + W_BYTES + G1_BYTES);
        px16* s_acc = (px16*)(smem + G_BYTES + W_BYTES + G1_BYTES + H_BYTES);
        float* s_ws = (float*)(smem + G_BYTES + W_BYTES + G1_BYTES + H_BYTES + ACC_BYTES);

        const int tile = p.tile_ids ? p.tile_ids[block] : block;
        const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
        const int X0 = tx * T, Y0 = ty * T;
        const int l = p.level;
        const bool top = (l == p.L);

        for (int i = tid; i < T * T; i += NT) { px16 z; z.b = z.g = z.r = z.a = 0; s_acc[i] = z; s_ws[i] = 0.f; }
        DS_SYNC();

        for (int fi = p.tile_off[tile]; fi < p.tile_off[tile + 1]; fi++) {
            const FrameDev& F = p.frames[p.tile_frames[fi]];
            const int rx = F.rx >> l, ry = F.ry >> l, rw = F.rw >> l, rh = F.rh >> l;
            const int ax0 = imax(X0, rx), ax1 = imin(X0 + T, rx + rw);
            const int ay0 = imax(imax(Y0, ry), p.own_y0), ay1 = imin(imin(Y0 + T, ry + rh), p.own_y1);
            if (ax0 >= ax1 || ay0 >= ay1) continue;  // block-uniform
            const int ox0 = ax0 - rx, ox1 = ax1 - rx, oy0 = ay0 - ry, oy1 = ay1 - ry;  // own, ROI-relative
            int n1x = 0, n1y = 0, jx0 = 0, jx1 = 0, jy0 = 0, jy1 = 0, gx0 = 0, gx1 = 0, gy0 = 0, gy1 = 0;
            int px0 = ox0, px1 = ox1 - 1, py0 = oy0, py1 = oy1 - 1;
            if (!top) {
                n1x = rw >> 1; n1y = rh >> 1;
                jx0 = ox0 >> 1; jx1 = (ox1 + 1) >> 1; jy0 = oy0 >> 1; jy1 = (oy1 + 1) >> 1;
                gx0 = imax(jx0 - 1, 0); gx1 = imin(jx1, n1x - 1);
                gy0 = imax(jy0 - 1, 0); gy1 = imin(jy1, n1y - 1);
                px0 = imax(2 * gx0 - 2, 0); px1 = imin(2 * gx1 + 2, rw - 1);
                py0 = imax(2 * gy0 - 2, 0); py1 = imin(2 * gy1 + 2, rh - 1);
            }
            const int pw = px1 - px0 + 1, ph = py1 - py0 + 1;

            // ---- phase 1: G_l (+ W_l) over [px0..px1] x [py0..py1]
            for (int i = tid; i < pw * ph; i += NT) {
                const int yy = i / pw, xx = i - yy * pw;
                const int px = px0 + xx, py = py0 + yy;
                if (LEVEL0) {
                    const int u = F.rx + px - F.cx, v = F.ry + py - F.cy;
                    const bool inside = (unsigned)u < (unsigned)F.w && (unsigned)v < (unsigned)F.h;
                    const int ur = refl(u, F.w, BORDER_REFL), vr = refl(v, F.h, BORDER_REFL);
                    const Coord c = eval_coord(F, ur, vr);
                    px8 s = sample_bilinear(F, c, ur, vr);
                    s.a = (unsigned char)(inside ? mask_value(F, c, u, v) : 0);
                    s_g8[i] = s;
                } else {
                    const size_t gi = (size_t)py * F.gp[l] + px;
                    s_g8[i] = F.G[l][gi];
                    s_w[i] = F.W[l][gi];
                }
            }
            DS_SYNC();

            if (!top) {
                const int gw = gx1 - gx0 + 1, gh = gy1 - gy0 + 1;
                // ---- phase 2a: G_{l+1} = pyrDown16S over the g-range
                for (int i = tid; i < gw * gh; i += NT) {
                    const int gyy = i / gw, gxx = i - gyy * gw;
                    const int gx = gx0 + gxx, gy = gy0 + gyy;
                    int cxi[5], sb = 0, sg = 0, sr = 0;
                    for (int k = 0; k < 5; k++) cxi[k] = refl101(2 * gx + k - 2, rw) - px0;
                    for (int ky = 0; ky < 5; ky++) {
                        const int sy = refl101(2 * gy + ky - 2, rh) - py0;
                        const int kwy = (ky == 0 || ky == 4) ? 1 : ((ky == 2) ? 6 : 4);
                        int rb = 0, rg = 0, rr = 0;
                        for (int kx = 0; kx < 5; kx++) {
                            const int kwx = (kx == 0 || kx == 4) ? 1 : ((kx == 2) ? 6 : 4);
                            const int si = sy * pw + cxi[kx];
                            int vb, vg, vr;
                            { const px8 q = s_g8[si]; vb = q.b; vg = q.g; vr = q.r; }
                            rb += kwx * vb; rg += kwx * vg; rr += kwx * vr;
                        }
                        sb += kwy * rb; sg += kwy * rg; sr += kwy * rr;
                    }
                    px16 o;
                    o.b = (short)((sb + 128) >> 8); o.g = (short)((sg + 128) >> 8); o.r = (short)((sr + 128) >> 8); o.a = 0;
                    s_g1[i] = o;
                    if (gx >= jx0 && gx < jx1 && gy >= jy0 && gy < jy1) {
                        px8 o8;
                        o8.b = (unsigned char)o.b; o8.g = (unsigned char)o.g; o8.r = (unsigned char)o.r; o8.a = 0;
                        F.G[l + 1][(size_t)gy * F.gp[l + 1] + gx] = o8;
                    }
                }
                // ---- phase 2b: horizontal pass of the weight pyrDown, rows hr0..hr1, own columns
                const int jw = jx1 - jx0;
                const int hr0 = imax(2 * jy0 - 2, 0), hr1 = imin(2 * jy1, rh - 1);
                const int hn = hr1 - hr0 + 1;
                for (int i = tid; i < hn * jw; i += NT) {
                    const int rr_ = i / jw, jj = i - rr_ * jw;
                    const int row = hr0 + rr_, j = jx0 + jj;
                    float t[5];
                    for (int k = 0; k < 5; k++) {
                        const int si = (row - py0) * pw + (refl101(2 * j + k - 2, rw) - px0);
                        t[k] = LEVEL0 ? f_mul((float)s_g8[si].a, 1.f / 255.f) : s_w[si];
                    }
                    s_h[i] = pd_h_is_simd(j, rw, n1x) ? pd_h_simd(t[0], t[1], t[2], t[3], t[4])
                                                       : pd_scalar(t[0], t[1], t[2], t[3], t[4]);
                }
                DS_SYNC();
                // ---- phase 2c: vertical pass, own (jy, j) -> W_{l+1}
                for (int i = tid; i < (jy1 - jy0) * jw; i += NT) {
                    const int jyy = i / jw, jj = i - jyy * jw;
                    const int jy = jy0 + jyy, j = jx0 + jj;
                    float t[5];
                    for (int k = 0; k < 5; k++) t[k] = s_h[(refl101(2 * jy + k - 2, rh) - hr0) * jw + jj];
                    const float v = pd_v_is_simd(j, n1x) ? pd_v_simd(t[0], t[1], t[2], t[3], t[4])
                                                         : pd_scalar(t[0], t[1], t[2], t[3], t[4]);
                    F.W[l + 1][(size_t)jy * F.gp[l + 1] + j] = f_mul(v, 1.f / 256.f);
                }
                // ---- phase 3: Laplacian + weighted accumulate over the own pixels
                const int ow = ox1 - ox0, oh = oy1 - oy0;
                for (int i = tid; i < ow * oh; i += NT) {
                    const int yy = i / ow, xx = i - yy * ow;
                    const int ox = ox0 + xx, oy = oy0 + yy;
                    const int c1x = ox >> 1, c1y = oy >> 1;
                    const int ixl = up_l(c1x, n1x) - gx0, ixc = c1x - gx0, ixr = up_r(c1x, n1x) - gx0;
                    const int iyl = up_l(c1y, n1y) - gy0, iyc = c1y - gy0, iyr = up_r(c1y, n1y) - gy0;
                    int up[3];
                    {
                        // horizontal on the three source rows, then vertical (A10)
                        int hl[3], hc[3], hrr[3];
                        const px16 a0 = s_g1[iyl * gw + ixl], a1 = s_g1[iyl * gw + ixc], a2 = s_g1[iyl * gw + ixr];
                        const px16 b0 = s_g1[iyc * gw + ixl], b1 = s_g1[iyc * gw + ixc], b2 = s_g1[iyc * gw + ixr];
                        const px16 c0 = s_g1[iyr * gw + ixl], c1 = s_g1[iyr * gw + ixc], c2 = s_g1[iyr * gw + ixr];
                        if (ox & 1) {
                            hl[0] = 4 * (a1.b + a2.b); hl[1] = 4 * (a1.g + a2.g); hl[2] = 4 * (a1.r + a2.r);
                            hc[0] = 4 * (b1.b + b2.b); hc[1] = 4 * (b1.g + b2.g); hc[2] = 4 * (b1.r + b2.r);
                            hrr[0] = 4 * (c1.b + c2.b); hrr[1] = 4 * (c1.g + c2.g); hrr[2] = 4 * (c1.r + c2.r);
                        } else {
                            hl[0] = a0.b + 6 * a1.b + a2.b; hl[1] = a0.g + 6 * a1.g + a2.g; hl[2] = a0.r + 6 * a1.r + a2.r;
                            hc[0] = b0.b + 6 * b1.b + b2.b; hc[1] = b0.g + 6 * b1.g + b2.g; hc[2] = b0.r + 6 * b1.r + b2.r;
                            hrr[0] = c0.b + 6 * c1.b + c2.b; hrr[1] = c0.g + 6 * c1.g + c2.g; hrr[2] = c0.r + 6 * c1.r + c2.r;
                        }
                        for (int ch = 0; ch < 3; ch++) {
                            const int vv = (oy & 1) ? 4 * (hc[ch] + hrr[ch]) : (hl[ch] + 6 * hc[ch] + hrr[ch]);
                            up[ch] = (int)(short)((vv + 32) >> 6);
                        }
                    }
                    const int si = (oy - py0) * pw + (ox - px0);
                    int gb, gg, gr; float wv;
                    { const px8 q = s_g8[si]; gb = q.b; gg = q.g; gr = q.r; wv = LEVEL0 ? f_mul((float)q.a, 1.f / 255.f) : s_w[si]; }
                    const int ti = (ay0 + yy - Y0) * T + (ax0 + xx - X0);
                    px16 a = s_acc[ti];
                    a.b = (short)(a.b + (short)f2i_rz(f_mul((float)sat16i(gb - up[0]), wv)));
                    a.g = (short)(a.g + (short)f2i_rz(f_mul((float)sat16i(gg - up[1]), wv)));
                    a.r = (short)(a.r + (short)f2i_rz(f_mul((float)sat16i(gr - up[2]), wv)));
                    s_acc[ti] = a;
                    s_ws[ti] = f_add(s_ws[ti], wv);
                }
            } else {
                const int ow = ox1 - ox0, oh = oy1 - oy0;
                for (int i = tid; i < ow * oh; i += NT) {
                    const int yy = i / ow, xx = i - yy * ow;
                    const int si = yy * pw + xx;
                    int gb, gg, gr; float wv;
                    { const px8 q = s_g8[si]; gb = q.b; gg = q.g; gr = q.r; wv = LEVEL0 ? f_mul((float)q.a, 1.f / 255.f) : s_w[si]; }
                    const int ti = (ay0 + yy - Y0) * T + (ax0 + xx - X0);
                    px16 a = s_acc[ti];
                    a.b = (short)(a.b + (short)f2i_rz(f_mul((float)gb, wv)));
                    a.g = (short)(a.g + (short)f2i_rz(f_mul((float)gg, wv)));
                    a.r = (short)(a.r + (short)f2i_rz(f_mul((float)gr, wv)));
                    s_acc[ti] = a;
                    s_ws[ti] = f_add(s_ws[ti], wv);
                }
            }
            DS_SYNC();
        }

        // ---- normalise and store the dst level (only rows the band owns)
        for (int i = tid; i < T * T; i += NT) {
            const int yy = i / T, xx = i - yy * T;
            const int X = X0 + xx, Y = Y0 + yy;
            if (X >= p.dst_w || Y >= p.dst_h || Y < p.acc_y0 || Y >= p.acc_y1) continue;
            const px16 a = s_acc[i];
            const float wsum = s_ws[i];
            const float den = f_add(wsum, 1e-5f);
            px16 o;
            o.b = (short)f2i_rz(f_div((float)a.b, den));
            o.g = (short)f2i_rz(f_div((float)a.g, den));
            o.r = (short)f2i_rz(f_div((float)a.r, den));
            o.a = (short)(wsum > 1e-5f ? 1 : 0);
            p.dst[(size_t)Y * p.dst_w + X] = o;
        }
    }
};

#if DS_CUDA
// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers: box-shaped tile loads of the per-frame pyramid levels
DS_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DS_D void mbar_init(void* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DS_D void mbar_expect_tx(void* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DS_D void tma_load_2d(void* dst, const void* tmap, int c0, int c1, void* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
DS_D void mbar_wait(void* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (int spin = 0; spin < (1 << 24); spin++) {
        uint32_t done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();   // a lost TMA transaction must fail loudly, never hang the device
}
// L2 prefetch of a tile (no shared memory, no completion tracking): warms L2 for loads that follow later
DS_D void tma_prefetch_l2_2d(const void* tmap, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}
DS_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// descriptors live in global memory (one per frame and level): acquire them for the tensormap proxy before use
DS_D void fence_tensormap_acquire(const void* tmap) {
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
}
DS_D void fence_proxy_async_smem() { fence_proxy_async(); }
// 16 bytes global -> shared without a register round trip; complete (for the issuing thread) after cp_async_wait_all
DS_D void cp_async16(void* dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory"); }
DS_D void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// One bw x bh box of a pitched plane of 4-byte elements into shared memory (TMA tile load, completion on `bar`); elements
// outside the plane arrive as zeros. The emulator copies.
// (The descriptors were written by an earlier grid - ds_meta_copy - so no tensormap-proxy fence is needed here.)
DS_D void box_load(void* dst, const void* tmap, const void*, int, int, int, int x0, int y0, int, int, void* bar) {
    tma_load_2d(dst, tmap, x0, y0, bar);
}
#else
DS_D void mbar_init(void*, int) {}
DS_D void mbar_expect_tx(void*, uint32_t) {}
DS_D void mbar_wait(void*, uint32_t) {}
DS_D void fence_proxy_async_smem() {}
DS_D void cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
DS_D void cp_async_wait_all() {}
DS_D void box_load(void* dst, const void*, const void* plane, int w, int h, int pitch, int x0, int y0, int bw, int bh, void*) {
    uint32_t* d = (uint32_t*)dst;
    const uint32_t* s = (const uint32_t*)plane;
    for (int y = 0; y < bh; y++)
        for (int x = 0; x < bw; x++) {
            const int sx = x0 + x, sy = y0 + y;
            d[y * bw + x] = ((unsigned)sx < (unsigned)w && (unsigned)sy < (unsigned)h) ? s[(size_t)sy * pitch + sx] : 0u;
        }
}
#endif

// ---------------------------------------------------------------------------------------------
// MULTIBAND feed, level 0, PLANE_F32 frames: the hot kernel. Same arithmetic and gather structure as
// MBBody<64, true>, restructured for instruction count:
//   * per tile-frame column / row tables hold the reflected bbox index and the per-column / per-row
//     products of the plane map (k0*u', k3*u', k6*u' / k1*v', k4*v', k7*v'), so a pixel costs two adds
//     per coordinate (each product and sum is still rounded separately, as OpenCV does);
//   * the source footprint of a tile-frame staged in shared memory by one TMA box load; bilinear taps as two-way dot
//     products with 16-bit 2-D weights (dp2a) on the BGRX words, cvRound on the FMA pipe (biased mantissa);
//   * G_0 / G_1 kept as packed 16-bit lanes (B|R<<16, G) so the separable 5-tap pyrDown and the 2x2
//     pyrUp quads run two channels per integer op (all partial sums fit 16 bits: <= 255*256);
//   * block-uniform shortcuts: a tile-frame whose mask is all 255 has W_1 == 1 exactly and
//     trunc(lap * 1) == lap; all 0 contributes nothing but still produces G_1;
//   * tile-frames that touch a ROI border use index-reflecting variants of the same phases.
//   * specialised straight-line variants of the warp loop and of the Laplacian phase for the common tile-frame (whole
//     71 x 71 region, mask 255 or - wholly in the gap of the feed ROI - 0): immediate-offset shared accesses, no clamps.
// acc lanes: B + 65536 * R in one int (exact while |sum| < 2^15; the host sends tiles with more than MAXF = 24 frames to
// the generic kernel).

// Normalised Laplacian pixel of a destination level (A11 blend): trunc(sum / (wsum + 1e-5)) per channel, flag = wsum >
// 1e-5, as the two words of a px16 (b | g << 16, r | flag << 16). The three quotients share the reciprocal refinement
// (div_by_rcp); the sums are taken modulo 2^16 like the reference's short accumulators.
DS_D void norm_px(int sb, int sg, int sr, float wsum, uint32_t& w0, uint32_t& w1) {
    const float den = f_add(wsum, 1e-5f);
    const float r = rcp_refined(den);
    const int vb = (int)(short)f2i_rz(div_by_rcp((float)(short)sb, den, r));
    const int vg = (int)(short)f2i_rz(div_by_rcp((float)(short)sg, den, r));
    const int vr = (int)(short)f2i_rz(div_by_rcp((float)(short)sr, den, r));
    w0 = ((uint32_t)vb & 0xffffu) | ((uint32_t)vg << 16);
    w1 = ((uint32_t)vr & 0xffffu) | (wsum > 1e-5f ? 0x10000u : 0u);
}

template <bool V> struct BoolTag { static constexpr bool value = V; };
template <int V> struct IntTag { static constexpr int value = V; };
struct L0Col { float a0, a3, a6; int u; };  // u = bbox column if inside the bbox, else ~(reflected column)
struct L0Row { float b1, b4, b7; int v; };

// AFF ("general"): the fast loop also builds cv::warpAffine (integer) coordinates and takes seam masks and gain maps;
// a separate instantiation so that the kernel of the headline configuration (plane maps, no per-pixel inputs) keeps
// its instruction footprint: with AFF = false such tile-frames go through the per-pixel loop.
template <int T_, bool LEVEL0, bool AFF = false>
struct MBFastBody {
    // PWS: row pitch of the needed region in smem. Levels >= 1 load it with TMA, whose innermost start
    // coordinate must be 16-byte aligned: the region starts at px0 rounded down to 4 px (up to 3 extra columns).
    static constexpr int T = T_, PWS = LEVEL0 ? T_ + 8 : T_ + 12, PHM = T_ + 7, GWS = T_ / 2 + 2, JW = T_ / 2, NQ = T_ * T_ / 4;
    static constexpr int W0_BYTES = LEVEL0 ? 0 : ds_al128(PHM * PWS * 4);   // W_l of the needed region (levels >= 1)
    static constexpr int G0_BYTES = ds_al128(PHM * PWS * 4);
    static constexpr int G1_BYTES = ds_al128(GWS * GWS * 8);
    // level 0, plane-only kernel: the source footprint of a tile-frame (the box the inverse map of the 71 x 71 region
    // fits in for rotations up to ~3 degrees) is brought into shared memory by one TMA tile load and the bilinear taps
    // are read from there. The box shares its buffer with the pyrDown H pass: it is free from the end of a frame's
    // vertical pass to the next frame's horizontal pass, which is when the next frame's box is in flight.
    static constexpr bool HAS_BOX = LEVEL0 && !AFF;
    // (the innermost start coordinate of a TMA tile must be 16-byte aligned: the box starts at a multiple of 4 pixels, up
    // to 3 columns left of the footprint)
    static constexpr int BOXW = 84, BOXH = 80, BOX_BYTES = BOXW * BOXH * 4;
    static constexpr int H_ONLY_BYTES = ds_al128(PHM * GWS * 8);   // pyrDown H pass; later the float H pass of the weights (PHM * JW * 4)
    static constexpr int H_BYTES = (HAS_BOX && BOX_BYTES > H_ONLY_BYTES) ? ds_al128(BOX_BYTES) : H_ONLY_BYTES;
    static constexpr int ACC_BYTES = T * T * 4;
    static constexpr int COL_BYTES = ds_al128(PWS * 16), ROW_BYTES = ds_al128(PHM * 16);
    static constexpr int NTAB = LEVEL0 ? 2 : 1;                // level 0: the tables of the next frame are built while this one is processed
    static constexpr int TAB_BYTES = NTAB * (COL_BYTES + ROW_BYTES);
    static constexpr int MAXF = LEVEL0 ? 24 : 64;              // frames per tile of the fast kernels (host-checked; the packed accumulators allow 64)
    static constexpr int GEO_OFF = G0_BYTES + G1_BYTES + H_BYTES + 3 * ACC_BYTES + TAB_BYTES + W0_BYTES;
    static constexpr int FDEV_BYTES = (int)((sizeof(FrameDev) + 15) & ~(size_t)15);
    static constexpr int GEO_BYTES = (LEVEL0 && !AFF) ? 128 : 96;   // sizeof(TFGeo)
    static constexpr int FDEV_OFF = GEO_OFF + MAXF * GEO_BYTES;      // two FrameDev slots: current frame / prefetch of the next
    // levels >= 1: second copy of the G / W boxes so the TMA load of the next frame overlaps this frame's compute
    static constexpr int G0B_OFF = ds_al128(FDEV_OFF + 2 * FDEV_BYTES);   // TMA destinations: 128-B aligned
    static constexpr int W0B_OFF = G0B_OFF + (LEVEL0 ? 0 : G0_BYTES);
    static constexpr int MBAR_OFF = W0B_OFF + W0_BYTES;
    static constexpr uint32_t TMA_BYTES = (uint32_t)(PWS * PHM * 4);   // one box (G or W) of the needed region
    static int smem_bytes() { return MBAR_OFF + 32; }

    struct U2 { uint32_t br, g; };
    // tile x frame geometry, computed once per tile by as many threads as the tile has frames
    struct alignas(16) TFGeo {
        int skip, rx, ry, rw;
        int rh, ax0, ax1, ay0;
        int ay1, gx0, gx1, gy0;
        int gy1, px0, py0, pw;
        int ph, jx0, jx1, jy0;
        int jy1, border, pad0, pad1;
    };
    // level 0, plane-only kernel: which warp loop the tile-frame takes (decided once, by the thread that builds its entry)
    struct alignas(16) TFGeo0 : TFGeo {
        int mode;         // -1: per-pixel loop; 0 / 1 / 2: v2 loop interior / gap / edge
        int gap_all0;     // mode 1: the needed region misses the bbox altogether (mask 0 throughout)
        long long boff;   // modes 0 / 1: byte offset of the biased tap address base from F.src
        int box;          // modes 0 / 1: the footprint fits the shared-memory box whose origin in the source is (bx0, by0)
        int bx0, by0, pad2;
    };
    typedef typename std::conditional<(LEVEL0 && !AFF), TFGeo0, TFGeo>::type Geo;
    static_assert(sizeof(Geo) == GEO_BYTES, "TFGeo layout");
    static_assert(LEVEL0 && T_ == 64, "level 0 only: the levels above are ds_mb_pyrdown / ds_mb_accum");
    // mirrored index interval [mn, mx] of [lo, hi] on an axis of length n (single reflection only)
    DS_DM bool fold_range(int lo, int hi, int n, int& mn, int& mx) {
        if (lo < -n || hi >= 2 * n) return false;
        mn = lo >= 0 ? imin(lo, n - 1) : (hi >= 0 ? 0 : -hi - 1);
        mx = hi < n ? imax(hi, 0) : (lo < n ? n - 1 : 2 * n - 1 - lo);
        if (lo < 0) mx = imax(mx, imin(-lo - 1, n - 1));
        if (hi >= n) mn = imin(mn, imax(2 * n - 1 - hi, 0));
        return true;
    }
    // Which v2 warp loop a tile-frame of the plane-only kernel can take (see the loop for the modes). The region's bbox
    // indices (mirrored into the bbox where it lies in the gap of the feed ROI) cover one contiguous interval per axis,
    // and x(u, v) is monotone in u and in v even in float arithmetic (a chain of monotone roundings), so the values at
    // the interval ends bound every pixel.
    DS_DM void classify_plane(const FrameDev& F, TFGeo0& g, int flags) {
        g.mode = -1; g.gap_all0 = 0; g.boff = 0; g.box = 0; g.bx0 = g.by0 = g.pad2 = 0;
        const bool proj = !(F.k[6] == 0.f && F.k[7] == 0.f && F.k8one == 1.f);
        if (g.skip || proj || F.kind != XF_PLANE || F.seam || F.gainmap || F.any_gain) return;
        const int sw = F.src_w, sh = F.src_h, pitch = F.src_pitch;
        if (sw > 32768 || sh > 32768) return;   // the remap tables saturate at 16 bits; the biased rounding needs |32 x| < 2^22
        const int u_lo = g.rx + g.px0 - F.cx, u_hi = u_lo + g.pw - 1, v_lo = g.ry + g.py0 - F.cy, v_hi = v_lo + g.ph - 1;
        int umn = 0, umx = 0, vmn = 0, vmx = 0;
        if (!fold_range(u_lo, u_hi, F.w, umn, umx) || !fold_range(v_lo, v_hi, F.h, vmn, vmx)) return;
        float ca0[2], ca3[2], rb1[2], rb4[2];
        DS_UNROLL
        for (int e = 0; e < 2; e++) {
            float U = (float)(F.tlx + (e ? umx : umn)), V = (float)(F.tly + (e ? vmx : vmn));
            if (F.scale != 1.f) { U = f_div(U, F.scale); V = f_div(V, F.scale); }
            const float up = f_sub(U, F.t0), vp = f_sub(V, F.t1);
            ca0[e] = f_mul(F.k[0], up); ca3[e] = f_mul(F.k[3], up);
            rb1[e] = f_mul(F.k[1], vp); rb4[e] = f_mul(F.k[4], vp);
        }
        float xmn = 3.0e38f, xmx = -3.0e38f, ymn = 3.0e38f, ymx = -3.0e38f;
        DS_UNROLL
        for (int e = 0; e < 4; e++) {
            const float xx_ = f_add(f_add(ca0[e & 1], rb1[e >> 1]), F.k2one), yy_ = f_add(f_add(ca3[e & 1], rb4[e >> 1]), F.k5one);
            xmn = fminf(xmn, xx_); xmx = fmaxf(xmx, xx_); ymn = fminf(ymn, yy_); ymx = fmaxf(ymx, yy_);
        }
        const bool inbounds = xmn >= 1.f && xmx <= (float)(sw - 3) && ymn >= 1.f && ymx <= (float)(sh - 3);
        const bool interior = inbounds && u_lo >= 0 && u_hi < F.w && v_lo >= 0 && v_hi < F.h;
        if (!interior && !(flags & 1)) return;
        if (inbounds) {
            // Tap addresses are formed from the biased mantissas directly: with C = bias >> 5,
            // t = (sy + C) * pitch + (sx + C) mod 2^32 = off + K, and base' = base + 4 * (off_lo - u_lo) makes
            // base' + 4 * t the tap address as long as u_lo + span does not wrap
            constexpr int CB = DS_RND_BIAS >> 5;
            const long long off_lo = (long long)((int)floorf(ymn) - 1) * pitch + ((int)floorf(xmn) - 1);
            const long long off_hi = (long long)((int)floorf(ymx) + 3) * pitch + ((int)floorf(xmx) + 3);
            const uint32_t K32 = (uint32_t)CB * (uint32_t)pitch + (uint32_t)CB;
            const uint32_t u_lo32 = (uint32_t)off_lo + K32;
            if (off_lo < 0 || (unsigned long long)u_lo32 + (unsigned long long)(off_hi - off_lo) > 0xFFFFFFFFull) return;   // one in ~10^4
            g.boff = 4 * (off_lo - (long long)u_lo32);
            g.mode = interior ? 0 : 1;
            // taps: columns floor(xmn) - 1 .. floor(xmx) + 2, rows likewise
            g.bx0 = ((int)floorf(xmn) - 1) & ~3; g.by0 = (int)floorf(ymn) - 1;
            g.box = ((flags & 2) && (int)floorf(xmx) + 2 - g.bx0 < BOXW && (int)floorf(ymx) + 2 - g.by0 < BOXH) ? 1 : 0;
            g.gap_all0 = (u_hi < 0 || u_lo >= F.w || v_hi < 0 || v_lo >= F.h) ? 1 : 0;
        } else if (F.border != BORDER_CONST &&
                   xmn >= -(float)(sw - 1) && xmx <= (float)(2 * sw - 3) && ymn >= -(float)(sh - 1) && ymx <= (float)(2 * sh - 3) &&
                   xmx <= 32766.f && ymx <= 32766.f) {
            g.mode = 2;   // some taps leave the source, but by less than its size: one reflection resolves them
        }
    }

    // 5-tap [1 4 6 4 1] on packed lanes
    DS_DM uint32_t tap5(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e) { return a + e + 6u * c + 4u * (b + d); }

    template <int NT>
    DS_DM void run(const MBParams& p, int block, int tid, unsigned char* smem) {
        uint32_t* s_g0 = (uint32_t*)smem;   // levels >= 1 with TMA: alternates between two box buffers
        U2* s_g1 = (U2*)(smem + G0_BYTES);
        U2* s_h = (U2*)(smem + G0_BYTES + G1_BYTES);
        float* s_hw = (float*)(smem + G0_BYTES + G1_BYTES);
        // accumulators: {B + 65536 R, G} interleaved (one 128-bit access per pixel pair), weight sums apart
        int2* s_acc = (int2*)(smem + G0_BYTES + G1_BYTES + H_BYTES);
        float* s_ws = (float*)(s_acc + T * T);
        constexpr int TAB_OFF = G0_BYTES + G1_BYTES + H_BYTES + 3 * ACC_BYTES;
        auto col_buf = [&](int b_) { return (L0Col*)(smem + TAB_OFF + (b_ & (NTAB - 1)) * (COL_BYTES + ROW_BYTES)); };
        auto row_buf = [&](int b_) { return (L0Row*)(smem + TAB_OFF + (b_ & (NTAB - 1)) * (COL_BYTES + ROW_BYTES) + COL_BYTES); };
        float* s_w = (float*)(smem + TAB_OFF + TAB_BYTES);
        const int l = LEVEL0 ? 0 : p.level;

        const int4 rec = ld_ro(p.tile_rec + block);
        const int tile = rec.x;
        const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
        const int X0 = tx * T, Y0 = ty * T;
        const float c255 = f_mul(255.f, 1.f / 255.f);

        for (int i = tid; i < T * T; i += NT) { s_acc[i] = make_i2(0, 0); s_ws[i] = 0.f; }
        int* const s_vote = (int*)(smem + MBAR_OFF + 16);   // block-wide OR of the uniformity bits (block_or_bits)
        uint32_t box_phase = 0u;                            // level 0: parity of the source-box barrier
        if (tid == 0) *s_vote = 0;
#if DS_CUDA
        unsigned long long* s_bar = (unsigned long long*)(smem + MBAR_OFF);
        if (HAS_BOX && (p.flags & 2) && tid == 0) mbar_init(s_bar, 1);   // one barrier: the source box
#endif
        // (no barrier: the accumulators are first touched after several more)

        Geo* s_geo = (Geo*)(smem + GEO_OFF);
        const int f_begin = rec.y, f_end = rec.z;
        // Frame descriptors are staged in shared memory: slot (fi & 1) holds frame fi, and while it is being
        // processed the first threads prefetch frame fi + 1 into the other slot (visible after the barriers
        // every iteration ends with). Keeps the dependent global loads off the per-tile-frame critical path.
        auto stage_frame_at = [&](int frame_idx, int fi_) {
            const uint4* srcw = (const uint4*)(p.frames + frame_idx);
            uint4* dstw = (uint4*)(smem + FDEV_OFF + (fi_ & 1) * FDEV_BYTES);
            // the last warps copy: the first ones build the tile-frame geometry at the same time. Asynchronous copies
            // (cp.async): the copying warp does not wait for the data; stage_wait() before the barrier the readers pass
            for (int w = NT - 1 - tid; w < (int)(sizeof(FrameDev) / 16); w += NT) cp_async16(dstw + w, srcw + w);
        };
        auto stage_frame = [&](int fi_) { stage_frame_at(p.tile_frames[fi_], fi_); };
        if (f_begin < f_end) stage_frame_at(rec.w, f_begin);   // the first frame's index came with the tile record
        for (int j = tid; j < f_end - f_begin && j < MAXF; j += NT) {
            const FrameDev& F = p.frames[j == 0 ? rec.w : p.tile_frames[f_begin + j]];
            Geo g;
            g.rx = F.rx >> l; g.ry = F.ry >> l; g.rw = F.rw >> l; g.rh = F.rh >> l;
            g.ax0 = imax(X0, g.rx); g.ax1 = imin(X0 + T, g.rx + g.rw);
            g.ay0 = imax(imax(Y0, g.ry), p.own_y0); g.ay1 = imin(imin(Y0 + T, g.ry + g.rh), p.own_y1);
            g.skip = (g.ax0 >= g.ax1 || g.ay0 >= g.ay1) ? 1 : 0;
            const int ox0 = g.ax0 - g.rx, ox1 = g.ax1 - g.rx, oy0 = g.ay0 - g.ry, oy1 = g.ay1 - g.ry;
            const int n1x = g.rw >> 1, n1y = g.rh >> 1;
            g.jx0 = ox0 >> 1; g.jx1 = (ox1 + 1) >> 1; g.jy0 = oy0 >> 1; g.jy1 = (oy1 + 1) >> 1;
            g.gx0 = imax(g.jx0 - 1, 0); g.gx1 = imin(g.jx1, n1x - 1);
            g.gy0 = imax(g.jy0 - 1, 0); g.gy1 = imin(g.jy1, n1y - 1);
            g.px0 = LEVEL0 ? imax(2 * g.gx0 - 2, 0) : (imax(2 * g.gx0 - 2, 0) & ~3);
            const int px1 = imin(2 * g.gx1 + 2, g.rw - 1);
            g.py0 = imax(2 * g.gy0 - 2, 0);
            const int py1 = imin(2 * g.gy1 + 2, g.rh - 1);
            g.pw = px1 - g.px0 + 1; g.ph = py1 - g.py0 + 1;
            // no index reflection anywhere in this tile-frame?
            g.border = !(2 * g.gx0 - 2 >= 0 && 2 * g.gx1 + 2 <= g.rw - 1 && 2 * g.gy0 - 2 >= 0 && 2 * g.gy1 + 2 <= g.rh - 1 &&
                         g.jx0 >= 1 && g.jx1 <= n1x - 1 && g.jy0 >= 1 && g.jy1 <= n1y - 1) ? 1 : 0;
            g.pad0 = g.pad1 = 0;
            if constexpr (LEVEL0 && !AFF) classify_plane(F, g, p.flags);
            s_geo[j] = g;
        }
        cp_async_wait_all();
        DS_SYNC();

        // ---- level-0 tables of one tile-frame: reflected bbox index + per-column / per-row map terms. Built for frame
        // fi + 1 while frame fi is processed (two buffers), so the warp loop starts without a barrier of its own.
        auto build_tables = [&](const FrameDev& Fx, int fj) {
            const Geo gx = s_geo[fj - f_begin];
            if (gx.skip) return;
            L0Col* const t_col = col_buf(fj);
            L0Row* const t_row = row_buf(fj);
            const bool aff = AFF && LEVEL0 && Fx.kind == XF_AFFINE;
            // (the last warps build the tables: the first ones have the extra items of the pyrDown pass that follows)
            for (int i = NT - 1 - tid; i < PWS + gx.ph; i += NT) {
                if (i < PWS) {
                    const int u = gx.rx + gx.px0 + imin(i, gx.pw - 1) - Fx.cx;   // padding columns repeat the last one
                    const int ur = refl(u, Fx.w, BORDER_REFL);
                    L0Col c;
                    if (aff) {
                        // cv::warpAffine: adelta / bdelta of the column, 10 fractional bits (A6), kept as integer bits
                        c.a0 = i2f_bits(d2i_rn(d_mul(d_mul(Fx.inv[0], (double)ur), 1024.0)));
                        c.a3 = i2f_bits(d2i_rn(d_mul(d_mul(Fx.inv[3], (double)ur), 1024.0)));
                        c.a6 = 0.f;
                    } else {
                        float U = (float)(Fx.tlx + ur);
                        if (Fx.scale != 1.f) U = f_div(U, Fx.scale);
                        const float up = f_sub(U, Fx.t0);
                        c.a0 = f_mul(Fx.k[0], up); c.a3 = f_mul(Fx.k[3], up); c.a6 = f_mul(Fx.k[6], up);
                    }
                    c.u = (unsigned)u < (unsigned)Fx.w ? u : ~ur;   // >= 0: inside (u == ur); < 0: ~(reflected index)
                    t_col[i] = c;
                } else {
                    const int yy = i - PWS;
                    const int v = gx.ry + gx.py0 + yy - Fx.cy;
                    const int vr = refl(v, Fx.h, BORDER_REFL);
                    L0Row r;
                    if (aff) {
                        r.b1 = i2f_bits(d2i_rn(d_mul(d_add(d_mul(Fx.inv[1], (double)vr), Fx.inv[2]), 1024.0)));
                        r.b4 = i2f_bits(d2i_rn(d_mul(d_add(d_mul(Fx.inv[4], (double)vr), Fx.inv[5]), 1024.0)));
                        r.b7 = 0.f;
                    } else {
                        float V = (float)(Fx.tly + vr);
                        if (Fx.scale != 1.f) V = f_div(V, Fx.scale);
                        const float vp = f_sub(V, Fx.t1);
                        r.b1 = f_mul(Fx.k[1], vp); r.b4 = f_mul(Fx.k[4], vp); r.b7 = f_mul(Fx.k[7], vp);
                    }
                    r.v = (unsigned)v < (unsigned)Fx.h ? v : ~vr;
                    t_row[yy] = r;
                }
            }
        };
        // Source box of tile-frame fj into the H buffer (free at the call sites: see the layout comment). One thread issues
        // the TMA tile load (out-of-image parts are zero-filled and never read); the emulator copies.
        auto issue_box = [&](int fj) {
            if constexpr (HAS_BOX) {
                if (fj >= f_end) return;
                const Geo gx = s_geo[fj - f_begin];
                if (gx.skip || !gx.box) return;
#if DS_CUDA
                if (tid == 0) {
                    const char* tm = (const char*)p.tmaps + ((size_t)p.tile_frames[fj] * 2 + 1) * 128;   // slot [frame][1]: source, shared-memory box
                    fence_tensormap_acquire(tm);
                    fence_proxy_async();   // generic-proxy accesses of the buffer ended before the last barrier
                    mbar_expect_tx(s_bar, (uint32_t)BOX_BYTES);
                    tma_load_2d(smem + G0_BYTES + G1_BYTES, tm, gx.bx0, gx.by0, s_bar);
                }
#else
                const FrameDev& Fx = p.frames[p.tile_frames[fj]];
                uint32_t* box = (uint32_t*)(smem + G0_BYTES + G1_BYTES);
                for (int i = tid; i < BOXW * BOXH; i += NT) {
                    const int by = i / BOXW, bx = i - by * BOXW;
                    const int sx = gx.bx0 + bx, sy = gx.by0 + by;
                    box[i] = ((unsigned)sx < (unsigned)Fx.src_w && (unsigned)sy < (unsigned)Fx.src_h) ? Fx.src[(size_t)sy * Fx.src_pitch + sx] : 0u;
                }
#endif
            }
        };
        // what follows the barrier after a frame's warp loop: tables and L2 prefetch for the next frame of the tile
        auto prepare_next = [&](int fi) {
            if (fi + 1 >= f_end) return;
            const FrameDev& N = *(const FrameDev*)(smem + FDEV_OFF + ((fi + 1) & 1) * FDEV_BYTES);   // staged at the top of iteration fi
            build_tables(N, fi + 1);
#if DS_CUDA
            // L2 prefetch (TMA) of the source footprint of the next frame. One thread, fire and forget.
            if (p.tmaps != nullptr && tid == 0) {
                const Geo ng = s_geo[fi + 1 - f_begin];
                bool boxed = false;
                if constexpr (HAS_BOX) boxed = ng.box != 0;   // its footprint comes in as a box anyway
                if (!ng.skip && !boxed && N.kind == XF_PLANE) {
                    const int u_lo = ng.rx + ng.px0 - N.cx, u_hi = u_lo + ng.pw - 1, v_lo = ng.ry + ng.py0 - N.cy, v_hi = v_lo + ng.ph - 1;
                    const int ulo = imax(imin(u_lo, N.w - 1), 0), uhi = imax(imin(u_hi, N.w - 1), 0);
                    const int vlo = imax(imin(v_lo, N.h - 1), 0), vhi = imax(imin(v_hi, N.h - 1), 0);
                    float xmin = 3.0e38f, ymin = 3.0e38f;
                    for (int cidx = 0; cidx < 4; cidx++) {
                        float U = (float)(N.tlx + ((cidx & 1) ? uhi : ulo)), V = (float)(N.tly + ((cidx & 2) ? vhi : vlo));
                        if (N.scale != 1.f) { U = U / N.scale; V = V / N.scale; }
                        const float up = U - N.t0, vp = V - N.t1;
                        xmin = fminf(xmin, N.k[0] * up + N.k[1] * vp + N.k2one);
                        ymin = fminf(ymin, N.k[3] * up + N.k[4] * vp + N.k5one);
                    }
                    if (xmin > -1.0e6f && xmin < 1.0e6f && ymin > -1.0e6f && ymin < 1.0e6f) {
                        const char* tm = (const char*)p.tmaps + (size_t)p.tile_frames[fi + 1] * 2 * 128;   // slot [frame][0]: source, L2 prefetch box
                        fence_tensormap_acquire(tm);
                        tma_prefetch_l2_2d(tm, ((int)floorf(xmin) - 1) & ~3, (int)floorf(ymin) - 1);
                    }
                }
            }
#endif
        };
        if constexpr (LEVEL0) {
            if (f_begin < f_end) { build_tables(*(const FrameDev*)(smem + FDEV_OFF + (f_begin & 1) * FDEV_BYTES), f_begin); issue_box(f_begin); }
        }
        DS_SYNC();

        for (int fi = f_begin; fi < f_end; fi++) {
            if (fi + 1 < f_end) stage_frame(fi + 1);   // slot ((fi + 1) & 1) was last read in iteration fi - 1
            const FrameDev& F = *(const FrameDev*)(smem + FDEV_OFF + (fi & 1) * FDEV_BYTES);
            int m_and = 255, m_or = 0;   // level 0: AND / OR of the mask bytes; levels >= 1: 255 / 0 flags of (w == 1) / (w != 0)
            int known_votes = -1;        // >= 0: the uniformity of the mask is known without a block-wide vote (bit 0: all 255, bit 1: all 0)
            // The tile-frame geometry is read from shared memory twice - here and again after the warp loop - so that it
            // does not stay in registers across the loop: at 64 registers per thread that spilled ~25 values per thread and
            // tile-frame to local memory, and with 216 of the SM's 228 KB taken by shared memory the L1 does not hold a
            // CTA's stack (the spill traffic to L2 was as large as the source reads).
#define DS_TF_LOCALS                                                                     \
            const int rx = g.rx, ry = g.ry, rw = g.rw, rh = g.rh;                         \
            const int ax0 = g.ax0, ax1 = g.ax1, ay0 = g.ay0, ay1 = g.ay1;                 \
            const int n1x = rw >> 1, n1y = rh >> 1;                                       \
            const int jx0 = g.jx0, jx1 = g.jx1, jy0 = g.jy0, jy1 = g.jy1;                 \
            const int gx0 = g.gx0, gx1 = g.gx1, gy0 = g.gy0, gy1 = g.gy1;                 \
            const int px0 = g.px0, py0 = g.py0, pw = g.pw, ph = g.ph;                     \
            const int gw = gx1 - gx0 + 1, gh = gy1 - gy0 + 1;                             \
            const int jw = jx1 - jx0, jh = jy1 - jy0;                                     \
            const bool border = g.border != 0;                                            \
            (void)rx; (void)ry; (void)rw; (void)rh; (void)ax0; (void)ax1; (void)ay0; (void)ay1; (void)n1x; (void)n1y; \
            (void)jx0; (void)jx1; (void)jy0; (void)jy1; (void)gx0; (void)gy0; (void)px0; (void)py0; (void)pw; (void)ph; \
            (void)gw; (void)gh; (void)jw; (void)jh; (void)border;
            if (s_geo[fi - f_begin].skip) {       // block-uniform
                cp_async_wait_all();
                DS_SYNC();
                if constexpr (LEVEL0) { prepare_next(fi); issue_box(fi + 1); DS_SYNC(); }
                continue;
            }
            {   // ======== first half: tables + phase 1
            const Geo g = s_geo[fi - f_begin];
            DS_TF_LOCALS
            const bool affine = AFF && LEVEL0 && F.kind == XF_AFFINE;   // cv::warpAffine coordinates (integer tables)
            const bool proj = !affine && !(F.k[6] == 0.f && F.k[7] == 0.f && F.k8one == 1.f);
            if constexpr (LEVEL0) {
            L0Col* const s_col = col_buf(fi);   // built during the previous frame (or before the loop)
            L0Row* const s_row = row_buf(fi);

            // ---- phase 1: inverse warp of the needed region into s_g0 (b | g<<8 | r<<16 | mask<<24)
            // Frame fields are copied to registers first: F lives in global memory and would otherwise be
            // re-read around every shared-memory store.
            {
                const uint32_t* const src = F.src;
                const int pitch = F.src_pitch, sw = F.src_w, sh = F.src_h;
                const float k2 = F.k2one, k5 = F.k5one, k8 = F.k8one;
                const uint8_t* const seam = F.seam;
                const int seam_pitch = F.seam_pitch;
                const bool has_gain = F.any_gain != 0;
                const bool bconst = F.border == BORDER_CONST;
                // Fast tile-frames: every tap of the needed region is an in-bounds inlier. The region's bbox
                // indices (mirrored into the bbox where it lies in the gap of the feed ROI, BORDER_REFLECT of
                // copyMakeBorder) cover one contiguous interval per axis, and x(u, v) is monotone in u and in v
                // even in float arithmetic (a chain of monotone roundings), so the values at the interval ends
                // bound every pixel: no border handling, no cvRound patch, nearest mask = "inside the bbox".
                // `interior`: the region is inside the bbox as well, so the mask is 255 everywhere.
                bool inbounds = false, interior = false;
                if constexpr (AFF) {
                if (!proj) {
                    const int u_lo = rx + px0 - F.cx, u_hi = u_lo + pw - 1, v_lo = ry + py0 - F.cy, v_hi = v_lo + ph - 1;
                    int umn = 0, umx = 0, vmn = 0, vmx = 0;
                    if (affine) {
                        if (fold_range(u_lo, u_hi, F.w, umn, umx) && fold_range(v_lo, v_hi, F.h, vmn, vmx)) {
                            int xmn = 0x7fffffff, xmx = (int)0x80000000, ymn = 0x7fffffff, ymx = (int)0x80000000;
                            DS_UNROLL
                            for (int e = 0; e < 4; e++) {
                                const double uu = (double)((e & 1) ? umx : umn), vv = (double)((e >> 1) ? vmx : vmn);
                                const int X = (d2i_rn(d_mul(d_add(d_mul(F.inv[1], vv), F.inv[2]), 1024.0)) + 16 + d2i_rn(d_mul(d_mul(F.inv[0], uu), 1024.0))) >> 5;
                                const int Y = (d2i_rn(d_mul(d_add(d_mul(F.inv[4], vv), F.inv[5]), 1024.0)) + 16 + d2i_rn(d_mul(d_mul(F.inv[3], uu), 1024.0))) >> 5;
                                xmn = imin(xmn, X); xmx = imax(xmx, X); ymn = imin(ymn, Y); ymx = imax(ymx, Y);
                            }
                            inbounds = (xmn >> 5) >= 1 && (xmx >> 5) <= sw - 3 && (ymn >> 5) >= 1 && (ymx >> 5) <= sh - 3;
                            interior = inbounds && u_lo >= 0 && u_hi < F.w && v_lo >= 0 && v_hi < F.h;
                            if (!interior && !(p.flags & 1)) inbounds = false;
                        }
                    } else if (fold_range(u_lo, u_hi, F.w, umn, umx) && fold_range(v_lo, v_hi, F.h, vmn, vmx)) {
                        float ca0[2], ca3[2], rb1[2], rb4[2];
                        DS_UNROLL
                        for (int e = 0; e < 2; e++) {
                            float U = (float)(F.tlx + (e ? umx : umn)), V = (float)(F.tly + (e ? vmx : vmn));
                            if (F.scale != 1.f) { U = f_div(U, F.scale); V = f_div(V, F.scale); }
                            const float up = f_sub(U, F.t0), vp = f_sub(V, F.t1);
                            ca0[e] = f_mul(F.k[0], up); ca3[e] = f_mul(F.k[3], up);
                            rb1[e] = f_mul(F.k[1], vp); rb4[e] = f_mul(F.k[4], vp);
                        }
                        float xmn = 3.0e38f, xmx = -3.0e38f, ymn = 3.0e38f, ymx = -3.0e38f;
                        DS_UNROLL
                        for (int e = 0; e < 4; e++) {
                            const float xx_ = f_add(f_add(ca0[e & 1], rb1[e >> 1]), k2), yy_ = f_add(f_add(ca3[e & 1], rb4[e >> 1]), k5);
                            xmn = fminf(xmn, xx_); xmx = fmaxf(xmx, xx_); ymn = fminf(ymn, yy_); ymx = fmaxf(ymx, yy_);
                        }
                        inbounds = xmn >= 1.f && xmx <= (float)(sw - 3) && ymn >= 1.f && ymx <= (float)(sh - 3);
                        interior = inbounds && u_lo >= 0 && u_hi < F.w && v_lo >= 0 && v_hi < F.h;
                        if (!interior && !(p.flags & 1)) inbounds = false;
                    }
                }
                }
                bool v2_done = false;
                if constexpr (!AFF) {
                // ---- v2 warp loop (plane maps without perspective terms, no per-pixel inputs). One thread per region
                // column: the column's map terms stay in registers, the row's terms come from one broadcast shared
                // load, so a pixel costs four adds and two fused multiply-adds (cvRound(32 x) as the mantissa of
                // x * 32 + 1.5 * 2^23: 32 x is exact, the sum rounds once, half to even - cvRound without the XU pipe),
                // an address, four taps, and the 15-bit bilinear as two-way dot products with 16-bit 2-D weights
                // ((32-ax)(32-ay) | ax(32-ay) << 16 etc.: dp2a) instead of byte dot products plus a vertical pass.
                //   MODE 0  region inside the bbox, taps inside the source: mask 255
                //   MODE 1  region reaches into the gap of the feed ROI (mirrored bbox indices): mask 0 there
                //   MODE 2  taps may leave the source (BORDER_REFLECT, one reflection), nearest mask per pixel
                // Tap addresses in modes 0 / 1 are formed from the biased mantissas directly: with C = bias >> 5,
                // t = (sy + C) * pitch + (sx + C) mod 2^32 = off + K, and base' = base + 4 * (off_lo - u_lo) makes
                // base' + 4 * t the tap address as long as u_lo + span does not wrap (checked here, per tile-frame).
                constexpr int CB = DS_RND_BIAS >> 5;
                const bool v2 = g.mode >= 0;
                const char* const basep = (const char*)src + g.boff;
                interior = g.mode == 0; inbounds = g.mode == 0 || g.mode == 1;
                const bool gap_all0 = g.gap_all0 != 0;
                if (v2) {
                    const SAddr a_col = s_addr(s_col), a_row = s_addr(s_row), a_g0 = s_addr(s_g0);
                    constexpr int NTV = 512, RG = NTV / 64;          // 64 columns x 8 row groups
                    static_assert((PHM + RG - 1) / RG == 9 && PWS - 64 == 8 && 7 * PHM <= NTV, "v2 warp loop layout");
                    struct PX { int bx, by; uint32_t p00, p01, p10, p11; int aux; };
                    // modes 0 / 1 with the footprint in the shared-memory box: the biased tap offset is taken against the
                    // box origin (all 32-bit, wrap-around is harmless: the final shared address is exact)
                    const bool boxed = HAS_BOX && g.box != 0;
                    // (the bias comes in as a kernel parameter: as a compile-time constant ptxas carries its part of the address -
                    // which does not fit a load's immediate field - to every single tap and adds it there, one instruction per tap)
                    const int CBr = p.rnd_bias >> 5;
                    const uint32_t boxK = (uint32_t)(CBr + g.by0) * (uint32_t)BOXW + (uint32_t)(CBr + g.bx0);
#if DS_CUDA
                    const SAddr a_boxb = s_addr(smem + G0_BYTES + G1_BYTES) - 4u * boxK;   // shared addresses are 32-bit: the bias folds in
                    auto box_addr = [&](uint32_t t) { return a_boxb + 4u * t; };
#else
                    unsigned char* const a_box0 = smem + G0_BYTES + G1_BYTES;
                    auto box_addr = [&](uint32_t t) { return a_box0 + (ptrdiff_t)(int32_t)(4u * (t - boxK)); };
#endif
                    if (boxed) {
#if DS_CUDA
                        mbar_wait(s_bar, box_phase);
#endif
                        box_phase ^= 1u;
                    }
                    auto fetch = [&](auto mode_tag, float ca0, float ca3, int cu, int r, PX& q) {
                        constexpr int MODE = decltype(mode_tag)::value / 4;
                        constexpr bool BOXED = (decltype(mode_tag)::value & 1) != 0;
                        float rb1, rb4; int rv = 0;
                        if constexpr (MODE == 0) lds_f2(a_row + r * 16, rb1, rb4);
                        else { uint32_t w0, w1, w2, w3; lds_u4(a_row + r * 16, w0, w1, w2, w3); rb1 = i2f_bits((int)w0); rb4 = i2f_bits((int)w1); rv = (int)w3; }
                        const float x = f_add(f_add(ca0, rb1), k2), y = f_add(f_add(ca3, rb4), k5);
                        q.bx = rnd32_bits(x); q.by = rnd32_bits(y);
                        if constexpr (MODE != 2 && BOXED) {
                            const uint32_t t = (uint32_t)(q.by >> 5) * (uint32_t)BOXW + (uint32_t)(q.bx >> 5);
                            const SAddr a = box_addr(t);
                            q.p00 = lds_u1(a); q.p01 = lds_u1(a + 4); q.p10 = lds_u1(a + BOXW * 4); q.p11 = lds_u1(a + BOXW * 4 + 4);
                            q.aux = rv;
                        } else if constexpr (MODE != 2) {
                            const uint32_t t = (uint32_t)(q.by >> 5) * (uint32_t)pitch + (uint32_t)(q.bx >> 5);
                            const uint32_t* r0 = (const uint32_t*)(basep + 4ull * (unsigned long long)t);
                            const uint32_t* r1 = (const uint32_t*)(basep + 4ull * (unsigned long long)(t + (uint32_t)pitch));
                            q.p00 = ld_ro(r0); q.p01 = ld_ro(r0 + 1); q.p10 = ld_ro(r1); q.p11 = ld_ro(r1 + 1);
                            q.aux = rv;
                        } else {
                            const int sx = (q.bx >> 5) - CB, sy = (q.by >> 5) - CB;
                            // one reflection: p < 0 -> -p - 1, p >= n -> 2n - 1 - p
                            auto refl1 = [](int p_, int n_) { const int q_ = p_ ^ (p_ >> 31); return imin(q_, 2 * n_ - 1 - q_); };
                            const int x0 = refl1(sx, sw), x1 = refl1(sx + 1, sw), y0 = refl1(sy, sh), y1 = refl1(sy + 1, sh);
                            const uint32_t* ra = src + (size_t)y0 * pitch; const uint32_t* rb = src + (size_t)y1 * pitch;
                            q.p00 = ld_ro(ra + x0); q.p01 = ld_ro(ra + x1); q.p10 = ld_ro(rb + x0); q.p11 = ld_ro(rb + x1);
                            const int nx = rnd1_bits(x) - DS_RND_BIAS, ny = rnd1_bits(y) - DS_RND_BIAS;
                            q.aux = ((cu | rv) >= 0 && (unsigned)nx < (unsigned)sw && (unsigned)ny < (unsigned)sh) ? 255 : 0;
                        }
                    };
                    auto finish = [&](auto mode_tag, const PX& q, SAddr dst, bool store, bool count, uint32_t cmask) {
                        constexpr int MODE = decltype(mode_tag)::value / 4;
                        const uint32_t ax = (uint32_t)q.bx & 31u, ay = (uint32_t)q.by & 31u;
                        const uint32_t wxp = ax * 65535u + 32u;                     // (32 - ax) | ax << 16
                        const uint32_t wbot = wxp * ay, wtop = wxp * 32u - wbot;    // rows weighted by ay / 32 - ay
                        const uint32_t t0 = byte_perm(q.p00, q.p01, 0x5140), t0r = byte_perm(q.p00, q.p01, 0x6262);
                        const uint32_t t1 = byte_perm(q.p10, q.p11, 0x5140), t1r = byte_perm(q.p10, q.p11, 0x6262);
                        // (sum + 2^14) >> 15 of OpenCV's weights * 32 == (sum + 512) >> 10
                        const uint32_t sb = dot2lo(wbot, t1, dot2lo(wtop, t0, 512u));
                        const uint32_t sg = dot2hi(wbot, t1, dot2hi(wtop, t0, 512u));
                        // the mask byte rides in the red accumulator: (s + 512 + (m << 18)) >> 10 == r + (m << 8); every sum,
                        // shifted left by 6, holds its 8-bit result in byte 2 (and red's mask in byte 3): two byte permutes pack
                        uint32_t minit;
                        if constexpr (MODE == 0) minit = 512u + (0xFFu << 18);
                        else if constexpr (MODE == 1) minit = 512u + ((cmask & ~(uint32_t)(q.aux >> 31)) >> 6);
                        else minit = 512u + ((uint32_t)q.aux << 18);
                        const uint32_t sr = dot2lo(wbot, t1r, dot2lo(wtop, t0r, minit));
                        const uint32_t val = byte_perm(byte_perm(sb << 6, sg << 6, 0x0062), sr << 6, 0x7610);
                        if constexpr (MODE == 2) { if (count) { m_and &= q.aux; m_or |= q.aux; } }
                        if (store) sts_u1(dst, val);
                    };
                    auto run_v2 = [&](auto mode_tag) {
                        for (int vt = tid; vt < NTV; vt += NT) {   // a single trip on the GPU (NT == NTV)
                            {
                                const int col = vt & 63, rg = vt >> 6;
                                uint32_t c0, c1, c2, c3;
                                lds_u4(a_col + col * 16, c0, c1, c2, c3);
                                const float ca0 = i2f_bits((int)c0), ca3 = i2f_bits((int)c1);
                                const int cu = (int)c3;
                                const uint32_t cmask = cu >= 0 ? 0xff000000u : 0u;
                                const bool count = col < pw;
                                const SAddr d0 = a_g0 + (rg * PWS + col) * 4;
                                // rows rg + 8k, k = 0..8, software-pipelined in two register sets of two rows: the taps of
                                // the next set are requested before this one is interpolated (occupancy is 2 CTAs / SM;
                                // other warps alone do not cover an L2 round trip)
                                PX A[2], B[2];
                                auto row_of = [&](int k) { return imin(rg + RG * k, ph - 1); };
                                auto fin = [&](const PX& q, int k) {
                                    const bool live = rg + RG * k < ph;
                                    finish(mode_tag, q, d0 + k * (RG * PWS * 4), live, live && count, cmask);
                                };
                                fetch(mode_tag, ca0, ca3, cu, row_of(0), A[0]); fetch(mode_tag, ca0, ca3, cu, row_of(1), A[1]);
                                fetch(mode_tag, ca0, ca3, cu, row_of(2), B[0]); fetch(mode_tag, ca0, ca3, cu, row_of(3), B[1]);
                                fin(A[0], 0); fin(A[1], 1);
                                fetch(mode_tag, ca0, ca3, cu, row_of(4), A[0]); fetch(mode_tag, ca0, ca3, cu, row_of(5), A[1]);
                                fin(B[0], 2); fin(B[1], 3);
                                fetch(mode_tag, ca0, ca3, cu, row_of(6), B[0]); fetch(mode_tag, ca0, ca3, cu, row_of(7), B[1]);
                                fin(A[0], 4); fin(A[1], 5);
                                // the tenth pixel of the thread: one of the 7 x 71 pixels of the columns beyond the 64th
                                // (497 items over the 512 threads), in flight together with row 8
                                const int er = vt / 7, ec = 64 + (vt - er * 7);
                                const bool elive = er < ph && ec < pw;
                                uint32_t e0, e1, e2, e3;
                                lds_u4(a_col + ec * 16, e0, e1, e2, e3);
                                fetch(mode_tag, ca0, ca3, cu, row_of(8), A[0]);
                                fetch(mode_tag, i2f_bits((int)e0), i2f_bits((int)e1), (int)e3, imin(er, ph - 1), A[1]);
                                fin(B[0], 6); fin(B[1], 7);
                                fin(A[0], 8);
                                finish(mode_tag, A[1], a_g0 + (imin(er, ph - 1) * PWS + ec) * 4, elive, elive, (int)e3 >= 0 ? 0xff000000u : 0u);
                            }
                        }
                    };
                    // The common tile-frame - mode 0, footprint in the box, the whole 71 x 71 region - with everything that is
                    // known at compile time folded into the instructions: rows rg + 8 k live for k < 8 (no clamps, no per-pixel
                    // predicates), row-table loads and stores at immediate offsets from one base register each, the four taps of
                    // a pixel off one address register.
                    auto run_full0 = [&](auto mask_tag) {   // mask 255 (inside the bbox) or 0 (wholly in the gap of the feed ROI)
                        constexpr uint32_t MINIT = 512u + ((uint32_t)decltype(mask_tag)::value << 18);
                        for (int vt = tid; vt < NTV; vt += NT) {
                            const int col = vt & 63, rg = vt >> 6;
                            uint32_t c0, c1, c2, c3;
                            lds_u4(a_col + col * 16, c0, c1, c2, c3);
                            const float ca0 = i2f_bits((int)c0), ca3 = i2f_bits((int)c1);
                            const SAddr arow = a_row + rg * 16, d0 = a_g0 + (rg * PWS + col) * 4;
                            auto taps = [&](float fa0, float fa3, float rb1, float rb4, PX& q) {
                                const float x = f_add(f_add(fa0, rb1), k2), y = f_add(f_add(fa3, rb4), k5);
                                q.bx = rnd32_bits(x); q.by = rnd32_bits(y);
                                const uint32_t t = (uint32_t)(q.by >> 5) * (uint32_t)BOXW + (uint32_t)(q.bx >> 5);
                                lds_tap4<BOXW * 4>(box_addr(t), q.p00, q.p01, q.p10, q.p11);
                            };
                            auto fetch_k = [&](auto ktag, PX& q) {
                                float rb1, rb4;
                                lds_f2_o<decltype(ktag)::value * RG * 16>(arow, rb1, rb4);
                                taps(ca0, ca3, rb1, rb4, q);
                            };
                            auto value = [&](const PX& q) {
                                const uint32_t ax = (uint32_t)q.bx & 31u, ay = (uint32_t)q.by & 31u;
                                const uint32_t wxp = ax * 65535u + 32u;                     // (32 - ax) | ax << 16
                                const uint32_t wbot = wxp * ay, wtop = wxp * 32u - wbot;    // rows weighted by ay / 32 - ay
                                const uint32_t t0 = byte_perm(q.p00, q.p01, 0x5140), t0r = byte_perm(q.p00, q.p01, 0x6262);
                                const uint32_t t1 = byte_perm(q.p10, q.p11, 0x5140), t1r = byte_perm(q.p10, q.p11, 0x6262);
                                const uint32_t sb = dot2lo(wbot, t1, dot2lo(wtop, t0, 512u));
                                const uint32_t sg = dot2hi(wbot, t1, dot2hi(wtop, t0, 512u));
                                const uint32_t sr = dot2lo(wbot, t1r, dot2lo(wtop, t0r, MINIT));   // the mask byte rides in red
                                return byte_perm(byte_perm(sb << 6, sg << 6, 0x0062), sr << 6, 0x7610);
                            };
                            auto fin_k = [&](auto ktag, const PX& q) { sts_u1_o<decltype(ktag)::value * RG * PWS * 4>(d0, value(q)); };
                            PX A[2], B[2];
                            fetch_k(IntTag<0>(), A[0]); fetch_k(IntTag<1>(), A[1]);
                            fetch_k(IntTag<2>(), B[0]); fetch_k(IntTag<3>(), B[1]);
                            fin_k(IntTag<0>(), A[0]); fin_k(IntTag<1>(), A[1]);
                            fetch_k(IntTag<4>(), A[0]); fetch_k(IntTag<5>(), A[1]);
                            fin_k(IntTag<2>(), B[0]); fin_k(IntTag<3>(), B[1]);
                            fetch_k(IntTag<6>(), B[0]); fetch_k(IntTag<7>(), B[1]);
                            fin_k(IntTag<4>(), A[0]); fin_k(IntTag<5>(), A[1]);
                            // row rg + 64 (live for rg < 7) and the thread's one pixel of the columns beyond the 64th (497 items)
                            {
                                float rb1, rb4;
                                lds_f2(a_row + imin(rg + RG * 8, PHM - 1) * 16, rb1, rb4);
                                taps(ca0, ca3, rb1, rb4, A[0]);
                                const int er_ = vt / 7, ec = 64 + (vt - er_ * 7), er = imin(er_, PHM - 1);
                                uint32_t e0, e1, e2, e3;
                                lds_u4(a_col + ec * 16, e0, e1, e2, e3);
                                lds_f2(a_row + er * 16, rb1, rb4);
                                taps(i2f_bits((int)e0), i2f_bits((int)e1), rb1, rb4, A[1]);
                                fin_k(IntTag<6>(), B[0]); fin_k(IntTag<7>(), B[1]);
                                if (rg < PHM - RG * 8) sts_u1_o<8 * RG * PWS * 4>(d0, value(A[0]));
                                if (er_ < PHM) sts_u1(a_g0 + (er * PWS + ec) * 4, value(A[1]));
                            }
                        }
                    };
#if !DS_CUDA
                    ds_emu_count(interior ? 0 : (inbounds ? 1 : 2));
                    if (boxed) ds_emu_count(5);
#endif
                    // tag = 4 * mode + (taps from the shared-memory box)
                    if (interior) {
                        if (boxed && ph == PHM && pw == PHM) run_full0(IntTag<255>());
                        else if (boxed) run_v2(IntTag<1>()); else run_v2(IntTag<0>());
                        m_or = 255; known_votes = 1;                 // uniform 255
                    } else if (inbounds) {
                        if (boxed && gap_all0 && ph == PHM && pw == PHM && !(p.flags & 8)) run_full0(IntTag<0>());   // mirror image, mask 0 throughout
                        else if (boxed) run_v2(IntTag<5>()); else run_v2(IntTag<4>());
                        known_votes = gap_all0 ? 2 : 0;              // not inside the bbox: never uniform 255
                    } else {
                        run_v2(IntTag<8>());
                    }
                    v2_done = true;
                }
                }
#if !DS_CUDA
                if (!v2_done) ds_emu_count(AFF && inbounds ? 3 : 4);
#endif
                if (v2_done) {
                } else if (AFF && inbounds) {
                    // Software-pipelined in two register sets of UB pixels per thread: the taps of stage k+1 are
                    // requested before stage k is interpolated, so the L1 / L2 latency of the gathers hides behind
                    // the arithmetic of the same warp (occupancy is 2 CTAs / SM; other warps alone do not cover it).
                    // Branch-free: stages are block-uniform, slots past the end of the region repeat its last
                    // pixel and do not store; the column table is valid over the whole pitch, so padding
                    // columns compute a harmless value nobody reads.
                    constexpr int UB = 2, STEP = UB * NT;
                    const int npx = PWS * ph, last = npx - 1;
                    const int nst = (npx + STEP - 1) / STEP;
                    const SAddr a_col = s_addr(s_col), a_row = s_addr(s_row), a_g0 = s_addr(s_g0);
                    // mk / gm: mask byte and gain-map value of the pixel, requested together with the taps (general variant)
                    struct Taps { int ix[UB], iy[UB]; uint32_t p00[UB], p01[UB], p10[UB], p11[UB]; int mk[UB]; float gm[UB]; };
                    const float* const gainmap = F.gainmap;
                    const int gainmap_pitch = F.gainmap_pitch;
                    auto fetch = [&](int i0, Taps& t, auto aff, auto masked) {   // integral_constant<bool> tags: warpAffine integer coordinates / per-pixel mask
                        DS_UNROLL
                        for (int b = 0; b < UB; b++) {
                            const int i = imin(i0 + b * NT, last);
                            const int yy = i / PWS, xx = i - yy * PWS;
                            if constexpr (decltype(aff)::value) {
                                uint32_t ad, bd, xr, yr;
                                lds_u2(a_col + xx * 16, ad, bd);
                                lds_u2(a_row + yy * 16, xr, yr);
                                t.ix[b] = ((int)xr + 16 + (int)ad) >> 5; t.iy[b] = ((int)yr + 16 + (int)bd) >> 5;
                            } else {
                                float ca0, ca3, rb1, rb4;
                                lds_f2(a_col + xx * 16, ca0, ca3);
                                lds_f2(a_row + yy * 16, rb1, rb4);
                                const float x = f_add(f_add(ca0, rb1), k2);
                                const float y = f_add(f_add(ca3, rb4), k5);
#if DS_CUDA
                                t.ix[b] = __float2int_rn(f_mul(x, 32.f)); t.iy[b] = __float2int_rn(f_mul(y, 32.f));
#else
                                t.ix[b] = f2i_rn(f_mul(x, 32.f)); t.iy[b] = f2i_rn(f_mul(y, 32.f));
#endif
                            }
                            const uint32_t* r0 = src + ((t.iy[b] >> 5) * pitch + (t.ix[b] >> 5));
                            t.p00[b] = ld_ro(r0); t.p01[b] = ld_ro(r0 + 1); t.p10[b] = ld_ro(r0 + pitch); t.p11[b] = ld_ro(r0 + pitch + 1);
                            if constexpr (AFF && decltype(masked)::value) {
                                // mask byte (seam mask inside the bbox, 0 in the gap) and gain-map value at the (mirrored)
                                // bbox position, in flight with the taps
                                const int cu = (int)lds_u1(a_col + xx * 16 + 12), rv = (int)lds_u1(a_row + yy * 16 + 12);
                                const int inside = (cu | rv) >= 0;
                                t.mk[b] = inside ? 255 : 0;
                                if (seam && inside) t.mk[b] = (int)ld_ro(seam + (size_t)rv * seam_pitch + cu);
                                t.gm[b] = 1.f;
                                if (gainmap) t.gm[b] = ld_ro(gainmap + (size_t)(rv >= 0 ? rv : ~rv) * gainmap_pitch + (cu >= 0 ? cu : ~cu));
                            }
                        }
                    };
                    auto finish = [&](int i0, const Taps& t, auto masked) {   // masked: integral_constant<bool>
                        DS_UNROLL
                        for (int b = 0; b < UB; b++) {
                            const int ax = t.ix[b] & 31, ay = t.iy[b] & 31;
                            const uint32_t wb = (uint32_t)(32 - ax) | ((uint32_t)ax << 8), wg = wb << 16;
                            const uint32_t t0 = byte_perm(t.p00[b], t.p01[b], 0x5140), t0r = byte_perm(t.p00[b], t.p01[b], 0x6262);
                            const uint32_t t1 = byte_perm(t.p10[b], t.p11[b], 0x5140), t1r = byte_perm(t.p10[b], t.p11[b], 0x6262);
                            const int wy1 = ay, wy0 = 32 - ay;
                            int ob = (dot4u(t0, wb, 0) * wy0 + dot4u(t1, wb, 0) * wy1 + 512) >> 10;
                            int og = (dot4u(t0, wg, 0) * wy0 + dot4u(t1, wg, 0) * wy1 + 512) >> 10;
                            int orr = (dot4u(t0r, wb, 0) * wy0 + dot4u(t1r, wb, 0) * wy1 + 512) >> 10;
                            const int i = i0 + b * NT;
                            uint32_t mbyte = 0xff000000u;
                            if constexpr (decltype(masked)::value) {
                                // the general form: a pixel in the gap of the feed ROI is a mirror image whose mask
                                // (CONSTANT 0 border) is 0; inside the bbox the seam mask, if any, is the mask; the
                                // gain map is read at the (mirrored) bbox position
                                const int ic = imin(i, last);
                                const int yy = ic / PWS, xx = ic - yy * PWS;
                                int m;
                                if constexpr (AFF) {
                                    if (has_gain) apply_gains_with(F, ob, og, orr, t.gm[b]);
                                    m = t.mk[b];
                                } else {
                                    const int cu = (int)lds_u1(a_col + xx * 16 + 12), rv = (int)lds_u1(a_row + yy * 16 + 12);
                                    m = (cu | rv) >= 0 ? 255 : 0;
                                    if (has_gain) apply_gains(F, ob, og, orr, 0, 0);   // no gain map, no seam mask here
                                }
                                mbyte = (uint32_t)m << 24;
                                if (i < npx && xx < pw) { m_and &= m; m_or |= m; }
                            } else {
                                if (has_gain) apply_gains(F, ob, og, orr, 0, 0);   // no gain map in this variant
                            }
                            if (i < npx) sts_u1(a_g0 + i * 4, (uint32_t)ob | ((uint32_t)og << 8) | ((uint32_t)orr << 16) | mbyte);
                        }
                    };
                    auto pipeline = [&](auto masked, auto aff) {
                        Taps A, B;
                        fetch(tid, A, aff, masked);
                        int st = 0;
                        for (; st + 2 <= nst; st += 2) {
                            fetch(tid + (st + 1) * STEP, B, aff, masked);
                            finish(tid + st * STEP, A, masked);
                            fetch(tid + (st + 2) * STEP, A, aff, masked);   // past the end on the last trip: clamped, unused
                            finish(tid + (st + 1) * STEP, B, masked);
                        }
                        if (st < nst) finish(tid + st * STEP, A, masked);
                    };
                    const bool plain = interior && !seam && !F.gainmap;   // mask 255 everywhere, no per-pixel inputs
                    bool done = false;
                    if constexpr (AFF) {
                        if (affine) {
                            if (plain) pipeline(BoolTag<false>(), BoolTag<true>()); else pipeline(BoolTag<true>(), BoolTag<true>());
                            done = true;
                        }
                    }
                    if (!done) {
                        if (plain) pipeline(BoolTag<false>(), BoolTag<false>()); else pipeline(BoolTag<true>(), BoolTag<false>());
                    }
                    if (plain) {
                        m_or = 255;   // m_and stays 255: the mask is uniform 255
                        known_votes = 1;
                    }
                } else
                for (int i = tid; i < PWS * ph; i += NT) {
                    const int yy = i / PWS, xx = i - yy * PWS;
                    if (xx >= pw) continue;
                    const L0Col c = s_col[xx];
                    const L0Row r = s_row[yy];
                    bool okx = true, oky = true;
                    int ix, iy, nx, ny;
                    if (AFF && F.kind == XF_HOMOGRAPHY) {
                        // cv::warpPerspective coordinates (A7): double arithmetic per pixel, at the (mirrored) bbox position;
                        // everything after the coordinates - taps, pyramids, accumulation - is the fast kernel's
                        const Coord hc = eval_coord(F, c.u >= 0 ? c.u : ~c.u, r.v >= 0 ? r.v : ~r.v);
                        ix = hc.sx * 32 + hc.ax; iy = hc.sy * 32 + hc.ay;   // sat16(ix >> 5) == hc.sx again below
                        nx = hc.m ? 0 : -1; ny = 0;                          // the nearest mask comes with the coordinate
                    } else if (affine) {
                        const int xr = f2i_bits(r.b1), yr = f2i_bits(r.b4), ad = f2i_bits(c.a0), bd = f2i_bits(c.a3);
                        ix = (xr + 16 + ad) >> 5; iy = (yr + 16 + bd) >> 5;
                        nx = sat16i((xr + 512 + ad) >> 10); ny = sat16i((yr + 512 + bd) >> 10);
                    } else {
                    float x = f_add(f_add(c.a0, r.b1), k2);
                    float y = f_add(f_add(c.a3, r.b4), k5);
                    if (proj) {
                        const float z = f_add(f_add(c.a6, r.b7), k8);
                        if (z != 1.f) { x = f_div(x, z); y = f_div(y, z); }
                    }
                    // cvRound: |32 x| >= 2^31 or NaN gives INT_MIN on the oracle's x86 (f2i_rn); below that the
                    // plain conversion is identical. For the nearest mask the patch is not needed: a
                    // saturated coordinate is out of the source either way, NaN is excluded through `ok`.
                    okx = x < 67108864.f; oky = y < 67108864.f;
#if DS_CUDA
                    ix = __float2int_rn(f_mul(x, 32.f)); iy = __float2int_rn(f_mul(y, 32.f));
                    nx = __float2int_rn(x); ny = __float2int_rn(y);
#else
                    ix = f2i_rn(f_mul(x, 32.f)); iy = f2i_rn(f_mul(y, 32.f));
                    nx = okx ? f2i_rn(x) : 0; ny = oky ? f2i_rn(y) : 0;
#endif
                    ix = okx ? ix : (int)0x80000000;
                    iy = oky ? iy : (int)0x80000000;
                    }
                    const int sx = sat16i(ix >> 5), sy = sat16i(iy >> 5);
                    const int ax = ix & 31, ay = iy & 31;
                    int m = ((c.u | r.v) >= 0 && okx && oky && (unsigned)nx < (unsigned)sw && (unsigned)ny < (unsigned)sh) ? 255 : 0;
                    if (seam && (c.u | r.v) >= 0) m &= (int)ld_ro(seam + (size_t)r.v * seam_pitch + c.u);
                    uint32_t p00, p01, p10, p11;
                    if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
                        const uint32_t* r0 = src + (sy * pitch + sx);
                        p00 = ld_ro(r0); p01 = ld_ro(r0 + 1);
                        p10 = ld_ro(r0 + pitch); p11 = ld_ro(r0 + pitch + 1);
                    } else if (bconst) {
                        const int x0 = (unsigned)sx < (unsigned)sw ? sx : -1;
                        const int x1 = (unsigned)(sx + 1) < (unsigned)sw ? sx + 1 : -1;
                        const int y0 = (unsigned)sy < (unsigned)sh ? sy : -1;
                        const int y1 = (unsigned)(sy + 1) < (unsigned)sh ? sy + 1 : -1;
                        p00 = src_tap(F, x0, y0); p01 = src_tap(F, x1, y0);
                        p10 = src_tap(F, x0, y1); p11 = src_tap(F, x1, y1);
                    } else {
                        const int x0 = refl(sx, sw, BORDER_REFL), x1 = refl(sx + 1, sw, BORDER_REFL);
                        const int y0 = refl(sy, sh, BORDER_REFL), y1 = refl(sy + 1, sh, BORDER_REFL);
                        p00 = src_tap(F, x0, y0); p01 = src_tap(F, x1, y0);
                        p10 = src_tap(F, x0, y1); p11 = src_tap(F, x1, y1);
                    }
                    // horizontal: bytes (B0,B1,G0,G1) . (wx0,wx1,0,0) etc.; vertical in 32 bit; (v + 512) >> 10
                    const uint32_t wb = (uint32_t)(32 - ax) | ((uint32_t)ax << 8), wg = wb << 16;
                    const uint32_t t0 = byte_perm(p00, p01, 0x5140), t0r = byte_perm(p00, p01, 0x6262);
                    const uint32_t t1 = byte_perm(p10, p11, 0x5140), t1r = byte_perm(p10, p11, 0x6262);
                    const int wy1 = ay, wy0 = 32 - ay;
                    int ob = (dot4u(t0, wb, 0) * wy0 + dot4u(t1, wb, 0) * wy1 + 512) >> 10;
                    int og = (dot4u(t0, wg, 0) * wy0 + dot4u(t1, wg, 0) * wy1 + 512) >> 10;
                    int orr = (dot4u(t0r, wb, 0) * wy0 + dot4u(t1r, wb, 0) * wy1 + 512) >> 10;
                    if (has_gain) apply_gains(F, ob, og, orr, c.u >= 0 ? c.u : ~c.u, r.v >= 0 ? r.v : ~r.v);
                    s_g0[i] = (uint32_t)ob | ((uint32_t)og << 8) | ((uint32_t)orr << 16) | ((uint32_t)m << 24);
                    m_and &= m;
                    m_or |= m;
                }
            }
            }
            }   // ======== end of the first half
            {   // ======== second half: votes, pyrDown, Laplacian + accumulate
            Geo g;
            {
                uint32_t gwords[GEO_BYTES / 4];
                const SAddr ga = s_addr(s_geo + (fi - f_begin));
                DS_UNROLL
                for (int k = 0; k < GEO_BYTES / 16; k++) lds_u4(ga + k * 16, gwords[4 * k], gwords[4 * k + 1], gwords[4 * k + 2], gwords[4 * k + 3]);
                memcpy(&g, gwords, sizeof(g));
            }
            DS_TF_LOCALS
            uint32_t* const G1out = (uint32_t*)F.G[l + 1];
            float* const W1out = F.W[l + 1];
            const int op1 = F.gp[l + 1];   // row pitch of the level-(l+1) arrays
            const int all255 = (m_and == 255), all0 = (m_or == 0);
            int uni255, uni0;
            cp_async_wait_all();   // the next frame's descriptor (requested at the top of the iteration) is read after this barrier
            if (known_votes >= 0) {
                DS_SYNC();
                uni255 = (known_votes & 1) && (c255 == 1.f); uni0 = (known_votes & 2) != 0;
            } else {
                // one barrier for both votes: bit 0 = some weight is not 1, bit 1 = some weight is not 0
                const int bits = block_or_bits((all255 ? 0 : 1) | (all0 ? 0 : 2), s_vote);
                uni255 = !(bits & 1) && (!LEVEL0 || c255 == 1.f);   // every weight of the needed region is exactly 1
                uni0 = !(bits & 2);
            }

            if constexpr (LEVEL0) prepare_next(fi);
            // ---- phase 2a: G_1 = pyrDown16S, separable, two channels per op. Away from the ROI border an item is a PAIR of
            // horizontally adjacent outputs: the seven (horizontal pass) / two x five (vertical pass) inputs come in 16-byte
            // shared loads and are split into lanes once.
            if (!border && (GWS & 1) == 0) {
                const SAddr a_g0 = s_addr(s_g0), a_h = s_addr(s_h), a_g1 = s_addr(s_g1); (void)a_g1;
                constexpr int HP = GWS / 2;   // pairs per row
                for (int i = tid; i < ph * HP; i += NT) {
                    const int yy = i / HP, pj = i - yy * HP;
                    // px0 == 2 gx0 - 2: the pair's window starts at region column 4 pj (16-byte aligned, 8 columns inside the pitch)
                    uint32_t q[8];
                    lds_u4(a_g0 + 4 * (yy * PWS + 4 * pj), q[0], q[1], q[2], q[3]);
                    lds_u4(a_g0 + 4 * (yy * PWS + 4 * pj) + 16, q[4], q[5], q[6], q[7]);
                    uint32_t br[7], gg[7];
                    DS_UNROLL
                    for (int t = 0; t < 7; t++) { br[t] = byte_perm(q[t], 0, 0x4240); gg[t] = byte_perm(q[t], 0, 0x4341); }
                    const uint32_t h0br = tap5(br[0], br[1], br[2], br[3], br[4]), h0g = tap5(gg[0], gg[1], gg[2], gg[3], gg[4]);
                    const uint32_t h1br = tap5(br[2], br[3], br[4], br[5], br[6]), h1g = tap5(gg[2], gg[3], gg[4], gg[5], gg[6]);
#if DS_CUDA
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a_h + 8 * (yy * GWS + 2 * pj)), "r"(h0br), "r"(h0g), "r"(h1br), "r"(h1g) : "memory");
#else
                    U2 h0, h1; h0.br = h0br; h0.g = h0g; h1.br = h1br; h1.g = h1g;
                    s_h[yy * GWS + 2 * pj] = h0; s_h[yy * GWS + 2 * pj + 1] = h1;
#endif
                }
                DS_SYNC();
                for (int i = tid; i < gh * HP; i += NT) {
                    const int gyy = i / HP, pj = i - gyy * HP;
                    const int rr = 2 * (gy0 + gyy) - 2 - py0;   // == 2 gyy
                    uint32_t a[5][4];   // rows rr .. rr + 4: {br, g} of the pair's two columns
                    DS_UNROLL
                    for (int t = 0; t < 5; t++) lds_u4(a_h + 8 * ((rr + t) * GWS + 2 * pj), a[t][0], a[t][1], a[t][2], a[t][3]);
                    const uint32_t o0br = ((tap5(a[0][0], a[1][0], a[2][0], a[3][0], a[4][0]) + 0x00800080u) >> 8) & 0x00FF00FFu;
                    const uint32_t o0g = ((tap5(a[0][1], a[1][1], a[2][1], a[3][1], a[4][1]) + 0x00000080u) >> 8) & 0x000000FFu;
                    const uint32_t o1br = ((tap5(a[0][2], a[1][2], a[2][2], a[3][2], a[4][2]) + 0x00800080u) >> 8) & 0x00FF00FFu;
                    const uint32_t o1g = ((tap5(a[0][3], a[1][3], a[2][3], a[3][3], a[4][3]) + 0x00000080u) >> 8) & 0x000000FFu;
#if DS_CUDA
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a_g1 + 8 * (gyy * GWS + 2 * pj)), "r"(o0br), "r"(o0g), "r"(o1br), "r"(o1g) : "memory");
#else
                    U2 o0, o1; o0.br = o0br; o0.g = o0g; o1.br = o1br; o1.g = o1g;
                    s_g1[gyy * GWS + 2 * pj] = o0; s_g1[gyy * GWS + 2 * pj + 1] = o1;
#endif
                    const int gx = gx0 + 2 * pj, gy = gy0 + gyy;
                    if (gy >= jy0 && gy < jy1) {
                        if (gx >= jx0 && gx < jx1 && 2 * pj < gw) G1out[(size_t)gy * op1 + gx] = byte_perm(o0br, o0g, 0x5240);
                        if (gx + 1 >= jx0 && gx + 1 < jx1 && 2 * pj + 1 < gw) G1out[(size_t)gy * op1 + gx + 1] = byte_perm(o1br, o1g, 0x5240);
                    }
                }
                DS_SYNC();
            } else {
                for (int i = tid; i < ph * GWS; i += NT) {
                    const int yy = i / GWS, gxx = i - yy * GWS;
                    if (gxx >= gw) continue;
                    const int c0 = 2 * (gx0 + gxx) - 2;
                    uint32_t q[5];
                    if (!border) {
                        const uint32_t* s = s_g0 + yy * PWS + (c0 - px0);
                        DS_UNROLL
                        for (int t = 0; t < 5; t++) q[t] = s[t];
                    } else {
                        DS_UNROLL
                        for (int t = 0; t < 5; t++) q[t] = s_g0[yy * PWS + (refl101(c0 + t, rw) - px0)];
                    }
                    U2 h;
                    h.br = tap5(byte_perm(q[0], 0, 0x4240), byte_perm(q[1], 0, 0x4240), byte_perm(q[2], 0, 0x4240),
                                byte_perm(q[3], 0, 0x4240), byte_perm(q[4], 0, 0x4240));
                    h.g = tap5(byte_perm(q[0], 0, 0x4341), byte_perm(q[1], 0, 0x4341), byte_perm(q[2], 0, 0x4341),
                               byte_perm(q[3], 0, 0x4341), byte_perm(q[4], 0, 0x4341));
                    s_h[i] = h;
                }
                DS_SYNC();
                for (int i = tid; i < gh * GWS; i += NT) {
                    const int gyy = i / GWS, gxx = i - gyy * GWS;
                    if (gxx >= gw) continue;
                    const int r0 = 2 * (gy0 + gyy) - 2;
                    U2 q[5];
                    if (!border) {
                        DS_UNROLL
                        for (int t = 0; t < 5; t++) q[t] = s_h[(r0 - py0 + t) * GWS + gxx];
                    } else {
                        DS_UNROLL
                        for (int t = 0; t < 5; t++) q[t] = s_h[(refl101(r0 + t, rh) - py0) * GWS + gxx];
                    }
                    U2 o;
                    o.br = ((tap5(q[0].br, q[1].br, q[2].br, q[3].br, q[4].br) + 0x00800080u) >> 8) & 0x00FF00FFu;
                    o.g = ((tap5(q[0].g, q[1].g, q[2].g, q[3].g, q[4].g) + 0x00000080u) >> 8) & 0x000000FFu;
                    s_g1[gyy * GWS + gxx] = o;
                    const int gx = gx0 + gxx, gy = gy0 + gyy;
                    if (gx >= jx0 && gx < jx1 && gy >= jy0 && gy < jy1)
                        G1out[(size_t)gy * op1 + gx] = byte_perm(o.br, o.g, 0x5240);
                }
                DS_SYNC();
            }
            if (tid == 0) *s_vote = 0;   // read by everyone two barriers ago; next written after this frame's last barrier

            // ---- W_1 = pyrDownF32(W_0) over the own range
            if (uni255 || uni0) {
                const float wv = uni255 ? 1.f : 0.f;
                if (!((jx0 | jw) & 3)) {
                    // rows of the level-1 plane are 16-byte aligned and so is the tile's part: four weights per store
                    float4 w4; w4.x = w4.y = w4.z = w4.w = wv;
                    for (int i = NT - 1 - tid; i < jh * (JW / 4); i += NT) {   // the last warps: the first ones had the tail of the pyrDown pass
                        const int jyy = i / (JW / 4), j4 = (i - jyy * (JW / 4)) * 4;
                        if (j4 < jw) *(float4*)(W1out + (size_t)(jy0 + jyy) * op1 + (jx0 + j4)) = w4;
                    }
                } else
                for (int i = tid; i < jh * JW; i += NT) {
                    const int jyy = i / JW, jj = i - jyy * JW;
                    if (jj < jw) W1out[(size_t)(jy0 + jyy) * op1 + (jx0 + jj)] = wv;
                }
            } else {
                const int hr0 = imax(2 * jy0 - 2, 0), hr1 = imin(2 * jy1, rh - 1);
                const int hn = hr1 - hr0 + 1;
                for (int i = tid; i < hn * JW; i += NT) {
                    const int rr_ = i / JW, jj = i - rr_ * JW;
                    if (jj >= jw) continue;
                    const int row = hr0 + rr_, j = jx0 + jj;
                    float t[5];
                    for (int q = 0; q < 5; q++) {
                        const int si = (row - py0) * PWS + (refl101(2 * j + q - 2, rw) - px0);
                        t[q] = LEVEL0 ? f_mul((float)(s_g0[si] >> 24), 1.f / 255.f) : s_w[si];
                    }
                    s_hw[i] = pd_h_is_simd(j, rw, n1x) ? pd_h_simd(t[0], t[1], t[2], t[3], t[4])
                                                        : pd_scalar(t[0], t[1], t[2], t[3], t[4]);
                }
                DS_SYNC();
                for (int i = tid; i < jh * JW; i += NT) {
                    const int jyy = i / JW, jj = i - jyy * JW;
                    if (jj >= jw) continue;
                    const int jy = jy0 + jyy, j = jx0 + jj;
                    float t[5];
                    for (int q = 0; q < 5; q++) t[q] = s_hw[(refl101(2 * jy + q - 2, rh) - hr0) * JW + jj];
                    const float v = pd_v_is_simd(j, n1x) ? pd_v_simd(t[0], t[1], t[2], t[3], t[4])
                                                         : pd_scalar(t[0], t[1], t[2], t[3], t[4]);
                    W1out[(size_t)jy * op1 + j] = f_mul(v, 1.f / 256.f);
                }
                if constexpr (HAS_BOX) DS_SYNC();   // the float H pass shares the buffer the next frame's box is loaded into
            }
            // the next frame's source box comes in while this frame's Laplacians are accumulated (the H buffer is free now)
            issue_box(fi + 1);

            // ---- phase 3: lap = G_0 - pyrUp(G_1), weighted accumulate, one 2x2 quad per item
            // The common tile-frame - all weights 1, no ROI border, the whole tile inside the frame's part and the rows that
            // accumulate - has every shared-memory address of a quad fixed per thread: the coarse 3 x 3 block starts at
            // (qy, qx) of s_g1, the fine pixels at (2 qy + 4, 2 qx + 4) of the warped region. One base register each,
            // immediate offsets, no window tests, no per-pixel predicates.
            const bool full3 = LEVEL0 && uni255 && !border && ax0 == X0 && ax1 == X0 + T && ay0 == Y0 && ay1 == Y0 + T &&
                               Y0 >= p.acc_y0 && Y0 + T <= p.acc_y1;
            if (full3) {
                const SAddr a_g1 = s_addr(s_g1), a_g0 = s_addr(s_g0), a_acc = s_addr(s_acc), a_ws = s_addr(s_ws);
                for (int q = tid; q < NQ; q += NT) {
                    const int qy = q / (T / 2), qx = q - qy * (T / 2);
                    const SAddr ac = a_g1 + (qy * GWS + qx) * 8;
                    U2 a0, a1, a2, b0, b1, b2, d0, d1, d2;
                    lds_u2_o<0>(ac, a0.br, a0.g); lds_u2_o<8>(ac, a1.br, a1.g); lds_u2_o<16>(ac, a2.br, a2.g);
                    lds_u2_o<GWS * 8>(ac, b0.br, b0.g); lds_u2_o<GWS * 8 + 8>(ac, b1.br, b1.g); lds_u2_o<GWS * 8 + 16>(ac, b2.br, b2.g);
                    lds_u2_o<GWS * 16>(ac, d0.br, d0.g); lds_u2_o<GWS * 16 + 8>(ac, d1.br, d1.g); lds_u2_o<GWS * 16 + 16>(ac, d2.br, d2.g);
                    const SAddr ag = a_g0 + ((2 * qy + 4) * PWS + 2 * qx + 4) * 4;
                    const SAddr aa = a_acc + (2 * qy * T + 2 * qx) * 8, aw = a_ws + (2 * qy * T + 2 * qx) * 4;
                    uint32_t g00, g01, g10, g11;
                    lds_u2_o<0>(ag, g00, g01); lds_u2_o<PWS * 4>(ag, g10, g11);
                    const uint32_t El_br = a0.br + 6u * a1.br + a2.br, Ol_br = 4u * (a1.br + a2.br);
                    const uint32_t Ec_br = b0.br + 6u * b1.br + b2.br, Oc_br = 4u * (b1.br + b2.br);
                    const uint32_t Er_br = d0.br + 6u * d1.br + d2.br, Or_br = 4u * (d1.br + d2.br);
                    const uint32_t El_g = a0.g + 6u * a1.g + a2.g, Ol_g = 4u * (a1.g + a2.g);
                    const uint32_t Ec_g = b0.g + 6u * b1.g + b2.g, Oc_g = 4u * (b1.g + b2.g);
                    const uint32_t Er_g = d0.g + 6u * d1.g + d2.g, Or_g = 4u * (d1.g + d2.g);
                    const uint32_t u00br = ((El_br + 6u * Ec_br + Er_br + 0x00200020u) >> 6) & 0x03FF03FFu, u01br = ((Ol_br + 6u * Oc_br + Or_br + 0x00200020u) >> 6) & 0x03FF03FFu;
                    const uint32_t u10br = ((4u * (Ec_br + Er_br) + 0x00200020u) >> 6) & 0x03FF03FFu, u11br = ((4u * (Oc_br + Or_br) + 0x00200020u) >> 6) & 0x03FF03FFu;
                    const uint32_t u00g = ((El_g + 6u * Ec_g + Er_g + 0x20u) >> 6) & 0x3FFu, u01g = ((Ol_g + 6u * Oc_g + Or_g + 0x20u) >> 6) & 0x3FFu;
                    const uint32_t u10g = ((4u * (Ec_g + Er_g) + 0x20u) >> 6) & 0x3FFu, u11g = ((4u * (Oc_g + Or_g) + 0x20u) >> 6) & 0x3FFu;
                    uint32_t v0, v1, v2, v3; float w0, w1;
                    // row 0 of the quad: lap_b + 65536 * lap_r == gbr - up_br as plain integers; trunc(lap * 1) == lap
                    lds_u4_o<0>(aa, v0, v1, v2, v3); lds_f2_o<0>(aw, w0, w1);
                    v0 += byte_perm(g00, 0, 0x4240) - u00br; v1 += ((g00 >> 8) & 255u) - u00g;
                    v2 += byte_perm(g01, 0, 0x4240) - u01br; v3 += ((g01 >> 8) & 255u) - u01g;
                    sts_u4_o<0>(aa, v0, v1, v2, v3); sts_f2_o<0>(aw, f_add(w0, 1.f), f_add(w1, 1.f));
                    lds_u4_o<T * 8>(aa, v0, v1, v2, v3); lds_f2_o<T * 4>(aw, w0, w1);
                    v0 += byte_perm(g10, 0, 0x4240) - u10br; v1 += ((g10 >> 8) & 255u) - u10g;
                    v2 += byte_perm(g11, 0, 0x4240) - u11br; v3 += ((g11 >> 8) & 255u) - u11g;
                    sts_u4_o<T * 8>(aa, v0, v1, v2, v3); sts_f2_o<T * 4>(aw, f_add(w0, 1.f), f_add(w1, 1.f));
                }
            } else
            if (!uni0) {
                for (int q = tid; q < NQ; q += NT) {
                    const int qy = q / (T / 2), qx = q - qy * (T / 2);
                    const int X = X0 + 2 * qx, Y = Y0 + 2 * qy;
                    if (X < ax0 || X >= ax1 || Y < ay0 || Y >= ay1) continue;
                    if (Y + 1 < p.acc_y0 || Y >= p.acc_y1) continue;   // halo rows of a band only produce G_{l+1} / W_{l+1}
                    const int ox = X - rx, oy = Y - ry;
                    const int c1x = ox >> 1, c1y = oy >> 1;
                    int ixl, ixr, iyl, iyr;
                    if (!border) { ixl = c1x - 1; ixr = c1x + 1; iyl = c1y - 1; iyr = c1y + 1; }
                    else { ixl = up_l(c1x, n1x); ixr = up_r(c1x, n1x); iyl = up_l(c1y, n1y); iyr = up_r(c1y, n1y); }
                    const int ixc = c1x - gx0, iyc = c1y - gy0;
                    ixl -= gx0; ixr -= gx0; iyl -= gy0; iyr -= gy0;
                    const U2 a0 = s_g1[iyl * GWS + ixl], a1 = s_g1[iyl * GWS + ixc], a2 = s_g1[iyl * GWS + ixr];
                    const U2 b0 = s_g1[iyc * GWS + ixl], b1 = s_g1[iyc * GWS + ixc], b2 = s_g1[iyc * GWS + ixr];
                    const U2 d0 = s_g1[iyr * GWS + ixl], d1 = s_g1[iyr * GWS + ixc], d2 = s_g1[iyr * GWS + ixr];
                    // horizontal: even = l + 6c + r, odd = 4(c + r), on rows l / c / r (packed lanes)
                    const uint32_t El_br = a0.br + 6u * a1.br + a2.br, Ol_br = 4u * (a1.br + a2.br);
                    const uint32_t Ec_br = b0.br + 6u * b1.br + b2.br, Oc_br = 4u * (b1.br + b2.br);
                    const uint32_t Er_br = d0.br + 6u * d1.br + d2.br, Or_br = 4u * (d1.br + d2.br);
                    const uint32_t El_g = a0.g + 6u * a1.g + a2.g, Ol_g = 4u * (a1.g + a2.g);
                    const uint32_t Ec_g = b0.g + 6u * b1.g + b2.g, Oc_g = 4u * (b1.g + b2.g);
                    const uint32_t Er_g = d0.g + 6u * d1.g + d2.g, Or_g = 4u * (d1.g + d2.g);
                    // vertical + (v + 32) >> 6; index [dy][dx]
                    uint32_t up_br[2][2], up_g[2][2];
                    up_br[0][0] = ((El_br + 6u * Ec_br + Er_br + 0x00200020u) >> 6) & 0x03FF03FFu;
                    up_br[0][1] = ((Ol_br + 6u * Oc_br + Or_br + 0x00200020u) >> 6) & 0x03FF03FFu;
                    up_br[1][0] = ((4u * (Ec_br + Er_br) + 0x00200020u) >> 6) & 0x03FF03FFu;
                    up_br[1][1] = ((4u * (Oc_br + Or_br) + 0x00200020u) >> 6) & 0x03FF03FFu;
                    up_g[0][0] = ((El_g + 6u * Ec_g + Er_g + 0x20u) >> 6) & 0x3FFu;
                    up_g[0][1] = ((Ol_g + 6u * Oc_g + Or_g + 0x20u) >> 6) & 0x3FFu;
                    up_g[1][0] = ((4u * (Ec_g + Er_g) + 0x20u) >> 6) & 0x3FFu;
                    up_g[1][1] = ((4u * (Oc_g + Or_g) + 0x20u) >> 6) & 0x3FFu;
                    DS_UNROLL
                    for (int dy = 0; dy < 2; dy++) {
                        // the two pixels of a quad row are adjacent in s_g0 / s_acc / s_ws: one wide access each
                        const int gsi = (oy + dy - py0) * PWS + (ox - px0);       // even
                        const int ti = (2 * qy + dy) * T + 2 * qx;                // even
                        const uint2 g2 = *(const uint2*)(s_g0 + gsi);
                        int4 av = *(const int4*)(s_acc + ti);
                        float2 wv2 = *(const float2*)(s_ws + ti);
                        const uint32_t g0v[2] = {g2.x, g2.y};
                        int accbr[2] = {av.x, av.z}, accg[2] = {av.y, av.w};
                        float wsv[2] = {wv2.x, wv2.y};
                        DS_UNROLL
                        for (int dx = 0; dx < 2; dx++) {
                            const uint32_t g0 = g0v[dx];
                            const uint32_t gbr = byte_perm(g0, 0, 0x4240), gg = (g0 >> 8) & 255u;
                            if (uni255) {
                                // lap_b + 65536 * lap_r == gbr - up_br as plain integers
                                accbr[dx] += (int)(gbr - up_br[dy][dx]);
                                accg[dx] += (int)gg - (int)up_g[dy][dx];
                                wsv[dx] = f_add(wsv[dx], 1.f);
                            } else {
                                const float wv = LEVEL0 ? f_mul((float)(g0 >> 24), 1.f / 255.f) : s_w[gsi + dx];
                                const int lb = (int)(gbr & 0xFFFFu) - (int)(up_br[dy][dx] & 0xFFFFu);
                                const int lr = (int)(gbr >> 16) - (int)(up_br[dy][dx] >> 16);
                                const int lg = (int)gg - (int)up_g[dy][dx];
                                const int tb = (int)(short)f2i_rz(f_mul((float)lb, wv));
                                const int tr = (int)(short)f2i_rz(f_mul((float)lr, wv));
                                const int tg = (int)(short)f2i_rz(f_mul((float)lg, wv));
                                accbr[dx] += tb + tr * 65536;
                                accg[dx] += tg;
                                wsv[dx] = f_add(wsv[dx], wv);
                            }
                        }
                        av.x = accbr[0]; av.y = accg[0]; av.z = accbr[1]; av.w = accg[1];
                        *(int4*)(s_acc + ti) = av;
                        wv2.x = wsv[0]; wv2.y = wsv[1];
                        *(float2*)(s_ws + ti) = wv2;
                    }
                }
            }
            }   // ======== end of the second half
#undef DS_TF_LOCALS
            DS_SYNC();
        }

        // ---- normalise and store level 0 of the canvas pyramid
        for (int i = tid; i < T * T; i += NT) {
            const int yy = i / T, xx = i - yy * T;
            const int X = X0 + xx, Y = Y0 + yy;
            if (X >= p.dst_w || Y >= p.dst_h || Y < p.acc_y0 || Y >= p.acc_y1) continue;
            const int2 av = s_acc[i];
            const int abr = av.x;
            const int sb = (int)(short)(abr & 0xFFFF);
            const int sr = (abr - sb) >> 16;
            uint2 o;
            norm_px(sb, av.y, sr, s_ws[i], o.x, o.y);
            *(uint2*)(p.dst + (size_t)Y * p.dst_w + X) = o;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// MULTIBAND, levels >= 1, streaming formulation. The per-frame pyramids are plain planes in HBM, so above level 0
// nothing has to be staged tile by tile:
//   ds_mb_pyrdown  G_{l+1} = pyrDown16S(G_l), W_{l+1} = pyrDownF32(W_l) of every frame, each output computed once
//                  (no tile halo), 2 x 2 outputs per thread from a 7 x 7 register window, no shared memory;
//   ds_mb_accum    a canvas tile of level l walks its frames in feed order; every thread owns one 2 x 2 quad and keeps
//                  its Laplacian and weight sums in registers: lap = G_l - pyrUp(G_{l+1}) straight from the planes
//                  (neighbouring threads share the taps through L1), acc += trunc(lap * W_l); no barrier inside the
//                  frame loop; normalised level written once.
// Same arithmetic as MBBody (A8 - A11); the tile-based kernels spent 4 barriers and a 39 x 39 halo region per 32 x 32
// tile-frame on this (level 1 of cfg2: 287 M warp instructions, 0.38 ms).

// "Every weight W_l over the pixels [x0, x1] x [y0, y1] (level-l coordinates relative to the frame's feed ROI) is exactly
// 1", from geometry alone: a plane-mapped frame without per-pixel mask whose level-0 support of the region (radius
// 2^(l+1) - 2 after l pyrDowns) lies inside the bbox and maps into the source at its four corners - x(u, v) is monotone
// in u and in v even in float arithmetic (a chain of monotone roundings), so the corners bound every pixel; the nearest
// mask is then 255 throughout, W_0 = 255 * (1 / 255f) = 1 and every pyrDown of ones is exactly 1. The emulator checks
// every claim against the data.
DS_D bool weights_all_ones(const FrameDev& F, int l, int x0, int x1, int y0, int y1) {
    if (!(F.kind == XF_PLANE && !F.seam && F.k[6] == 0.f && F.k[7] == 0.f && F.k8one == 1.f && f_mul(255.f, 1.f / 255.f) == 1.f)) return false;
    const int rl = (2 << l) - 2;
    const int u_lo = F.rx + (x0 << l) - rl - F.cx, u_hi = F.rx + (x1 << l) + rl - F.cx;
    const int v_lo = F.ry + (y0 << l) - rl - F.cy, v_hi = F.ry + (y1 << l) + rl - F.cy;
    if (!(u_lo >= 0 && u_hi <= F.w - 1 && v_lo >= 0 && v_hi <= F.h - 1)) return false;
    float ca0[2], ca3[2], rb1[2], rb4[2];
    DS_UNROLL
    for (int e = 0; e < 2; e++) {
        float U = (float)(F.tlx + (e ? u_hi : u_lo)), V = (float)(F.tly + (e ? v_hi : v_lo));
        if (F.scale != 1.f) { U = f_div(U, F.scale); V = f_div(V, F.scale); }
        const float up = f_sub(U, F.t0), vp = f_sub(V, F.t1);
        ca0[e] = f_mul(F.k[0], up); ca3[e] = f_mul(F.k[3], up);
        rb1[e] = f_mul(F.k[1], vp); rb4[e] = f_mul(F.k[4], vp);
    }
    float xmn = 3.0e38f, xmx = -3.0e38f, ymn = 3.0e38f, ymx = -3.0e38f;
    DS_UNROLL
    for (int e = 0; e < 4; e++) {
        const float xx_ = f_add(f_add(ca0[e & 1], rb1[e >> 1]), F.k2one), yy_ = f_add(f_add(ca3[e & 1], rb4[e >> 1]), F.k5one);
        xmn = fminf(xmn, xx_); xmx = fmaxf(xmx, xx_); ymn = fminf(ymn, yy_); ymx = fmaxf(ymx, yy_);
    }
    return xmn >= 0.f && xmx <= (float)(F.src_w - 1) && ymn >= 0.f && ymx <= (float)(F.src_h - 1);
}

struct PyrParams {
    const FrameDev* frames; int nframes;
    int level;            // input level l (>= 1); output level l + 1
    int txmax, R;         // CTAs per frame: txmax column blocks x R row blocks (first row block = the frame's first needed row)
    uint32_t m_per, m_tx; // floor(2^32 / (txmax * R)), floor(2^32 / txmax): block index decode without divisions
    int own_y0, own_y1;   // canvas rows of level l + 1 to produce
    const void* lmaps; int lstride;   // tensor maps of the per-frame planes (MBParams::lmaps), or NULL
    int boxes;            // input windows staged in shared memory (TMA; the emulator copies) where no border index is involved
};
struct PyrDownBody {
    // One CTA = 32 x 64 outputs; one thread = a strip of 2 columns x 8 rows, marched down two output rows at a time: of
    // the seven input rows a 2 x 2 output block reads, three were filtered horizontally for the block above.
    static constexpr int BW = 32, STRIP = 4, BH = 8 * STRIP, NTH = 128;
    // CTAs whose input window lies inside the plane (no BORDER_REFLECT_101 index to resolve) bring it into shared memory
    // with one TMA box per plane - (2 BW + 8) x (2 BH + 3) elements from a 16-byte aligned column - and filter from there.
    static constexpr int XW = 2 * BW + 8, XH = 2 * BH + 3, XBOX = XW * XH * 4, XBOX_AL = (XBOX + 127) & ~127;
    static constexpr int LM_PG = 3, LM_PW = 4, LM_N = 5;   // == AccumBody::LM_*
    static int smem_bytes() { return 2 * XBOX_AL + 16; }
    struct HRow { uint32_t br[2], g[2]; float w[2]; };   // horizontally filtered input row at the thread's two output columns
    // q = i / d with m = floor(2^32 / d): the estimate is at most 2 short
    DS_DM int div_m(int i, int d, uint32_t m) {
        int q = (int)(((unsigned long long)(uint32_t)i * m) >> 32);
        int r = i - q * d;
        if (r >= d) { q++; r -= d; }
        if (r >= d) q++;
        return q;
    }
    template <int NT>
    DS_DM void run(const PyrParams& p, int block, int tid, unsigned char* smem) {
        const int per = p.txmax * p.R;
        const int fslot = div_m(block, per, p.m_per), rem = block - fslot * per;
        const int byr = div_m(rem, p.txmax, p.m_tx), bx = rem - byr * p.txmax;
        const FrameDev& F = p.frames[fslot];
        const int l = p.level;
        const int w_in = F.rw >> l, h_in = F.rh >> l, w_out = w_in >> 1, h_out = h_in >> 1;
        const int ry1 = F.ry >> (l + 1);
        const int jlo = imax(p.own_y0 - ry1, 0), jhi = imin(p.own_y1 - ry1, h_out);
        if (jlo >= jhi) return;   // unused slot (rw == rh == 0) or a frame outside the rows
        const int by = jlo / BH + byr;
        if (by * BH >= jhi || bx * BW >= w_out) return;
        const uint32_t* const Gin = (const uint32_t*)F.G[l];
        const float* const Win = F.W[l];
        uint32_t* const Gout = (uint32_t*)F.G[l + 1];
        float* const Wout = F.W[l + 1];
        const int ip = F.gp[l], op = F.gp[l + 1];
        // the CTA's outputs all have weight 1 by geometry: W_l is not even read
        const bool ones = weights_all_ones(F, l + 1, bx * BW, imin(bx * BW + BW, w_out) - 1, imax(by * BH, jlo), imin(by * BH + BH, jhi) - 1);
        const bool need_w = !ones || !DS_CUDA;   // (the emulator computes them anyway and checks the claim)
        // input window of the CTA: columns 2 bx BW - 2 .. + 2 BW + 2, rows 2 by BH - 2 .. + 2 BH + 2
        const int xs = 2 * bx * BW - 4, ys = 2 * by * BH - 2;   // box origin (column rounded down to 4 elements)
        const bool boxed = p.boxes && xs >= 0 && ys >= 0 && 2 * bx * BW + 2 * BW + 2 <= w_in - 1 && 2 * by * BH + 2 * BH + 2 <= h_in - 1;
        unsigned long long* const s_bar = (unsigned long long*)(smem + 2 * XBOX_AL);
        if (boxed) {
            if (tid == 0) {
                const char* lm = (const char*)((uintptr_t)p.lmaps + ((size_t)fslot * p.lstride + l) * (LM_N * 128));
                mbar_init(s_bar, 1);
                mbar_expect_tx(s_bar, (uint32_t)(need_w && DS_CUDA ? 2 * XBOX : XBOX));
                box_load(smem, lm + LM_PG * 128, Gin, w_in, h_in, ip, xs, ys, XW, XH, s_bar);
                if (need_w) box_load(smem + XBOX_AL, lm + LM_PW * 128, Win, w_in, h_in, ip, xs, ys, XW, XH, s_bar);
            }
            DS_SYNC();          // the barrier word is initialised for everyone
            mbar_wait(s_bar, 0u);
        }
        const SAddr a_gbox = s_addr(smem), a_wbox = s_addr(smem + XBOX_AL);
        for (int t = tid; t < NTH; t += NT) {
            const int tx = t & 15, ty = t >> 4;
            const int jx = bx * BW + 2 * tx, jy0 = by * BH + ty * STRIP;   // the strip's first output
            if (jx >= w_out || jy0 >= jhi || jy0 + STRIP <= jlo) continue;
            const int c0 = 2 * jx - 2;                                     // seven input columns from c0
            const bool xin = c0 >= 0 && c0 + 6 <= w_in - 1;
            int cols[7];
            DS_UNROLL
            for (int k = 0; k < 7; k++) cols[k] = xin ? c0 + k : refl101(c0 + k, w_in);
            const bool hs0 = pd_h_is_simd(jx, w_in, w_out), hs1 = pd_h_is_simd(jx + 1, w_in, w_out);
            const bool vs0 = pd_v_is_simd(jx, w_out), vs1 = pd_v_is_simd(jx + 1, w_out);
            // input row r (reflected into the plane, BORDER_REFLECT_101) filtered horizontally: [1 4 6 4 1] on packed lanes
            // (B | R << 16, G); the weights in the op order of each column's class (A9)
            auto hrow = [&](int r, HRow& h) {
                const int rr = refl101(r, h_in);
                const uint32_t* gr = Gin + (size_t)rr * ip;
                const float* wr = Win + (size_t)rr * ip;
                uint32_t q[7]; float w[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (boxed) {
                    // (inside the plane: rr == r) 8 + 16 + 4 bytes at 16-byte aligned shared addresses
                    const int o = (r - ys) * XW + (c0 - xs);
                    lds_u2(a_gbox + 4 * o, q[0], q[1]); lds_u4(a_gbox + 4 * o + 8, q[2], q[3], q[4], q[5]); q[6] = lds_u1(a_gbox + 4 * o + 24);
                    if (need_w) {
                        uint32_t u2, u3, u4, u5;
                        lds_f2(a_wbox + 4 * o, w[0], w[1]);
                        lds_u4(a_wbox + 4 * o + 8, u2, u3, u4, u5);
                        w[2] = i2f_bits((int)u2); w[3] = i2f_bits((int)u3); w[4] = i2f_bits((int)u4); w[5] = i2f_bits((int)u5);
                        w[6] = i2f_bits((int)lds_u1(a_wbox + 4 * o + 24));
                    }
                } else if (xin) {
                    // c0 = 2 mod 4 and rows are 16-byte aligned: 8 + 16 + 4 bytes
                    const uint2 a = ld_ro((const uint2*)(gr + c0)); const uint4 b = ld_ro((const uint4*)(gr + c0 + 2)); const uint32_t c = ld_ro(gr + c0 + 6);
                    q[0] = a.x; q[1] = a.y; q[2] = b.x; q[3] = b.y; q[4] = b.z; q[5] = b.w; q[6] = c;
                    if (need_w) {
                        const float2 fa = ld_ro((const float2*)(wr + c0)); const float4 fb = ld_ro((const float4*)(wr + c0 + 2)); const float fc = ld_ro(wr + c0 + 6);
                        w[0] = fa.x; w[1] = fa.y; w[2] = fb.x; w[3] = fb.y; w[4] = fb.z; w[5] = fb.w; w[6] = fc;
                    }
                } else {
                    DS_UNROLL
                    for (int c = 0; c < 7; c++) { q[c] = ld_ro(gr + cols[c]); if (need_w) w[c] = ld_ro(wr + cols[c]); }
                }
                uint32_t br[7], gg[7];
                DS_UNROLL
                for (int c = 0; c < 7; c++) { br[c] = byte_perm(q[c], 0, 0x4240); gg[c] = byte_perm(q[c], 0, 0x4441); }
                DS_UNROLL
                for (int o = 0; o < 2; o++) {
                    h.br[o] = br[2 * o] + br[2 * o + 4] + 6u * br[2 * o + 2] + 4u * (br[2 * o + 1] + br[2 * o + 3]);
                    h.g[o] = gg[2 * o] + gg[2 * o + 4] + 6u * gg[2 * o + 2] + 4u * (gg[2 * o + 1] + gg[2 * o + 3]);
                    h.w[o] = 0.f;
                    if (need_w) {
                        const bool hs = o ? hs1 : hs0;
                        h.w[o] = hs ? pd_h_simd(w[2 * o], w[2 * o + 1], w[2 * o + 2], w[2 * o + 3], w[2 * o + 4])
                                    : pd_scalar(w[2 * o], w[2 * o + 1], w[2 * o + 2], w[2 * o + 3], w[2 * o + 4]);
                    }
                }
            };
            // output row y from five filtered rows
            auto emit = [&](int y, const HRow& a, const HRow& b, const HRow& c, const HRow& d, const HRow& e) {
                if (y < jlo || y >= jhi) return;
                uint32_t go[2]; float wo[2];
                DS_UNROLL
                for (int o = 0; o < 2; o++) {
                    const uint32_t vbr = a.br[o] + e.br[o] + 6u * c.br[o] + 4u * (b.br[o] + d.br[o]);
                    const uint32_t vg = a.g[o] + e.g[o] + 6u * c.g[o] + 4u * (b.g[o] + d.g[o]);
                    const uint32_t obr = ((vbr + 0x00800080u) >> 8) & 0x00FF00FFu, og = ((vg + 0x80u) >> 8) & 0xFFu;
                    go[o] = byte_perm(obr, og, 0x5240);
                    wo[o] = 1.f;
                    if (need_w) {
                        const bool vs = o ? vs1 : vs0;
                        const float v = vs ? pd_v_simd(a.w[o], b.w[o], c.w[o], d.w[o], e.w[o]) : pd_scalar(a.w[o], b.w[o], c.w[o], d.w[o], e.w[o]);
                        wo[o] = f_mul(v, 1.f / 256.f);
#if !DS_CUDA
                        if (ones && wo[o] != 1.f && jx + o < w_out) { fprintf(stderr, "ds emu: pyrdown level %d frame %d: weights claimed 1 by geometry are not\n", l, fslot); abort(); }
#endif
                    }
                }
                uint32_t* gq = Gout + (size_t)y * op + jx;
                float* wq = Wout + (size_t)y * op + jx;
                if (jx + 1 < w_out) {   // jx is even and the pitch a multiple of 4: 8-byte aligned pairs
                    uint2 gv; gv.x = go[0]; gv.y = go[1]; *(uint2*)gq = gv;
                    float2 wv; wv.x = wo[0]; wv.y = wo[1]; *(float2*)wq = wv;
                } else { gq[0] = go[0]; wq[0] = wo[0]; }
            };
            HRow h0, h1, h2, h3, h4, h5, h6;
            int r0 = 2 * jy0 - 2;
            hrow(r0, h0); hrow(r0 + 1, h1); hrow(r0 + 2, h2);
            DS_UNROLL
            for (int s = 0; s < STRIP / 2; s++) {
                const int y = jy0 + 2 * s;
                if (y >= jhi) break;
                hrow(r0 + 3, h3); hrow(r0 + 4, h4); hrow(r0 + 5, h5); hrow(r0 + 6, h6);
                emit(y, h0, h1, h2, h3, h4);
                emit(y + 1, h2, h3, h4, h5, h6);
                h0 = h4; h1 = h5; h2 = h6; r0 += 4;
            }
        }
    }
};

#ifndef DS_ACC_TW
#define DS_ACC_TW 32
#define DS_ACC_TH 16
#endif
struct AccumBody {
    static constexpr int TW = DS_ACC_TW, TH = DS_ACC_TH, NQ = TW * TH / 4, QW = TW / 2;   // one thread = one 2 x 2 quad
    static constexpr int GCH = 32;                     // frames whose geometry is staged at a time; also the span of the packed sums
    // Shared-memory ring (below the top level, when the tensor maps exist): the boxes of up to NST frames of the tile's list
    // are in flight at once - G_l and W_l of the tile (36 x 16: the innermost start of a TMA box must be 16-byte aligned, so
    // up to 3 extra columns on the left) and G_{l+1} of the tile / 2 + its pyrUp ring (24 x 10) - each stage completing on
    // its own mbarrier. Loads in flight no longer cost registers, and the taps become shared-memory reads at fixed offsets.
    static constexpr int NST = 4, GBW = TW + 4, GBH = TH, CBW = (TW / 2 + 2 + 3 + 3) & ~3, CBH = TH / 2 + 2;
    static constexpr int GBOX = GBW * GBH * 4, CBOX = CBW * CBH * 4;
    static constexpr int STAGE_BYTES = 2 * GBOX + ds_al128(CBOX);
    static constexpr int LM_G = 0, LM_W = 1, LM_C = 2, LM_PG = 3, LM_PW = 4, LM_N = 5;   // tensor maps per frame and level (MBParams::lmaps)
    // tile x frame geometry: built by one thread, read by all. Pointers are pre-offset to the tile origin (they may point
    // outside the plane; the quads inside the frame's ROI add what brings them back in).
    struct alignas(16) AGeo {
        const uint32_t* G; const float* W;   // level l at ROI-relative (X0 - rx, Y0 - ry)
        const uint32_t* G1; int gp, gp1;     // level l + 1 at ((X0 - rx) / 2, (Y0 - ry) / 2); row pitches
        int x0, nx, y0, ny;                  // the tile's part inside the ROI and the rows: tile-relative quads (top level: pixels)
        int c1x0, c1y0, n1x, n1y;            // level-(l+1) coordinate of the tile origin in the frame, plane size (pyrUp border rules)
        int skip, ones, border, frame;
        int gbx, gby, cbx, cby;              // ring: box origins in the level-l / level-(l+1) plane
        int goff, coff, pad0, pad1;          // ring: element offset of the tile origin inside the G / W box and the coarse box
    };
    static constexpr int GEO_BYTES = GCH * (int)sizeof(AGeo), RING_OFF = ds_al128(GEO_BYTES);
    static int smem_bytes() { return RING_OFF + NST * STAGE_BYTES + 64; }
    struct U2 { uint32_t br, g; };
    DS_DM U2 split(uint32_t v) { U2 o; o.br = byte_perm(v, 0, 0x4240); o.g = byte_perm(v, 0, 0x4441); return o; }

    DS_DM void make_geo(const MBParams& p, int frame, int X0, int Y0, bool top, AGeo& g) {
        const FrameDev& F = p.frames[frame];
        const int l = p.level;
        const int rx = F.rx >> l, ry = F.ry >> l, rw = F.rw >> l, rh = F.rh >> l;
        const int ax0 = imax(X0, rx), ax1 = imin(X0 + TW, rx + rw);
        // below the top level the ROI is even-aligned and quads are whole: rows rounded out to quads (the store filters)
        const int a0 = top ? p.acc_y0 : (p.acc_y0 & ~1), a1 = top ? p.acc_y1 : ((p.acc_y1 + 1) & ~1);
        const int ay0 = imax(imax(Y0, ry), a0), ay1 = imin(imin(Y0 + TH, ry + rh), a1);
        g.skip = (ax0 >= ax1 || ay0 >= ay1) ? 1 : 0;
        g.frame = frame;
        g.gp = F.gp[l]; g.gp1 = top ? 0 : F.gp[l + 1];
        const long long o0 = (long long)(Y0 - ry) * g.gp + (X0 - rx);
        g.G = (const uint32_t*)F.G[l] + o0; g.W = F.W[l] + o0;
        g.c1x0 = (X0 - rx) >> 1; g.c1y0 = (Y0 - ry) >> 1; g.n1x = rw >> 1; g.n1y = rh >> 1;
        g.G1 = top ? nullptr : (const uint32_t*)F.G[l + 1] + ((long long)g.c1y0 * g.gp1 + g.c1x0);
        const int sh = top ? 0 : 1;
        g.x0 = (ax0 - X0) >> sh; g.nx = (ax1 - ax0) >> sh; g.y0 = (ay0 - Y0) >> sh; g.ny = (ay1 - ay0) >> sh;
        // does any quad of the part tap beyond the level-(l+1) plane (pyrUp border rules)?
        g.border = (((ax0 - rx) >> 1) < 1 || ((ax1 - 1 - rx) >> 1) + 1 > g.n1x - 1 || ((ay0 - ry) >> 1) < 1 || ((ay1 - 1 - ry) >> 1) + 1 > g.n1y - 1) ? 1 : 0;
        g.ones = 0; g.pad0 = g.pad1 = 0;
        if (!g.skip && weights_all_ones(F, l, ax0 - rx, ax1 - 1 - rx, ay0 - ry, ay1 - 1 - ry)) g.ones = 1;
        // ring boxes: origins clamped into the plane (whatever lies left of / above it belongs to no active quad)
        g.gbx = imax(X0 - rx, 0) & ~3; g.gby = imax(Y0 - ry, 0);
        g.goff = (Y0 - ry - g.gby) * GBW + (X0 - rx - g.gbx);
        g.cbx = imax(g.c1x0 - 1, 0) & ~3; g.cby = imax(g.c1y0 - 1, 0);
        g.coff = (g.c1y0 - g.cby) * CBW + (g.c1x0 - g.cbx);
    }

    // pyrUp of the 3 x 3 coarse neighbourhood (A10, packed lanes) and the weighted Laplacians of the quad
    DS_DM void quad(const uint32_t (&cw)[9], const uint32_t (&g0v)[4], const float (&wv4)[4], bool ones, int (&pbr)[4], int (&pg)[4], float (&ws)[4]) {
        const U2 a0 = split(cw[0]), a1 = split(cw[1]), a2 = split(cw[2]), b0 = split(cw[3]), b1 = split(cw[4]), b2 = split(cw[5]), d0 = split(cw[6]), d1 = split(cw[7]), d2 = split(cw[8]);
        // horizontal: even = l + 6c + r, odd = 4(c + r), on rows l / c / r
        const uint32_t El_br = a0.br + 6u * a1.br + a2.br, Ol_br = 4u * (a1.br + a2.br);
        const uint32_t Ec_br = b0.br + 6u * b1.br + b2.br, Oc_br = 4u * (b1.br + b2.br);
        const uint32_t Er_br = d0.br + 6u * d1.br + d2.br, Or_br = 4u * (d1.br + d2.br);
        const uint32_t El_g = a0.g + 6u * a1.g + a2.g, Ol_g = 4u * (a1.g + a2.g);
        const uint32_t Ec_g = b0.g + 6u * b1.g + b2.g, Oc_g = 4u * (b1.g + b2.g);
        const uint32_t Er_g = d0.g + 6u * d1.g + d2.g, Or_g = 4u * (d1.g + d2.g);
        // vertical + (v + 32) >> 6; index = 2 dy + dx
        uint32_t up_br[4], up_g[4];
        up_br[0] = ((El_br + 6u * Ec_br + Er_br + 0x00200020u) >> 6) & 0x03FF03FFu;
        up_br[1] = ((Ol_br + 6u * Oc_br + Or_br + 0x00200020u) >> 6) & 0x03FF03FFu;
        up_br[2] = ((4u * (Ec_br + Er_br) + 0x00200020u) >> 6) & 0x03FF03FFu;
        up_br[3] = ((4u * (Oc_br + Or_br) + 0x00200020u) >> 6) & 0x03FF03FFu;
        up_g[0] = ((El_g + 6u * Ec_g + Er_g + 0x20u) >> 6) & 0x3FFu;
        up_g[1] = ((Ol_g + 6u * Oc_g + Or_g + 0x20u) >> 6) & 0x3FFu;
        up_g[2] = ((4u * (Ec_g + Er_g) + 0x20u) >> 6) & 0x3FFu;
        up_g[3] = ((4u * (Oc_g + Or_g) + 0x20u) >> 6) & 0x3FFu;
        DS_UNROLL
        for (int i = 0; i < 4; i++) {
            const U2 gv = split(g0v[i]);
            if (ones) {
                // trunc(lap * 1) == lap, and lap_b + 65536 * lap_r == gbr - up_br as plain integers
                pbr[i] += (int)(gv.br - up_br[i]);
                pg[i] += (int)gv.g - (int)up_g[i];
                ws[i] = f_add(ws[i], 1.f);
            } else {
                const float wv = wv4[i];
                const int lb = (int)(gv.br & 0xFFFFu) - (int)(up_br[i] & 0xFFFFu);
                const int lr = (int)(gv.br >> 16) - (int)(up_br[i] >> 16);
                const int lg = (int)gv.g - (int)up_g[i];
                const int tb = (int)(short)f2i_rz(f_mul((float)lb, wv)), tr = (int)(short)f2i_rz(f_mul((float)lr, wv));
                pbr[i] += tb + tr * 65536;
                pg[i] += (int)(short)f2i_rz(f_mul((float)lg, wv));
                ws[i] = f_add(ws[i], wv);
            }
        }
    }

    template <int NT>
    DS_DM void run(const MBParams& p, int block, int tid, unsigned char* smem) {
        constexpr int K = NQ / NT;   // quads per thread: 1 on the GPU; the emulator's single thread holds them all
        AGeo* s_geo = (AGeo*)smem;
        unsigned char* const s_ring = smem + RING_OFF;
        unsigned long long* const s_bar = (unsigned long long*)(s_ring + NST * STAGE_BYTES);
        const int4 rec = ld_ro(p.tile_rec + block);
        const int tile = rec.x, f_begin = rec.y, f_end = rec.z;
        const int ty = PyrDownBody::div_m(tile, p.tiles_x, p.m_tiles_x), tx = tile - ty * p.tiles_x;
        const int X0 = tx * TW, Y0 = ty * TH;
        const bool top = p.level == p.L;
        const bool ring = !top && (p.flags & 4) && (p.lmaps != nullptr || !DS_CUDA);
        if (ring && tid == 0) { for (int s = 0; s < NST; s++) mbar_init(s_bar + s, 1); }   // visible after the first barrier below
        // per pixel: packed sums B + 65536 R and G of the current chunk of frames (exact while |sum B| < 2^15: a chunk is 32
        // frames of |lap| <= 255), folded into wide sums after every chunk; weight sum in feed order
        int pbr[K][4], pg[K][4], sB[K][4], sR[K][4]; float ws[K][4];
        DS_UNROLL
        for (int k = 0; k < K; k++) { DS_UNROLL for (int i = 0; i < 4; i++) { pbr[k][i] = pg[k][i] = sB[k][i] = sR[k][i] = 0; ws[k][i] = 0.f; } }
        int m_issue = 0, m_use = 0;   // live (non-skipped) tile-frames issued into / consumed from the ring so far
        for (int base = f_begin; base < f_end; base += GCH) {
            const int n = imin(GCH, f_end - base);
            if (base > f_begin) DS_SYNC();   // everyone is done with the previous chunk's entries
            for (int j = tid; j < n; j += NT) make_geo(p, (base + j == f_begin) ? rec.w : p.tile_frames[base + j], X0, Y0, top, s_geo[j]);
            DS_SYNC();
            if (ring) {
                // ---- ring path. Thread 0 alone walks the list and issues the boxes; the others only need the number of live
                // entries of the chunk to know - block-uniformly - when a consumed stage has to be handed back for a refill.
                const SAddr a_geo = s_addr(s_geo);
                int nl = 0;
                for (int j = 0; j < n; j++) nl += (int)lds_u1(a_geo + j * (int)sizeof(AGeo) + 64) ? 0 : 1;
                int jn = 0;   // thread 0: next entry to issue
                auto issue_one = [&](int stage) {   // thread 0 only
                    while (s_geo[jn].skip) jn++;
                    const AGeo& g = s_geo[jn++];
                    unsigned char* st = s_ring + stage * STAGE_BYTES;
                    const char* lm = (const char*)((uintptr_t)p.lmaps + ((size_t)g.frame * p.lstride + p.level) * (LM_N * 128));
                    const char* lc = (const char*)((uintptr_t)p.lmaps + ((size_t)g.frame * p.lstride + p.level + 1) * (LM_N * 128));
#if DS_CUDA
                    const void* pG = nullptr; const void* pW = nullptr; const void* pC = nullptr;   // the tensor maps know the planes
                    const int pw_ = 0, ph_ = 0;
#else
                    const FrameDev& F = p.frames[g.frame];
                    const void* pG = F.G[p.level]; const void* pW = F.W[p.level]; const void* pC = F.G[p.level + 1];
                    const int pw_ = F.rw >> p.level, ph_ = F.rh >> p.level;
#endif
                    mbar_expect_tx(s_bar + stage, (uint32_t)(GBOX + CBOX + (g.ones ? 0 : GBOX)));
                    box_load(st, lm + LM_G * 128, pG, pw_, ph_, g.gp, g.gbx, g.gby, GBW, GBH, s_bar + stage);
                    if (!g.ones || !DS_CUDA) box_load(st + GBOX, lm + LM_W * 128, pW, pw_, ph_, g.gp, g.gbx, g.gby, GBW, GBH, s_bar + stage);
                    box_load(st + 2 * GBOX, lc + LM_C * 128, pC, g.n1x, g.n1y, g.gp1, g.cbx, g.cby, CBW, CBH, s_bar + stage);
                };
                const int first = imin(nl, NST);
                if (tid == 0) {
                    if (m_issue > 0) fence_proxy_async_smem();   // the stages were read through the generic proxy in the previous chunk
                    for (int i = 0; i < first; i++) issue_one((m_issue + i) % NST);
                }
                m_issue += first;
                int lm_ = 0;   // live entries of this chunk consumed so far
                for (int j = 0; j < n; j++) {
                    uint32_t gx0_, gnx, gy0_, gny, gskip, gones, gborder, gframe, goff, coff;
                    lds_u4(a_geo + j * (int)sizeof(AGeo) + 64, gskip, gones, gborder, gframe);
                    if (gskip) continue;   // block-uniform
                    lds_u4(a_geo + j * (int)sizeof(AGeo) + 32, gx0_, gnx, gy0_, gny);
                    lds_u2(a_geo + j * (int)sizeof(AGeo) + 96, goff, coff);
                    const int s = m_use % NST;
                    const SAddr a_st = s_addr(s_ring + s * STAGE_BYTES);
                    mbar_wait(s_bar + s, (uint32_t)((m_use / NST) & 1));
                    DS_UNROLL
                    for (int k = 0; k < K; k++) {
                        const int q = tid + k * NT;
                        const int qx = q % QW, qy = q / QW;
                        if ((unsigned)(qx - (int)gx0_) >= gnx || (unsigned)(qy - (int)gy0_) >= gny) continue;
                        uint32_t cw[9], g0v[4]; float wv4[4] = {1.f, 1.f, 1.f, 1.f};
                        const SAddr ag = a_st + 4 * ((int)goff + 2 * qy * GBW + 2 * qx);
                        const SAddr ac = a_st + 2 * GBOX + 4 * ((int)coff + qy * CBW + qx);
                        if (!gborder) {
                            cw[0] = lds_u1(ac - 4 * CBW - 4); cw[1] = lds_u1(ac - 4 * CBW); cw[2] = lds_u1(ac - 4 * CBW + 4);
                            cw[3] = lds_u1(ac - 4); cw[4] = lds_u1(ac); cw[5] = lds_u1(ac + 4);
                            cw[6] = lds_u1(ac + 4 * CBW - 4); cw[7] = lds_u1(ac + 4 * CBW); cw[8] = lds_u1(ac + 4 * CBW + 4);
                        } else {
                            uint32_t c1x0, c1y0, n1x, n1y;
                            lds_u4(a_geo + j * (int)sizeof(AGeo) + 48, c1x0, c1y0, n1x, n1y);
                            const int c1x = (int)c1x0 + qx, c1y = (int)c1y0 + qy;
                            const int dxl = up_l(c1x, (int)n1x) - c1x, dxr = up_r(c1x, (int)n1x) - c1x;
                            const SAddr aa = ac + 4 * CBW * (up_l(c1y, (int)n1y) - c1y), ad = ac + 4 * CBW * (up_r(c1y, (int)n1y) - c1y);
                            cw[0] = lds_u1(aa + 4 * dxl); cw[1] = lds_u1(aa); cw[2] = lds_u1(aa + 4 * dxr);
                            cw[3] = lds_u1(ac + 4 * dxl); cw[4] = lds_u1(ac); cw[5] = lds_u1(ac + 4 * dxr);
                            cw[6] = lds_u1(ad + 4 * dxl); cw[7] = lds_u1(ad); cw[8] = lds_u1(ad + 4 * dxr);
                        }
                        lds_u2(ag, g0v[0], g0v[1]); lds_u2(ag + 4 * GBW, g0v[2], g0v[3]);
                        if (!gones || !DS_CUDA) { lds_f2(ag + GBOX, wv4[0], wv4[1]); lds_f2(ag + GBOX + 4 * GBW, wv4[2], wv4[3]); }
#if !DS_CUDA
                        if (gones && (wv4[0] != 1.f || wv4[1] != 1.f || wv4[2] != 1.f || wv4[3] != 1.f)) { fprintf(stderr, "ds emu: level %d tile %d: weights claimed 1 by geometry are not\n", p.level, tile); abort(); }
#endif
                        quad(cw, g0v, wv4, gones != 0, pbr[k], pg[k], ws[k]);
                    }
                    m_use++; lm_++;
                    if (lm_ + NST - 1 < nl) {   // live entry lm_ - 1 + NST exists: it goes into the stage just consumed
                        DS_SYNC();
                        if (tid == 0) { fence_proxy_async_smem(); issue_one(s); }
                        m_issue++;
                    }
                }
            } else {
            // ---- the top level (accumulates G_L itself, per pixel: its ROI need not be even-aligned) and, without tensor
            // maps, the levels below it with direct loads
            for (int j = 0; j < n; j++) {
                const AGeo& g = s_geo[j];
                if (g.skip) continue;   // block-uniform
                DS_UNROLL
                for (int k = 0; k < K; k++) {
                    const int q = tid + k * NT;
                    const int qx = q % QW, qy = q / QW;
                    if (top) {
                        DS_UNROLL
                        for (int i = 0; i < 4; i++) {
                            const int xp = 2 * qx + (i & 1), yp = 2 * qy + (i >> 1);
                            if ((unsigned)(xp - g.x0) >= (unsigned)g.nx || (unsigned)(yp - g.y0) >= (unsigned)g.ny) continue;
                            const int gi = yp * g.gp + xp;
                            const uint32_t v = ld_ro(g.G + gi);
                            const float wv = ld_ro(g.W + gi);
#if !DS_CUDA
                            if (g.ones && wv != 1.f) { fprintf(stderr, "ds emu: level %d tile %d: weights claimed 1 by geometry are not\n", p.level, tile); abort(); }
#endif
                            const int tb = (int)(short)f2i_rz(f_mul((float)(v & 255u), wv)), tg = (int)(short)f2i_rz(f_mul((float)((v >> 8) & 255u), wv));
                            const int tr = (int)(short)f2i_rz(f_mul((float)((v >> 16) & 255u), wv));
                            pbr[k][i] += tb + tr * 65536; pg[k][i] += tg;
                            ws[k][i] = f_add(ws[k][i], wv);
                        }
                        continue;
                    }
                    if ((unsigned)(qx - g.x0) >= (unsigned)g.nx || (unsigned)(qy - g.y0) >= (unsigned)g.ny) continue;
                    uint32_t cw[9], g0v[4]; float wv4[4] = {1.f, 1.f, 1.f, 1.f};
                    int dxl = -1, dxr = 1, dyl = -1, dyr = 1;   // pyrUp neighbours, relative to the quad's coarse pixel
                    if (g.border) {
                        const int c1x = g.c1x0 + qx, c1y = g.c1y0 + qy;
                        dxl = up_l(c1x, g.n1x) - c1x; dxr = up_r(c1x, g.n1x) - c1x; dyl = up_l(c1y, g.n1y) - c1y; dyr = up_r(c1y, g.n1y) - c1y;
                    }
                    const uint32_t* const g0p = g.G + (2 * qy * g.gp + 2 * qx);
                    const uint32_t* const c = g.G1 + (qy * g.gp1 + qx);
                    const uint32_t* const ca = c + dyl * g.gp1; const uint32_t* const cd = c + dyr * g.gp1;
                    // every load of the quad is requested before the first use
                    cw[0] = ld_ro(ca + dxl); cw[1] = ld_ro(ca); cw[2] = ld_ro(ca + dxr);
                    cw[3] = ld_ro(c + dxl); cw[4] = ld_ro(c); cw[5] = ld_ro(c + dxr);
                    cw[6] = ld_ro(cd + dxl); cw[7] = ld_ro(cd); cw[8] = ld_ro(cd + dxr);
                    const uint2 gr0 = ld_ro((const uint2*)g0p), gr1 = ld_ro((const uint2*)(g0p + g.gp));
                    g0v[0] = gr0.x; g0v[1] = gr0.y; g0v[2] = gr1.x; g0v[3] = gr1.y;
                    if (!g.ones || !DS_CUDA) {
                        const float* wp = g.W + (2 * qy * g.gp + 2 * qx);
                        const float2 wr0 = ld_ro((const float2*)wp), wr1 = ld_ro((const float2*)(wp + g.gp));
                        wv4[0] = wr0.x; wv4[1] = wr0.y; wv4[2] = wr1.x; wv4[3] = wr1.y;
                    }
#if !DS_CUDA
                    if (g.ones && (wv4[0] != 1.f || wv4[1] != 1.f || wv4[2] != 1.f || wv4[3] != 1.f)) { fprintf(stderr, "ds emu: level %d tile %d: weights claimed 1 by geometry are not\n", p.level, tile); abort(); }
#endif
                    quad(cw, g0v, wv4, g.ones != 0, pbr[k], pg[k], ws[k]);
                }
            }
            }
            // fold the chunk's packed sums into the wide ones
            DS_UNROLL
            for (int k = 0; k < K; k++) {
                DS_UNROLL
                for (int i = 0; i < 4; i++) {
                    const int cb = (int)(short)(pbr[k][i] & 0xFFFF);
                    sB[k][i] += cb; sR[k][i] += (pbr[k][i] - cb) >> 16;
                    pbr[k][i] = 0;
                }
            }
        }
        // ---- normalise and store
        DS_UNROLL
        for (int k = 0; k < K; k++) {
            const int q = tid + k * NT;
            const int X = X0 + 2 * (q % QW), Y = Y0 + 2 * (q / QW);
            DS_UNROLL
            for (int dy = 0; dy < 2; dy++) {
                const int Yp = Y + dy;
                if (Yp >= p.dst_h || Yp < p.acc_y0 || Yp >= p.acc_y1) continue;
                uint32_t w0[2], w1[2];
                DS_UNROLL
                for (int dx = 0; dx < 2; dx++) { const int i = 2 * dy + dx; norm_px(sB[k][i], pg[k][i], sR[k][i], ws[k][i], w0[dx], w1[dx]); }
                px16* q16 = p.dst + (size_t)Yp * p.dst_w + X;
                if (X + 1 < p.dst_w && !(p.dst_w & 1)) *(uint4*)q16 = make_u4(w0[0], w1[0], w0[1], w1[1]);   // X even, even pitch: 16-byte aligned
                else {
                    if (X < p.dst_w) { uint2 v; v.x = w0[0]; v.y = w1[0]; *(uint2*)q16 = v; }
                    if (X + 1 < p.dst_w) { uint2 v; v.x = w0[1]; v.y = w1[1]; *(uint2*)(q16 + 1) = v; }
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// MULTIBAND collapse of one level: fine = sat(pyrUp(coarse) + fine). The last step (fine = level 0)
// also crops to the unpadded canvas, applies the result mask and saturates to 8U (A11 blend).
// Each work item is 4 horizontally adjacent fine pixels.

struct CollapseParams {
    const px16* coarse; int cw, ch;
    px16* fine; int fw, fh;
    int y0, y1;   // fine rows to process
    int final;    // 1: write `o` instead of updating `fine`
    OutParams o;
};
struct CollapseBody {
    static constexpr int PER_BLOCK = 256;  // items per block; one item = 4 columns x 2 rows (an even / odd row pair)
    static int smem_bytes() { return 0; }
    static long long items(const CollapseParams& p) {
        const int m0 = p.y0 >> 1, m1 = (p.y1 + 1) >> 1;   // row pairs (2m, 2m+1) touching [y0, y1)
        return (long long)((p.fw + 3) / 4) * (m1 - m0);
    }

    // one pyrUp output from the horizontally filtered rows (A10), plus the fine value, saturated
    // (vv + 32) >> 6 of 16-bit inputs always fits 16 bits (the taps sum to 64), so pyrUp's cast to short changes nothing.
    // FINAL: the result is only saturated to 8 bits afterwards, which subsumes the 16-bit saturation of the add.
    template <bool FINAL>
    DS_DM int up1(int hl, int hc, int hr, bool odd_y, int fine) {
        const int vv = odd_y ? 4 * (hc + hr) : (hl + 6 * hc + hr);
        const int v = ((vv + 32) >> 6) + fine;
        return FINAL ? v : sat16i(v);
    }

    // ---- interior items: no pyrUp border rule applies, four valid fine columns, coarse columns c0-1 .. c0+2 readable as
    // 8 + 16 + 8 bytes per row. Straight-line arithmetic with the scale factors of the odd taps folded into the final
    // shifts: with E = l + 6c + r and S = c + r per coarse row (horizontal pass, odd columns would be 4S),
    //   even column, even row: (E0 + 6 E1 + E2 + 32) >> 6        even column, odd row: (4 (E1 + E2) + 32) >> 6 = (E1 + E2 + 8) >> 4
    //   odd column,  even row: (4 (S0 + 6 S1 + S2) + 32) >> 6 = (S0 + 6 S1 + S2 + 8) >> 4
    //   odd column,  odd row:  (16 (S1 + S2) + 32) >> 6 = (S1 + S2 + 2) >> 2          (arithmetic shifts: exact for negatives)
    DS_DM int sx16(uint32_t w) { return (int)(short)(w & 0xffffu); }
    template <bool FINAL>
    DS_DM int fin(int v) {
#if DS_CUDA
        if (FINAL) return __vimin_s32_relu(v, 255);   // min(max(v, 0), 255) in one instruction; subsumes the 16-bit saturation
#endif
        return FINAL ? sat8i(v) : sat16i(v);
    }
    // one channel: v[row][col] coarse values, f[dy][kx] fine values -> o[dy][kx]
    template <bool FINAL>
    DS_DM void chan(const int (&v)[3][4], const int (&f)[2][4], int (&o)[2][4]) {
        int E0[3], E1[3], S0[3], S1[3];
        DS_UNROLL
        for (int r = 0; r < 3; r++) {
            E0[r] = v[r][0] + v[r][2] + 6 * v[r][1]; S0[r] = v[r][1] + v[r][2];
            E1[r] = v[r][1] + v[r][3] + 6 * v[r][2]; S1[r] = v[r][2] + v[r][3];
        }
        o[0][0] = fin<FINAL>(((E0[0] + E0[2] + 32 + 6 * E0[1]) >> 6) + f[0][0]);
        o[1][0] = fin<FINAL>(((E0[1] + E0[2] + 8) >> 4) + f[1][0]);
        o[0][1] = fin<FINAL>(((S0[0] + S0[2] + 8 + 6 * S0[1]) >> 4) + f[0][1]);
        o[1][1] = fin<FINAL>(((S0[1] + S0[2] + 2) >> 2) + f[1][1]);
        o[0][2] = fin<FINAL>(((E1[0] + E1[2] + 32 + 6 * E1[1]) >> 6) + f[0][2]);
        o[1][2] = fin<FINAL>(((E1[1] + E1[2] + 8) >> 4) + f[1][2]);
        o[0][3] = fin<FINAL>(((S1[0] + S1[2] + 8 + 6 * S1[1]) >> 4) + f[0][3]);
        o[1][3] = fin<FINAL>(((S1[1] + S1[2] + 2) >> 2) + f[1][3]);
    }
    template <bool FINAL>
    DS_DM void item_interior(const CollapseParams& p, int c1y, int Xq) {
        const int cw = p.cw, fw = p.fw;
        const px16* cq = p.coarse + (size_t)(c1y - 1) * cw + ((Xq >> 1) - 1);
        uint32_t w0[3][4], w1[3][4];   // coarse pixels as words: b | g << 16, r | a << 16
        DS_UNROLL
        for (int r = 0; r < 3; r++) {
            const px16* q = cq + (size_t)r * cw;
            const uint2 a = *(const uint2*)q; const uint4 b = *(const uint4*)(q + 1); const uint2 c = *(const uint2*)(q + 3);
            w0[r][0] = a.x; w1[r][0] = a.y; w0[r][1] = b.x; w1[r][1] = b.y; w0[r][2] = b.z; w1[r][2] = b.w; w0[r][3] = c.x; w1[r][3] = c.y;
        }
        uint32_t g0[2][4], g1[2][4];   // fine pixels, same word layout
        bool live[2];
        DS_UNROLL
        for (int dy = 0; dy < 2; dy++) {
            const int Y = 2 * c1y + dy;
            live[dy] = Y >= p.y0 && Y < p.y1;
            uint4 u0 = make_u4(0u, 0u, 0u, 0u), u1 = u0;
            if (live[dy]) { const px16* frow = p.fine + (size_t)Y * fw + Xq; u0 = *(const uint4*)frow; u1 = *(const uint4*)(frow + 2); }
            g0[dy][0] = u0.x; g1[dy][0] = u0.y; g0[dy][1] = u0.z; g1[dy][1] = u0.w;
            g0[dy][2] = u1.x; g1[dy][2] = u1.y; g0[dy][3] = u1.z; g1[dy][3] = u1.w;
        }
        int v[3][4], f[2][4], ob[2][4], og[2][4], orr[2][4];
        DS_UNROLL
        for (int r = 0; r < 3; r++) { DS_UNROLL for (int k = 0; k < 4; k++) v[r][k] = sx16(w0[r][k]); }
        DS_UNROLL
        for (int dy = 0; dy < 2; dy++) { DS_UNROLL for (int k = 0; k < 4; k++) f[dy][k] = sx16(g0[dy][k]); }
        chan<FINAL>(v, f, ob);
        DS_UNROLL
        for (int r = 0; r < 3; r++) { DS_UNROLL for (int k = 0; k < 4; k++) v[r][k] = (int)w0[r][k] >> 16; }
        DS_UNROLL
        for (int dy = 0; dy < 2; dy++) { DS_UNROLL for (int k = 0; k < 4; k++) f[dy][k] = (int)g0[dy][k] >> 16; }
        chan<FINAL>(v, f, og);
        DS_UNROLL
        for (int r = 0; r < 3; r++) { DS_UNROLL for (int k = 0; k < 4; k++) v[r][k] = sx16(w1[r][k]); }
        DS_UNROLL
        for (int dy = 0; dy < 2; dy++) { DS_UNROLL for (int k = 0; k < 4; k++) f[dy][k] = sx16(g1[dy][k]); }
        chan<FINAL>(v, f, orr);
        DS_UNROLL
        for (int dy = 0; dy < 2; dy++) {
            const int Y = 2 * c1y + dy;
            if (!live[dy]) continue;
            if (!FINAL) {
                px16* frow = p.fine + (size_t)Y * fw + Xq;
                uint32_t a0[4], a1[4];
                DS_UNROLL
                for (int k = 0; k < 4; k++) {
                    a0[k] = ((uint32_t)ob[dy][k] & 0xffffu) | ((uint32_t)og[dy][k] << 16);
                    a1[k] = ((uint32_t)orr[dy][k] & 0xffffu) | (g1[dy][k] & 0xffff0000u);
                }
                *(uint4*)frow = make_u4(a0[0], a1[0], a0[1], a1[1]);
                *(uint4*)(frow + 2) = make_u4(a0[2], a1[2], a0[3], a1[3]);
            } else if (Y < p.o.h) {
                uint32_t px[4];
                DS_UNROLL
                for (int k = 0; k < 4; k++) {
                    const bool m = (g1[dy][k] >> 16) != 0u;
                    px[k] = m ? ((uint32_t)ob[dy][k] | ((uint32_t)og[dy][k] << 8) | ((uint32_t)orr[dy][k] << 16) | 0xff000000u) : 0u;
                }
                if (p.o.fmt == 1) {
                    *(uint4*)((uint32_t*)(p.o.out + (size_t)Y * p.o.out_pitch) + Xq) = make_u4(px[0], px[1], px[2], px[3]);
                } else {
                    uint32_t* q = (uint32_t*)(p.o.out + (size_t)Y * p.o.out_pitch + (size_t)Xq * 3);
                    q[0] = (px[0] & 0xffffffu) | (px[1] << 24);
                    q[1] = ((px[1] >> 8) & 0xffffu) | (px[2] << 16);
                    q[2] = ((px[2] >> 16) & 0xffu) | (px[3] << 8);
                    *(uint32_t*)(p.o.mask + (size_t)Y * p.o.mask_pitch + Xq) =
                        (px[0] >> 24) | ((px[1] >> 24) << 8) | ((px[2] >> 24) << 16) | ((px[3] >> 24) << 24);
                }
            }
        }
    }

    template <int NT>
    DS_DM void run(const CollapseParams& p, int block, int tid, unsigned char*) {
        const int qw = (p.fw + 3) / 4;
        const int m0 = p.y0 >> 1, m1 = (p.y1 + 1) >> 1;
        const long long n = (long long)qw * (m1 - m0);
        const px16* const coarse = p.coarse;
        const int cw = p.cw, chh = p.ch, fw = p.fw;
        for (int it = tid; it < PER_BLOCK; it += NT) {
            const long long idx = (long long)block * PER_BLOCK + it;
            if (idx >= n) break;
            // 32-bit index arithmetic whenever the launch allows it (a 64-bit division costs ~80 instructions)
            int rowi, coli;
            if (n <= 0x7fffffffLL) { const unsigned u = (unsigned)idx; rowi = (int)(u / (unsigned)qw); coli = (int)(u - (unsigned)rowi * (unsigned)qw); }
            else { rowi = (int)(idx / qw); coli = (int)(idx - (long long)rowi * qw); }
            const int c1y = m0 + rowi, Xq = coli * 4;
            {
                const int c0i = Xq >> 1;
                if (c0i >= 1 && c0i + 2 <= cw - 1 && c1y >= 1 && c1y + 1 <= chh - 1 && fw - Xq >= 4 && !(cw & 1) && !(fw & 1) &&
                    (!p.final || Xq + 4 <= p.o.w)) {
                    if (p.final) item_interior<true>(p, c1y, Xq); else item_interior<false>(p, c1y, Xq);
                    continue;
                }
            }
            // both rows of the pair read coarse rows (l, c, r) around c1y; columns c0-1 .. c0+2 with the pyrUp border rules
            const int ry[3] = {up_l(c1y, chh), c1y, up_r(c1y, chh)};
            const int c0 = Xq >> 1;
            const int cx[4] = {up_l(c0, cw), c0, up_r(c0, cw), up_r(imin(c0 + 1, cw - 1), cw)};
            int cb[3][4], cg[3][4], cr[3][4];
            DS_UNROLL
            for (int r = 0; r < 3; r++) {
                const px16* row = coarse + (size_t)ry[r] * cw;
                DS_UNROLL
                for (int c = 0; c < 4; c++) { const px16 q = row[cx[c]]; cb[r][c] = q.b; cg[r][c] = q.g; cr[r][c] = q.r; }
            }
            // horizontal pass once for the 4 columns x 3 coarse rows
            int hb[3][4], hg[3][4], hr[3][4];
            DS_UNROLL
            for (int kx = 0; kx < 4; kx++) {
                const int j = kx >> 1;   // even X: (l, c, r) = columns (j, j+1, j+2); odd X: (c, r) = (j+1, j+2)
                DS_UNROLL
                for (int r = 0; r < 3; r++) {
                    if (kx & 1) {
                        hb[r][kx] = 4 * (cb[r][j + 1] + cb[r][j + 2]); hg[r][kx] = 4 * (cg[r][j + 1] + cg[r][j + 2]);
                        hr[r][kx] = 4 * (cr[r][j + 1] + cr[r][j + 2]);
                    } else {
                        hb[r][kx] = cb[r][j] + 6 * cb[r][j + 1] + cb[r][j + 2]; hg[r][kx] = cg[r][j] + 6 * cg[r][j + 1] + cg[r][j + 2];
                        hr[r][kx] = cr[r][j] + 6 * cr[r][j + 1] + cr[r][j + 2];
                    }
                }
            }
            const int nvalid = imin(4, fw - Xq);
            // the fine pixels of BOTH rows are requested before anything is stored: the stores below may alias the
            // loads as far as the compiler knows, so a load placed after them would wait for the first row to finish
            uint4 fv[2][2];
            bool live[2];
            DS_UNROLL
            for (int dy = 0; dy < 2; dy++) {
                const int Y = 2 * c1y + dy;
                live[dy] = Y >= p.y0 && Y < p.y1;
                fv[dy][0] = fv[dy][1] = make_u4(0u, 0u, 0u, 0u);
                if (live[dy] && nvalid == 4) {
                    const px16* frow = p.fine + (size_t)Y * fw + Xq;
                    fv[dy][0] = *(const uint4*)frow; fv[dy][1] = *(const uint4*)(frow + 2);
                }
            }
            DS_UNROLL
            for (int dy = 0; dy < 2; dy++) {
                const int Y = 2 * c1y + dy;
                if (!live[dy]) continue;
                const bool oddy = dy != 0;
                // the four fine pixels as words (b | g << 16, r | flag << 16): kept in registers, never as an
                // addressable array (that would live in local memory)
                uint32_t fw0[4], fw1[4];
                px16* frow = p.fine + (size_t)Y * fw + Xq;
                if (nvalid == 4) {
                    const uint4 v0 = fv[dy][0], v1 = fv[dy][1];
                    fw0[0] = v0.x; fw1[0] = v0.y; fw0[1] = v0.z; fw1[1] = v0.w;
                    fw0[2] = v1.x; fw1[2] = v1.y; fw0[3] = v1.z; fw1[3] = v1.w;
                } else {
                    DS_UNROLL
                    for (int kx = 0; kx < 4; kx++) {
                        fw0[kx] = fw1[kx] = 0u;
                        if (kx < nvalid) { const uint2 v = *(const uint2*)(frow + kx); fw0[kx] = v.x; fw1[kx] = v.y; }
                    }
                }
                int ob[4], og[4], orr[4];
                DS_UNROLL
                for (int kx = 0; kx < 4; kx++) {
                    const int fb = (int)(short)(fw0[kx] & 0xffffu), fg = (int)(fw0[kx]) >> 16, fr = (int)(short)(fw1[kx] & 0xffffu);
                    if (p.final) {
                        ob[kx] = up1<true>(hb[0][kx], hb[1][kx], hb[2][kx], oddy, fb);
                        og[kx] = up1<true>(hg[0][kx], hg[1][kx], hg[2][kx], oddy, fg);
                        orr[kx] = up1<true>(hr[0][kx], hr[1][kx], hr[2][kx], oddy, fr);
                    } else {
                        ob[kx] = up1<false>(hb[0][kx], hb[1][kx], hb[2][kx], oddy, fb);
                        og[kx] = up1<false>(hg[0][kx], hg[1][kx], hg[2][kx], oddy, fg);
                        orr[kx] = up1<false>(hr[0][kx], hr[1][kx], hr[2][kx], oddy, fr);
                    }
                }
                if (!p.final) {
                    DS_UNROLL
                    for (int kx = 0; kx < 4; kx++) {
                        fw0[kx] = ((uint32_t)ob[kx] & 0xffffu) | ((uint32_t)og[kx] << 16);
                        fw1[kx] = ((uint32_t)orr[kx] & 0xffffu) | (fw1[kx] & 0xffff0000u);
                    }
                    if (nvalid == 4) {
                        *(uint4*)frow = make_u4(fw0[0], fw1[0], fw0[1], fw1[1]);
                        *(uint4*)(frow + 2) = make_u4(fw0[2], fw1[2], fw0[3], fw1[3]);
                    } else {
                        DS_UNROLL
                        for (int kx = 0; kx < 4; kx++)
                            if (kx < nvalid) { uint2 v; v.x = fw0[kx]; v.y = fw1[kx]; *(uint2*)(frow + kx) = v; }
                    }
                } else if (Y < p.o.h) {
                    uint32_t px[4];
                    DS_UNROLL
                    for (int kx = 0; kx < 4; kx++) {
                        const bool m = (fw1[kx] >> 16) != 0u;
                        px[kx] = m ? ((uint32_t)sat8i(ob[kx]) | ((uint32_t)sat8i(og[kx]) << 8) | ((uint32_t)sat8i(orr[kx]) << 16) | 0xff000000u) : 0u;
                    }
                    const int nout = imin(nvalid, p.o.w - Xq);
                    if (p.o.fmt == 1) {
                        uint32_t* q = (uint32_t*)(p.o.out + (size_t)Y * p.o.out_pitch) + Xq;
                        if (nout == 4) *(uint4*)q = make_u4(px[0], px[1], px[2], px[3]);
                        else for (int kx = 0; kx < nout; kx++) q[kx] = px[kx];
                    } else if (nout == 4) {
                        // 12 BGR bytes = 3 aligned words; 4 mask bytes = 1 word (Xq is a multiple of 4, pitches of 256)
                        uint32_t* q = (uint32_t*)(p.o.out + (size_t)Y * p.o.out_pitch + (size_t)Xq * 3);
                        q[0] = (px[0] & 0xffffffu) | (px[1] << 24);
                        q[1] = ((px[1] >> 8) & 0xffffu) | (px[2] << 16);
                        q[2] = ((px[2] >> 16) & 0xffu) | (px[3] << 8);
                        *(uint32_t*)(p.o.mask + (size_t)Y * p.o.mask_pitch + Xq) =
                            (px[0] >> 24) | ((px[1] >> 24) << 8) | ((px[2] >> 24) << 16) | ((px[3] >> 24) << 24);
                    } else {
                        for (int kx = 0; kx < nout; kx++) {
                            uint8_t* q = p.o.out + (size_t)Y * p.o.out_pitch + (size_t)(Xq + kx) * 3;
                            q[0] = (uint8_t)px[kx]; q[1] = (uint8_t)(px[kx] >> 8); q[2] = (uint8_t)(px[kx] >> 16);
                            p.o.mask[(size_t)Y * p.o.mask_pitch + Xq + kx] = (uint8_t)(px[kx] >> 24);
                        }
                    }
                }
            }
        }
    }
};

// MULTIBAND with zero bands (L == 0): level 0 is the top level; the normalised level IS the result.
struct FinalizeL0Params {
    const px16* lvl0; int fw, fh; int y0, y1;
    OutParams o;
};
struct FinalizeL0Body {
    static constexpr int PER_BLOCK = 1024;
    static int smem_bytes() { return 0; }
    template <int NT>
    DS_DM void run(const FinalizeL0Params& p, int block, int tid, unsigned char*) {
        const long long n = (long long)p.fw * (p.y1 - p.y0);
        for (int it = tid; it < PER_BLOCK; it += NT) {
            const long long idx = (long long)block * PER_BLOCK + it;
            if (idx >= n) break;
            const int Y = p.y0 + (int)(idx / p.fw), X = (int)(idx % p.fw);
            if (X >= p.o.w || Y >= p.o.h) continue;
            const px16 f = p.lvl0[(size_t)Y * p.fw + X];
            const int m = f.a != 0;
            const int ob = m ? sat8i(f.b) : 0, og = m ? sat8i(f.g) : 0, orr = m ? sat8i(f.r) : 0;
            if (p.o.fmt == 1) {
                *(uint32_t*)(p.o.out + (size_t)Y * p.o.out_pitch + (size_t)X * 4) =
                    (uint32_t)ob | ((uint32_t)og << 8) | ((uint32_t)orr << 16) | (m ? 0xff000000u : 0u);
            } else {
                uint8_t* q = p.o.out + (size_t)Y * p.o.out_pitch + (size_t)X * 3;
                q[0] = (uint8_t)ob; q[1] = (uint8_t)og; q[2] = (uint8_t)orr;
                p.o.mask[(size_t)Y * p.o.mask_pitch + X] = m ? 255 : 0;
            }
        }
    }
};

#include "ds_mask_kernels.h"

#if DS_CUDA
typedef MBBody<64, true> MBBodyL0;
DS_DEFINE_KERNEL(ds_expand_bgrx, ExpandBody, 256, ExpandParams, 1)
DS_DEFINE_KERNEL(ds_meta_copy, MetaCopyBody, 256, MetaCopyParams, 1)
DS_DEFINE_KERNEL(ds_p2p_pull, PullBody, 256, PullParams, 1)
DS_DEFINE_KERNEL(ds_p2p_signal, SignalBody, 32, SignalParams, 1)
DS_DEFINE_KERNEL(ds_debug_tap, TapBody, 256, TapParams, 1)
DS_DEFINE_KERNEL(ds_seam_upsize, SeamUpBody, 256, SeamUpParams, 1)
DS_DEFINE_KERNEL(ds_mask_prep, MaskPrepBody, 256, MaskPrepParams, 1)
DS_DEFINE_KERNEL(ds_soft_mask, SoftMaskBody, 256, SoftMaskParams, 3)
DS_DEFINE_KERNEL(ds_crop_row_runs, RowRunsBody, 256, RowRunsParams, 1)
DS_DEFINE_KERNEL(ds_gain_resize, GainResizeBody, 256, GainResizeParams, 1)
DS_DEFINE_KERNEL(ds_feather_mask_bits, MaskBitsBody, 256, MaskBitsParams, 1)
DS_DEFINE_KERNEL(ds_feather_mask_rows, MaskBitsRowBody, 256, MaskBitsRowParams, 1)
DS_DEFINE_KERNEL(ds_feather_dist, FeatherDistBody, 256, FeatherDistParams, 1)
DS_DEFINE_KERNEL(ds_feather_blend, FeatherBody, 256, FeatherParams, 4)
DS_DEFINE_KERNEL(ds_mb_feed_l0_generic, MBBodyL0, 512, MBParams, 2)
typedef MBFastBody<64, true> MBFastL0;
typedef MBFastBody<64, true, true> MBFastL0A;
DS_DEFINE_KERNEL(ds_mb_feed_l0, MBFastL0, 512, MBParams, 2)
DS_DEFINE_KERNEL(ds_mb_feed_l0_affine, MBFastL0A, 512, MBParams, 2)
DS_DEFINE_KERNEL(ds_mb_pyrdown, PyrDownBody, 128, PyrParams, 5)
DS_DEFINE_KERNEL(ds_mb_accum, AccumBody, 128, MBParams, 8)
DS_DEFINE_KERNEL(ds_mb_collapse, CollapseBody, 256, CollapseParams, 4)
DS_DEFINE_KERNEL(ds_mb_finalize_l0, FinalizeL0Body, 256, FinalizeL0Params, 1)
#endif
