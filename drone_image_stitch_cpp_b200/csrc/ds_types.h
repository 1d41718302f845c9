// ds_types.h — POD structures shared by the host runtime and the kernels.
#pragma once
#include "ds_device.h"

#define DS_MAXL 13  // levels 0..12

enum { XF_PLANE = 0, XF_AFFINE = 1, XF_HOMOGRAPHY = 2 };
enum { BORDER_CONST = 0, BORDER_REFL = 1, BORDER_REFL101 = 2 };

// One uploaded frame as the kernels see it (array in device memory).
struct alignas(16) FrameDev {
    const uint32_t* src;  // BGRX pixels, row pitch in pixels
    int src_w, src_h, src_pitch;
    int kind, border;
    // PLANE_F32 (cv::detail::PlaneProjector::mapBackward): k = K * R^-1 (float32), t, scale.
    float k[9];
    float t0, t1, k2one, k5one, k8one, scale;
    // AFFINE_F64: inv = {m0, m1, b1, m3, m4, b2} (inverse 2x3); HOMOGRAPHY_F64: inv = H^-1 (3x3).
    double inv[9];
    int tlx, tly;        // PLANE: absolute plane coords of the bbox top-left
    int cx, cy;          // bbox top-left relative to the padded canvas origin
    int w, h;            // warped bbox size (cv sizes[i])
    int rx, ry, rw, rh;  // MultiBandBlender::feed aligned ROI at level 0, relative to padded canvas origin
    px8* G[DS_MAXL];     // per-frame Gaussian levels 1..L over the feed ROI (values 0..255; index 0 unused)
    float* W[DS_MAXL];   // per-frame weight levels 1..L
    int gp[DS_MAXL];     // row pitch (elements) of G[l] / W[l]: (rw >> l) rounded up to 4 (16-B rows, TMA-able)
    const uint32_t* mbits;  // FEATHER: warped-mask bit plane over the bbox (1 bit/px, tail bits set)
    int mbits_pitch;        // words per row
    const uint8_t* seam;    // optional seam mask over the bbox
    int seam_pitch;
    float gain[3];          // applyChannelGainInPlace: float32 multiply (stitch_global.cpp:291-305)
    int has_gain;
    double cgain[3];        // ExposureCompensator::apply with scalar gains: float64 multiply (stitch_global.cpp:644)
    int has_cgain;
    const float* gainmap;   // BlocksGainCompensator::apply: per-pixel float32 gain over the bbox (stitch_robust.cpp:209-211)
    int gainmap_pitch;
    int any_gain;           // has_gain | has_cgain | (gainmap != 0)
    const uint8_t* dist;    // FEATHER: min(L1 distance to the nearest zero of the warped mask, 255) over the bbox, exact below the radius
    int dist_pitch;
};

struct Coord {
    int sx, sy;  // integer source coordinate (sat16)
    int ax, ay;  // 5-bit fractions
    int m;       // nearest-warped all-255 mask: 255 / 0
};
