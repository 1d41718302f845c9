// ds_geometry.h — host-side placement geometry of the compositing path (float32 / int, no device).
//
// What the caller of the reference gets from OpenCV before the hot loop starts:
//   - warper.warpRoi(size, K, R)            (inside composePanorama, called at
//                                            /root/reference/src/stitch_robust.cpp:256)
//   - MultiBandBlender::prepare / feed ROI  (/root/reference/src/stitch_global.cpp:636-638, :658)
// The float32 op order is part of the contract: corners and sizes must equal OpenCV's exactly or
// every pyramid phase shifts (SURVEY.md "Hard parts"). Host code in this TU is compiled without
// FMA contraction (x86-64 baseline / -ffp-contract=off).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace dsgeo {

struct Projector {
    float k_rinv[9], r_kinv[9], t[3];
};

inline void mat3_mul_f32(const float* A, const float* B, float* C) {
    // cv::Mat_<float> 3x3 product: one float32 accumulator per element, k ascending
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float acc = 0.f;
            for (int k = 0; k < 3; k++) acc = acc + A[3 * i + k] * B[3 * k + j];
            C[3 * i + j] = acc;
        }
}

inline void mat3_inv_f32(const float* m, float* out) {
    // cv::invert on a 3x3 CV_32F: cofactors and determinant in double, result narrowed to float
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == 0.0) { memset(out, 0, 9 * sizeof(float)); return; }
    det = 1.0 / det;
    out[0] = (float)((e * i - f * h) * det); out[1] = (float)((c * h - b * i) * det); out[2] = (float)((b * f - c * e) * det);
    out[3] = (float)((f * g - d * i) * det); out[4] = (float)((a * i - c * g) * det); out[5] = (float)((c * d - a * f) * det);
    out[6] = (float)((d * h - e * g) * det); out[7] = (float)((b * g - a * h) * det); out[8] = (float)((a * e - b * d) * det);
}

// ProjectorBase::setCameraParams after AffineWarper::getRTfromHomogeneous (affine != 0) or with
// T = 0 (PlaneWarper).
inline Projector make_projector(const float* K, const float* R_in, bool affine) {
    Projector P;
    float R[9], T[3] = {0.f, 0.f, 0.f};
    memcpy(R, R_in, sizeof(R));
    if (affine) {
        const float tx = R[2], ty = R[5];
        R[2] = 0.f; R[5] = 0.f;
        float Rt[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]};
        memcpy(R, Rt, sizeof(R));
        const float tv[3] = {tx, ty, 0.f};
        for (int r = 0; r < 3; r++) {
            float acc = 0.f;
            for (int k = 0; k < 3; k++) acc = acc + R[3 * r + k] * tv[k];
            T[r] = acc * -1.f;
        }
    }
    const float Rinv[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]};
    float Kinv[9];
    mat3_inv_f32(K, Kinv);
    mat3_mul_f32(R, Kinv, P.r_kinv);
    mat3_mul_f32(K, Rinv, P.k_rinv);
    P.t[0] = T[0]; P.t[1] = T[1]; P.t[2] = T[2];
    return P;
}

inline void map_forward(const Projector& P, float scale, float x, float y, float& u, float& v) {
    float x_ = P.r_kinv[0] * x + P.r_kinv[1] * y + P.r_kinv[2];
    float y_ = P.r_kinv[3] * x + P.r_kinv[4] * y + P.r_kinv[5];
    float z_ = P.r_kinv[6] * x + P.r_kinv[7] * y + P.r_kinv[8];
    x_ = P.t[0] + x_ / z_ * (1 - P.t[2]);
    y_ = P.t[1] + y_ / z_ * (1 - P.t[2]);
    u = scale * x_;
    v = scale * y_;
}

// PlaneWarper::detectResultRoi: project the 4 corners, truncate. Returns tl and inclusive br.
inline void plane_result_roi(const Projector& P, float scale, int w, int h, int& tlx, int& tly, int& brx, int& bry) {
    const float xs[4] = {0.f, 0.f, (float)(w - 1), (float)(w - 1)};
    const float ys[4] = {0.f, (float)(h - 1), 0.f, (float)(h - 1)};
    float lo_u = 3.402823466e+38f, lo_v = 3.402823466e+38f, hi_u = -3.402823466e+38f, hi_v = -3.402823466e+38f;
    for (int i = 0; i < 4; i++) {
        float u, v;
        map_forward(P, scale, xs[i], ys[i], u, v);
        lo_u = fminf(lo_u, u); lo_v = fminf(lo_v, v);
        hi_u = fmaxf(hi_u, u); hi_v = fmaxf(hi_v, v);
    }
    tlx = (int)lo_u; tly = (int)lo_v; brx = (int)hi_u; bry = (int)hi_v;
}

// MultiBandBlender::prepare band cropping: min(requested, ceil(log2(max(w, h)))).
inline int effective_bands(int requested, int w, int h) {
    const double max_len = (double)(w > h ? w : h);
    const int cap = (int)ceil(log(max_len) / log(2.0));
    int b = requested < cap ? requested : cap;
    return b < 0 ? 0 : b;
}

inline int pad_to(int v, int m) { return v + ((m - v % m) % m); }

// MultiBandBlender::feed: the aligned ROI a frame occupies in the padded canvas.
// canvas (cx, cy, cw, ch) is the padded dst_roi_. Result relative to the canvas origin.
inline void feed_roi(int cx, int cy, int cw, int ch, int bands, int tlx, int tly, int iw, int ih,
                     int& rx, int& ry, int& rw, int& rh) {
    const int gap = 3 * (1 << bands);
    const int cbx = cx + cw, cby = cy + ch;
    int x0 = tlx - gap > cx ? tlx - gap : cx;
    int y0 = tly - gap > cy ? tly - gap : cy;
    int x1 = tlx + iw + gap < cbx ? tlx + iw + gap : cbx;
    int y1 = tly + ih + gap < cby ? tly + ih + gap : cby;
    x0 = cx + (((x0 - cx) >> bands) << bands);
    y0 = cy + (((y0 - cy) >> bands) << bands);
    int width = pad_to(x1 - x0, 1 << bands), height = pad_to(y1 - y0, 1 << bands);
    x1 = x0 + width; y1 = y0 + height;
    const int dx = x1 - cbx > 0 ? x1 - cbx : 0, dy = y1 - cby > 0 ? y1 - cby : 0;
    x0 -= dx; y0 -= dy;
    rx = x0 - cx; ry = y0 - cy; rw = width; rh = height;
}

inline bool invert_affine_f64(const double* M, double* o /* m0 m1 b1 m3 m4 b2 */) {
    // cv::invertAffineTransform as used by cv::warpAffine
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = M[4] * D, A22 = M[0] * D;
    const double m0 = A11, m1 = M[1] * (-D), m3 = M[3] * (-D), m4 = A22;
    o[0] = m0; o[1] = m1; o[2] = -m0 * M[2] - m1 * M[5];
    o[3] = m3; o[4] = m4; o[5] = -m3 * M[2] - m4 * M[5];
    return D != 0;
}

inline bool invert_3x3_f64(const double* m, double* o) {
    double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (d == 0.) return false;
    d = 1. / d;
    o[0] = (m[4] * m[8] - m[5] * m[7]) * d; o[1] = (m[2] * m[7] - m[1] * m[8]) * d; o[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    o[3] = (m[5] * m[6] - m[3] * m[8]) * d; o[4] = (m[0] * m[8] - m[2] * m[6]) * d; o[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    o[6] = (m[3] * m[7] - m[4] * m[6]) * d; o[7] = (m[1] * m[6] - m[0] * m[7]) * d; o[8] = (m[0] * m[4] - m[1] * m[3]) * d;
    return true;
}

}  // namespace dsgeo
