// dronestitch.hpp - C++17 host side above the C ABI (include/dronestitch.h), header only.
//
// The reference is a C++ application (C++20 + OpenCV); this header is what its two compositing call sites
// include. It mirrors the reference's own vocabulary - cameras with K() and R as cv::detail::CameraParams has
// them, StitchTuning's compose-related fields (src/stitch_config.hpp:50-100), a Blender-like prepare / feed /
// blend object (cv::detail::Blender, src/stitch_global.cpp:632-666) - and needs no OpenCV headers itself: images
// are (data, cols, rows, step) views with cv::Mat's layout, so `ds::view(mat)` is a one-liner on the caller's
// side. Errors become ds::Error (std::runtime_error), which the reference's top-level handler
// already catches (src/stitch_app.cpp:265-268).
#pragma once
#include <algorithm>
#include <array>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "dronestitch.h"

namespace ds {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc) {
    if (rc != DS_OK) throw Error(rc, std::string("dronestitch: ") + ds_last_error());
}

// 8UC3 BGR (or 8UC1 for masks) image view with cv::Mat's layout: data, cols, rows, step in bytes.
struct ImageView {
    const uint8_t* data = nullptr;
    int cols = 0, rows = 0;
    size_t step = 0;
};
// Owning 8-bit image the results are returned in (rows * step bytes, step = cols * channels).
struct Image {
    std::vector<uint8_t> data;
    int cols = 0, rows = 0, channels = 3;
    size_t step() const { return (size_t)cols * channels; }
    void create(int r, int c, int ch) { rows = r; cols = c; channels = ch; data.assign((size_t)r * c * ch, 0); }
    ImageView view() const { return ImageView{data.data(), cols, rows, step()}; }
};

// cv::detail::CameraParams as cv::Stitcher::cameras() returns it after estimateTransform (stitch_robust.cpp:251).
struct CameraParams {
    double focal = 1, aspect = 1, ppx = 0, ppy = 0;
    std::array<float, 9> R{{1, 0, 0, 0, 1, 0, 0, 0, 1}};   // CV_32F 3x3; with the affine pipeline the 3x3 affine "H"
    // K() of cv::detail::CameraParams, converted to CV_32F as composePanorama does (K.convertTo(K, CV_32F))
    std::array<float, 9> K() const {
        return {{(float)focal, 0.f, (float)ppx, 0.f, (float)(focal * aspect), (float)ppy, 0.f, 0.f, 1.f}};
    }
};

// The compose-related fields of the reference's StitchTuning (src/stitch_config.hpp:50-100).
struct StitchTuning {
    bool use_affine_warper = true;   // cv::AffineWarper, else cv::PlaneWarper (stitch_robust.cpp:203-205)
    int blend_bands = 5;             // MultiBandBlender(try_gpu, blend_bands) (stitch_robust.cpp:213)
    bool feather = false;            // FeatherBlender(0.02f) instead (BASELINE config 1)
    float feather_sharpness = 0.02f;
};

inline ds_transform planeTransform(const CameraParams& cam, float warped_image_scale, bool affine_warper) {
    ds_transform t;
    std::memset(&t, 0, sizeof(t));
    t.kind = DS_XF_PLANE_F32;
    t.affine_warper = affine_warper ? 1 : 0;
    const std::array<float, 9> K = cam.K();
    std::memcpy(t.K, K.data(), sizeof(t.K));
    std::memcpy(t.R, cam.R.data(), sizeof(t.R));
    t.scale = warped_image_scale;
    t.border = DS_BORDER_REFLECT;
    return t;
}
// stitch_global.cpp:474-480: forward 2x3 (double) of a strip into its own bbox placed at `corner`
inline ds_transform affineTransform(const double M[6], int corner_x, int corner_y, int width, int height) {
    ds_transform t;
    std::memset(&t, 0, sizeof(t));
    t.kind = DS_XF_AFFINE_F64;
    t.border = DS_BORDER_CONSTANT;
    std::memcpy(t.M, M, 6 * sizeof(double));
    t.M[8] = 1.0;
    t.corner_x = corner_x; t.corner_y = corner_y; t.width = width; t.height = height;
    return t;
}

struct Rect { int x = 0, y = 0, width = 0, height = 0; };
// transformedBoundingRect(size, h) of the global stage (src/stitch_global.cpp:71-98): the four corners (0,0), (w,0),
// (w,h), (0,h) through the 3x3 double transform; floor(min), ceil(max) - floor(min), at least 1.
inline Rect transformedBoundingRect(int cols, int rows, const double H[9]) {
    const double px[4] = {0.0, (double)cols, (double)cols, 0.0}, py[4] = {0.0, 0.0, (double)rows, (double)rows};
    double min_x = 1.7976931348623157e308, min_y = min_x, max_x = -min_x, max_y = -min_x;
    for (int i = 0; i < 4; i++) {
        const double x = H[0] * px[i] + H[1] * py[i] + H[2] * 1.0, y = H[3] * px[i] + H[4] * py[i] + H[5] * 1.0;
        min_x = std::min(min_x, x); min_y = std::min(min_y, y);
        max_x = std::max(max_x, x); max_y = std::max(max_y, y);
    }
    const int x = (int)std::floor(min_x), y = (int)std::floor(min_y);
    return Rect{x, y, std::max(1, (int)std::ceil(max_x) - x), std::max(1, (int)std::ceil(max_y) - y)};
}
// Band count of the global stage's blender (src/stitch_global.cpp:632-635): auto = min(12, ceil(log2(max(w, h))) - 1),
// final = max(max(5, tuning.blend_bands), auto).
inline int globalBlendBands(int canvas_w, int canvas_h, int configured_bands) {
    const int auto_blend_bands = std::min(12, static_cast<int>(std::ceil(std::log2(static_cast<double>(std::max(canvas_w, canvas_h))))) - 1);
    return std::max(std::max(5, configured_bands), auto_blend_bands);
}
// cv::detail::resultRoi(corners, sizes)
inline Rect resultRoi(const std::vector<Rect>& placed) {
    int x0 = INT_MAX, y0 = INT_MAX, x1 = INT_MIN, y1 = INT_MIN;
    for (const Rect& r : placed) {
        x0 = std::min(x0, r.x); y0 = std::min(y0, r.y);
        x1 = std::max(x1, r.x + r.width); y1 = std::max(y1, r.y + r.height);
    }
    return Rect{x0, y0, x1 - x0, y1 - y0};
}
// warper->warpRoi(size, K, R)
inline Rect warpRoi(const ds_transform& t, int cols, int rows) {
    int32_t r[4];
    check(ds_warp_roi(&t, cols, rows, r));
    return Rect{r[0], r[1], r[2], r[3]};
}

// The blender the reference drives (prepare / feed / blend, stitch_global.cpp:636-666), on the GPU.
class Blender {
public:
    Blender() = default;
    Blender(const Blender&) = delete;
    Blender& operator=(const Blender&) = delete;
    ~Blender() { release(); }
    void release() { if (c_) { ds_destroy_canvas(c_); c_ = nullptr; } }

    // blender->prepare(corners, sizes): dst_roi = resultRoi(corners, sizes)
    void prepare(const Rect& dst_roi, const StitchTuning& tuning, int device = 0, int band_y0 = 0, int band_y1 = 0) {
        release();
        ds_canvas_desc d;
        std::memset(&d, 0, sizeof(d));
        d.x = dst_roi.x; d.y = dst_roi.y; d.width = dst_roi.width; d.height = dst_roi.height;
        d.blend_mode = tuning.feather ? DS_BLEND_FEATHER : DS_BLEND_MULTIBAND;
        d.num_bands = tuning.blend_bands;
        d.sharpness = tuning.feather_sharpness;
        d.out_format = DS_OUT_BGR8;
        d.device = device;
        d.band_y0 = band_y0; d.band_y1 = band_y1;
        check(ds_create_canvas(&d, &c_));
        roi_ = dst_roi;
        fed_ = 0;
    }
    // warper->warp(img) + warp(mask) + convertTo(16S) + blender->feed(img_s, mask, corner). The image must stay
    // valid until blend() has returned (the copy is queued, DS_UPLOAD_ASYNC).
    void feed(const ImageView& img, const ds_transform& t, const ds_frame_opts* opts = nullptr) {
        ds_frame_opts o;
        if (opts) o = *opts; else std::memset(&o, 0, sizeof(o));
        o.flags |= DS_UPLOAD_ASYNC;
        check(ds_upload_frame(c_, fed_++, img.data, img.cols, img.rows, img.step, &t, &o));
    }
    // The global stage's feed loop (stitch_global.cpp:643-660) runs after the strips were warped: gains and seam masks
    // for strip `idx`, fed earlier, without sending its pixels again (ds_update_frame_opts).
    void update(int idx, const ds_frame_opts* opts) { check(ds_update_frame_opts(c_, idx, opts)); }
    // warped_masks[idx] (which = 1, needs DS_MASK_CONTENT) or the mask the blender is fed with (which = 0)
    void frameMask(int idx, int which, Image& out) {
        int32_t pl[4];
        check(ds_debug_get_placement(c_, idx, pl));
        out.create(pl[3], pl[2], 1);
        check(ds_download_frame_mask(c_, idx, which, out.data.data(), out.step()));
    }
    // warped_imgs[idx] / its nearest-warped mask, for the CPU-side exposure and seam steps (stitch_global.cpp:497-630)
    void warped(int idx, Image& img, Image& mask) {
        int32_t pl[4];
        check(ds_debug_get_placement(c_, idx, pl));
        img.create(pl[3], pl[2], 3); mask.create(pl[3], pl[2], 1);
        check(ds_debug_get_warped(c_, idx, img.data.data(), mask.data.data()));
    }
    // blender->blend(result, result_mask); result.convertTo(result, CV_8U)
    void blend(Image& result, Image* result_mask = nullptr) {
        check(ds_composite_async(c_));
        ds_canvas_info info;
        check(ds_get_info(c_, &info));
        const int y0 = info.band_y0, y1 = std::min(info.band_y1, roi_.height);
        result.create(y1 - y0, roi_.width, 3);
        if (result_mask) result_mask->create(y1 - y0, roi_.width, 1);
        check(ds_download_tile(c_, 0, y0, roi_.width, y1 - y0, result.data.data(), result.step(),
                               result_mask ? result_mask->data.data() : nullptr, result_mask ? result_mask->step() : 0));
        check(ds_synchronize(c_));
    }
    // autoCropBlackBorder(pano) (src/stitch_common.cpp:4-27) decided on the canvas in device memory; throws
    // ds::Error(DS_ERR_UNSUPPORTED) when the caller should run the reference's findContours on the panorama instead
    Rect autoCropRect() {
        int32_t r[4];
        check(ds_auto_crop_rect(c_, r));
        return Rect{r[0], r[1], r[2], r[3]};
    }
    // pano(rect).clone(): only the kept rectangle crosses PCIe
    void download(const Rect& r, Image& out) {
        out.create(r.height, r.width, 3);
        check(ds_download_tile(c_, r.x, r.y, r.width, r.height, out.data.data(), out.step(), nullptr, 0));
    }
    ds_canvas* handle() const { return c_; }
    const Rect& roi() const { return roi_; }

private:
    ds_canvas* c_ = nullptr;
    Rect roi_;
    int fed_ = 0;
};

// Replaces `stitcher->composePanorama(output)` (src/stitch_robust.cpp:256) with the reference's component choices
// (:203-213), full-resolution compositing (compositing_resol_mpx = -1): `cameras` and `work_scale` are what
// cv::Stitcher holds after estimateTransform (:251); `warped_image_scale` is its median focal (1 in the affine /
// SCANS pipeline).
inline void composePanorama(const std::vector<ImageView>& images, std::vector<CameraParams> cameras, double work_scale,
                            float warped_image_scale, const StitchTuning& tuning, Image& pano, Image* pano_mask = nullptr,
                            Rect* roi_out = nullptr) {
    if (images.empty() || images.size() != cameras.size()) throw Error(DS_ERR_BAD_ARG, "dronestitch: images / cameras mismatch");
    const double compose_work_aspect = 1.0 / work_scale;              // compose_scale (1) / work_scale
    const float scale = (float)((double)warped_image_scale * compose_work_aspect);
    std::vector<ds_transform> xf(images.size());
    std::vector<Rect> placed(images.size());
    for (size_t i = 0; i < images.size(); ++i) {
        cameras[i].focal *= compose_work_aspect;
        cameras[i].ppx *= compose_work_aspect;
        cameras[i].ppy *= compose_work_aspect;
        xf[i] = planeTransform(cameras[i], scale, tuning.use_affine_warper);
        placed[i] = warpRoi(xf[i], images[i].cols, images[i].rows);
    }
    Blender blender;
    blender.prepare(resultRoi(placed), tuning);
    for (size_t i = 0; i < images.size(); ++i) blender.feed(images[i], xf[i]);
    blender.blend(pano, pano_mask);
    if (roi_out) *roi_out = blender.roi();
}

}  // namespace ds
