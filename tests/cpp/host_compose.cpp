// C++ host-side parity driver: runs ds::composePanorama (include/dronestitch.hpp) on a case file written by
// tests/test_cpp_host.py and writes the panorama back; the Python test compares it with the oracle. Also checks the
// error behaviour of the wrapper (exceptions instead of status codes, as the reference's call sites expect).
#include <array>
#include <climits>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "dronestitch.hpp"

template <class T>
static T rd(std::ifstream& f) { T v; f.read(reinterpret_cast<char*>(&v), sizeof(T)); return v; }

// The compose half of stitchInterStripsCustom (src/stitch_global.cpp:439-486, :632-666) through ds::Blender: canvas
// from transformedBoundingRect, warpAffine placement, content masks at feed time, seam masks + gains afterwards
// (ds_update_frame_opts), soft blend masks, multi-band blend.
static int global_stage(std::ifstream& f, const char* out_path) {
    const int n = rd<int32_t>(f), bands = rd<int32_t>(f);
    std::vector<std::vector<uint8_t>> pixels(n), seams(n);
    std::vector<ds::ImageView> strips(n);
    std::vector<std::array<double, 9>> H(n);
    std::vector<int> sw(n), sh(n);
    std::vector<std::array<float, 3>> gain(n);
    for (int i = 0; i < n; i++) {
        const int cols = rd<int32_t>(f), rows = rd<int32_t>(f);
        f.read(reinterpret_cast<char*>(H[i].data()), 9 * sizeof(double));
        f.read(reinterpret_cast<char*>(gain[i].data()), 3 * sizeof(float));
        pixels[i].resize((size_t)cols * rows * 3);
        f.read(reinterpret_cast<char*>(pixels[i].data()), (std::streamsize)pixels[i].size());
        strips[i] = ds::ImageView{pixels[i].data(), cols, rows, (size_t)cols * 3};
        sw[i] = rd<int32_t>(f); sh[i] = rd<int32_t>(f);
        seams[i].resize((size_t)sw[i] * sh[i]);
        f.read(reinterpret_cast<char*>(seams[i].data()), (std::streamsize)seams[i].size());
    }
    try {
        // :439-468 canvas, shift, corners, sizes
        int min_x = INT_MAX, min_y = INT_MAX, max_x = INT_MIN, max_y = INT_MIN;
        for (int i = 0; i < n; i++) {
            const ds::Rect r = ds::transformedBoundingRect(strips[i].cols, strips[i].rows, H[i].data());
            min_x = std::min(min_x, r.x); min_y = std::min(min_y, r.y);
            max_x = std::max(max_x, r.x + r.width); max_y = std::max(max_y, r.y + r.height);
        }
        std::vector<ds::Rect> placed(n);
        std::vector<ds_transform> xf(n);
        for (int i = 0; i < n; i++) {
            std::array<double, 9> S = H[i];
            S[2] += (double)-min_x; S[5] += (double)-min_y;   // shift * global_transforms[i]
            placed[i] = ds::transformedBoundingRect(strips[i].cols, strips[i].rows, S.data());
            double M[6] = {S[0], S[1], S[2] - (double)placed[i].x, S[3], S[4], S[5] - (double)placed[i].y};
            xf[i] = ds::affineTransform(M, placed[i].x, placed[i].y, placed[i].width, placed[i].height);
        }
        ds::StitchTuning tuning;
        // :632-635 band count from the canvas size and the configured value (canvas_w / canvas_h of :455-456)
        tuning.blend_bands = ds::globalBlendBands(max_x - min_x, max_y - min_y, bands);
        if (tuning.blend_bands != ds_global_blend_bands(max_x - min_x, max_y - min_y, bands)) throw std::runtime_error("band rule: header and library disagree");
        ds::Blender blender;
        blender.prepare(ds::resultRoi(placed), tuning);
        ds_frame_opts o;
        std::memset(&o, 0, sizeof(o));
        o.flags = DS_MASK_CONTENT;
        for (int i = 0; i < n; i++) blender.feed(strips[i], xf[i], &o);   // :470-486 warp + content mask
        std::ofstream out(out_path, std::ios::binary);
        for (int i = 0; i < n; i++) {
            ds::Image cm;
            blender.frameMask(i, 1, cm);                                   // warped_masks[i] for the CPU-side steps
            out.write(reinterpret_cast<const char*>(cm.data.data()), (std::streamsize)cm.data.size());
        }
        for (int i = 0; i < n; i++) {                                      // :643-660
            std::memset(&o, 0, sizeof(o));
            o.flags = DS_MASK_CONTENT | DS_SEAM_NEAREST | DS_MASK_SOFT;
            o.seam_lowres = seams[i].data(); o.seam_lowres_w = sw[i]; o.seam_lowres_h = sh[i];
            o.channel_gain = gain[i].data();
            blender.update(i, &o);
        }
        ds::Image pano, mask;
        blender.blend(pano, &mask);
        const ds::Rect roi = blender.roi();
        const int32_t hdr[4] = {roi.x, roi.y, roi.width, roi.height};
        out.write(reinterpret_cast<const char*>(hdr), sizeof(hdr));
        out.write(reinterpret_cast<const char*>(pano.data.data()), (std::streamsize)pano.data.size());
        out.write(reinterpret_cast<const char*>(mask.data.data()), (std::streamsize)mask.data.size());
        // autoCropBlackBorder (stitch_app.cpp:262): the rectangle decided on the device, then only that tile downloaded
        int32_t keep[5] = {0, 0, roi.width, roi.height, 0};
        try {
            const ds::Rect r = blender.autoCropRect();
            keep[0] = r.x; keep[1] = r.y; keep[2] = r.width; keep[3] = r.height; keep[4] = 1;
        } catch (const ds::Error& e) {
            if (e.code != DS_ERR_UNSUPPORTED) throw;
        }
        out.write(reinterpret_cast<const char*>(keep), sizeof(keep));
        ds::Image cropped;
        blender.download(ds::Rect{keep[0], keep[1], keep[2], keep[3]}, cropped);
        out.write(reinterpret_cast<const char*>(cropped.data.data()), (std::streamsize)cropped.data.size());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    std::puts("ok");
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: host_compose <case.bin> <out.bin>\n"); return 2; }
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    char magic[4];
    f.read(magic, 4);
    if (std::memcmp(magic, "DSG1", 4) == 0) return global_stage(f, argv[2]);
    const int n = rd<int32_t>(f), bands = rd<int32_t>(f), feather = rd<int32_t>(f), affine = rd<int32_t>(f);
    const float warped_image_scale = rd<float>(f);
    const double work_scale = rd<double>(f);
    std::vector<std::vector<uint8_t>> pixels(n);
    std::vector<ds::ImageView> images(n);
    std::vector<ds::CameraParams> cameras(n);
    for (int i = 0; i < n; i++) {
        const int cols = rd<int32_t>(f), rows = rd<int32_t>(f);
        cameras[i].focal = rd<double>(f); cameras[i].aspect = rd<double>(f);
        cameras[i].ppx = rd<double>(f); cameras[i].ppy = rd<double>(f);
        f.read(reinterpret_cast<char*>(cameras[i].R.data()), 9 * sizeof(float));
        pixels[i].resize((size_t)cols * rows * 3);
        f.read(reinterpret_cast<char*>(pixels[i].data()), (std::streamsize)pixels[i].size());
        images[i] = ds::ImageView{pixels[i].data(), cols, rows, (size_t)cols * 3};
    }
    ds::StitchTuning tuning;
    tuning.blend_bands = bands;
    tuning.feather = feather != 0;
    tuning.use_affine_warper = affine != 0;
    try {
        // error behaviour first: exceptions carrying the C status code
        bool threw = false;
        try {
            std::vector<ds::CameraParams> fewer(cameras.begin(), cameras.end() - (n > 1 ? 1 : 0));
            if (n > 1) { ds::Image p; ds::composePanorama(images, fewer, work_scale, warped_image_scale, tuning, p); }
            else threw = true;
        } catch (const ds::Error& e) { threw = e.code == DS_ERR_BAD_ARG; }
        if (!threw) { std::fprintf(stderr, "mismatched inputs were accepted\n"); return 3; }
        threw = false;
        try {
            ds::Blender b;
            b.prepare(ds::Rect{0, 0, 64, 64}, tuning);
            ds_transform t = ds::planeTransform(cameras[0], warped_image_scale, tuning.use_affine_warper);
            b.feed(images[0], t);   // the frame's bbox leaves a 64 x 64 canvas
        } catch (const ds::Error& e) { threw = e.code == DS_ERR_BAD_ARG; }
        if (!threw) { std::fprintf(stderr, "a frame outside the canvas was accepted\n"); return 3; }

        ds::Image pano, mask;
        ds::Rect roi;
        ds::composePanorama(images, cameras, work_scale, warped_image_scale, tuning, pano, &mask, &roi);
        std::ofstream o(argv[2], std::ios::binary);
        const int32_t hdr[4] = {roi.x, roi.y, roi.width, roi.height};
        o.write(reinterpret_cast<const char*>(hdr), sizeof(hdr));
        o.write(reinterpret_cast<const char*>(pano.data.data()), (std::streamsize)pano.data.size());
        o.write(reinterpret_cast<const char*>(mask.data.data()), (std::streamsize)mask.data.size());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    std::puts("ok");
    return 0;
}
