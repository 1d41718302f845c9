// C++ host-side parity driver: runs ds::composePanorama (include/dronestitch.hpp) on a case file written by
// tests/test_cpp_host.py and writes the panorama back; the Python test compares it with the oracle. Also checks the
// error behaviour of the wrapper (exceptions instead of status codes, as the reference's call sites expect).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "dronestitch.hpp"

template <class T>
static T rd(std::ifstream& f) { T v; f.read(reinterpret_cast<char*>(&v), sizeof(T)); return v; }

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: host_compose <case.bin> <out.bin>\n"); return 2; }
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    char magic[4];
    f.read(magic, 4);
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
