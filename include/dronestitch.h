/*
 * dronestitch.h — C ABI of libdronestitch_cuda (B200 / sm_100a compositing path).
 *
 * Replaces, as a drop-in for that path only, the warp + blend loop of
 * Akika404/drone_image_stitch_cpp (all file:line below are in /root/reference):
 *   - src/stitch_robust.cpp:256        stitcher->composePanorama(output)      (HOT PATH #1)
 *       per frame: AffineWarper::warp (LINEAR/REFLECT), mask warp (NEAREST/CONSTANT),
 *       ->16S, MultiBandBlender::feed; then blend, ->8U   (components chosen at :203-213)
 *   - src/stitch_global.cpp:470-486    cv::warpAffine loop                    (HOT PATH #2, coords)
 *   - src/stitch_global.cpp:632-666    MultiBandBlender prepare / feed / blend / ->8U
 * The reference has no plugin / FFI layer of its own (direct C++ calls into OpenCV); the four entry
 * points create_canvas / upload_frame / composite / download_tile are the boundary its north-star
 * names, with argument conventions taken from the reference's data (8UC3 BGR cv::Mat + step,
 * cameras[i].K / R float32 3x3, cv::Rect ROI, cv::Stitcher::Status-like int codes).
 *
 * Plain C: pointers, sizes and POD structs only. No exceptions cross this boundary.
 * A ds_canvas is not thread-safe; distinct canvases may be used from distinct threads.
 * There is no CPU fallback: every entry point that computes fails with DS_ERR_NO_DEVICE
 * when no CUDA device is usable.
 */
#ifndef DRONESTITCH_H_
#define DRONESTITCH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define DS_API __declspec(dllexport)
#else
#define DS_API __attribute__((visibility("default")))
#endif

#define DS_MAX_LEVELS 13 /* pyramid levels 0..12 (reference caps bands at 12, stitch_global.cpp:632-635) */

typedef struct ds_canvas ds_canvas; /* opaque; owns all device memory */

/* Status codes (reference: cv::Stitcher::Status ints, src/stitch_common.cpp:29-42). 0 = OK. */
typedef enum ds_status {
    DS_OK = 0,
    DS_ERR_BAD_ARG = 1,
    DS_ERR_OOM = 2,
    DS_ERR_CUDA = 3,
    DS_ERR_P2P_UNAVAILABLE = 4,
    DS_ERR_STATE = 5,
    DS_ERR_NO_DEVICE = 6,
    DS_ERR_UNSUPPORTED = 7
} ds_status;

typedef enum ds_blend_mode {
    DS_BLEND_FEATHER = 0,   /* cv::detail::FeatherBlender(sharpness)            (BASELINE cfg 1) */
    DS_BLEND_MULTIBAND = 1  /* cv::detail::MultiBandBlender(false, bands, 32F)  (stitch_robust.cpp:213) */
} ds_blend_mode;

typedef enum ds_out_format {
    DS_OUT_BGR8 = 0, /* 3 B/px, matches result.convertTo(CV_8U); mask is a separate plane */
    DS_OUT_BGRA8 = 1 /* 4 B/px, alpha = result mask (255 / 0) */
} ds_out_format;

typedef enum ds_xf_kind {
    DS_XF_PLANE_F32 = 0,     /* cv::detail::PlaneWarper / cv::AffineWarper (K, R float32, scale) */
    DS_XF_AFFINE_F64 = 1,    /* cv::warpAffine semantics, forward 2x3 double (stitch_global.cpp:474-480) */
    DS_XF_HOMOGRAPHY_F64 = 2 /* cv::warpPerspective semantics, forward 3x3 double */
} ds_xf_kind;

typedef enum ds_border { DS_BORDER_CONSTANT = 0, DS_BORDER_REFLECT = 1 } ds_border;

/* Per-frame transform. Row-major matrices.
 * PLANE_F32: exactly what cv::Stitcher holds per camera after estimateTransform (:251) and the
 *   compose-scale update: K (float32 3x3), R (float32 3x3; with affine_warper != 0 it is the 3x3
 *   affine "H" of cv::AffineWarper), scale = warped_image_scale * compose_work_aspect.
 *   Placement (corner, size) is computed by the library like warper.warpRoi().
 * AFFINE_F64 / HOMOGRAPHY_F64: M maps source pixels to pixels of the frame's own warped bbox whose
 *   top-left is `corner` in canvas coordinates and whose size is `size` (the reference subtracts the
 *   corner from the translation, stitch_global.cpp:474-476). */
typedef struct ds_transform {
    int32_t kind;          /* ds_xf_kind */
    int32_t affine_warper; /* PLANE_F32 only: 1 = cv::AffineWarper, 0 = cv::PlaneWarper (T = 0) */
    float K[9];
    float R[9];
    float scale;
    int32_t border;        /* ds_border for the image taps; PLANE default REFLECT, others CONSTANT */
    double M[9];           /* AFFINE_F64: M[0..5]; HOMOGRAPHY_F64: M[0..8] */
    int32_t corner_x, corner_y; /* AFFINE / HOMOGRAPHY only */
    int32_t width, height;      /* AFFINE / HOMOGRAPHY only */
} ds_transform;

/* Optional per-frame inputs. Pass NULL for none. */
typedef struct ds_frame_opts {
    const uint8_t* seam_mask; /* 8UC1, size of the frame's warped bbox, ANDed into the warped mask */
    size_t seam_mask_stride;
    const float* channel_gain; /* 3 floats (B, G, R) applied as sat_u8(float(p) * g), stitch_global.cpp:291-305 */
    /* Low-resolution seam mask as cv::Stitcher holds it after seam finding (masks_warped[i], 8UC1).
     * The library does what composePanorama does with it per frame: dilate 3x3, resize to the warped bbox
     * with INTER_LINEAR_EXACT, AND into the warped mask. Overrides seam_mask when non-NULL. */
    const uint8_t* seam_lowres;
    int32_t seam_lowres_w, seam_lowres_h;
    size_t seam_lowres_stride;
    /* ExposureCompensator::apply with scalar gains (GainCompensator / ChannelsCompensator,
     * stitch_global.cpp:644): 3 doubles (B, G, R), sat_u8(rint(double(p) * g)); applied after channel_gain. */
    const double* compensator_gain;
    /* BlocksGainCompensator::apply (stitch_robust.cpp:209-211): float32 gain per pixel of the warped bbox
     * (the caller's resize of its block gain map), sat_u8(rint(float(p) * g)); applied last. */
    const float* gain_map;
    size_t gain_map_stride; /* bytes */
    /* DS_UPLOAD_ASYNC: return as soon as the copies are queued. The caller keeps `bgr` and every buffer named
     * here valid and unchanged until ds_composite / ds_synchronize / a ds_download_tile covering the canvas has
     * returned; host buffers should be pinned (cudaHostAlloc / cudaHostRegister), or the copy is synchronous
     * anyway. With frames still in flight, ds_composite_async works through the canvas in row slices, each
     * starting as soon as the frames it reads have arrived, and ds_download_tile copies every slice out as soon
     * as it is final - upload, compute and download overlap. Results are identical. */
    uint32_t flags;
    /* DS_MASK_SOFT: sigma of buildSoftBlendMask's GaussianBlur; 0 = the reference's 10.0 (stitch_global.cpp:345).
     * Supported up to 10 (kernel of 81 taps). */
    double soft_sigma;   /* double like cv::GaussianBlur's sigmaX: the kernel taps are computed from it in double */
    /* BlocksGainCompensator::apply as the reference configures it (stitch_robust.cpp:209-211), from the compensator's own
     * block gain map (gain_maps_[i], CV_32F, one value per 32x32 block of the seam-scale frame): the library resizes it to
     * the warped bbox on the device exactly like cv::resize(INTER_LINEAR) does for CV_32FC1 - fractions in double,
     * fused multiply-adds - and applies sat_u8(rint(float(p) * g)) last. Overrides gain_map when non-NULL. Block maps with
     * a single row or column (frames under 33 px) are refused with DS_ERR_UNSUPPORTED (cv::resize treats them differently). */
    const float* gain_blocks;
    int32_t gain_blocks_w, gain_blocks_h;
    size_t gain_blocks_stride; /* bytes; 0 = dense */
} ds_frame_opts;

/* ds_frame_opts.flags */
enum {
    DS_UPLOAD_ASYNC = 1u,
    /* The global stage's masks (stitchInterStripsCustom), built on the device:
     * DS_MASK_CONTENT  buildWarpedContentMask (stitch_global.cpp:353-383): BGR2GRAY > 3 of the source as float 0/1,
     *                  warped with the frame's own transform (INTER_LINEAR, BORDER_CONSTANT 0), > 0.999 -> 255. It is
     *                  ANDed into the frame's mask and can be read back with ds_download_frame_mask(.., 1, ..) - the
     *                  reference hands it to the exposure compensator and the seam finder (:577, :599-607).
     * DS_SEAM_NEAREST  seam_lowres is brought to the warped bbox by resize(INTER_NEAREST) + threshold(> 1)
     *                  (:649-655) instead of composePanorama's dilate + INTER_LINEAR_EXACT.
     * DS_MASK_SOFT     buildSoftBlendMask (:332-351): (seam AND content) > 1 as float 0/1 -> GaussianBlur(sigma,
     *                  BORDER_REPLICATE) -> * binary -> * 255 -> 8U is the mask the blender is fed with (:657-658).
     * A frame uploaded with DS_MASK_CONTENT is copied at once even with DS_UPLOAD_ASYNC (the mask reads its pixels). */
    DS_MASK_CONTENT = 2u,
    DS_SEAM_NEAREST = 4u,
    DS_MASK_SOFT = 8u
};

typedef struct ds_canvas_desc {
    int32_t x, y, width, height; /* canvas ROI = cv::detail::resultRoi(corners, sizes) */
    int32_t blend_mode;          /* ds_blend_mode */
    int32_t num_bands;           /* MULTIBAND: requested bands (cropped like MultiBandBlender::prepare) */
    float sharpness;             /* FEATHER: 0.02f is OpenCV's default */
    int32_t out_format;          /* ds_out_format */
    int32_t device;              /* CUDA device ordinal */
    /* Row band of the 2^bands-padded canvas this handle computes: rows [band_y0, band_y1) relative
     * to the canvas ROI origin. band_y1 <= 0 means the whole canvas. Band edges must be multiples of
     * 2^bands (or the canvas end). Frames that do not touch the band (+ its pyramid halo) may be
     * skipped by the caller; ds_frame_touches_band() tells. */
    int32_t band_y0, band_y1;
    void* stream;                /* cudaStream_t to run on, or NULL for a library-owned stream */
    /* Row slices of ds_composite: 0 = automatic (slices of a default height only while DS_UPLOAD_ASYNC uploads
     * are in flight), > 0 = always slices of about this many rows, < 0 = never. Output does not depend on it. */
    int32_t pipeline_rows;
    int32_t reserved[7];
} ds_canvas_desc;

typedef struct ds_canvas_info {
    int32_t padded_width, padded_height; /* canvas padded to a multiple of 2^bands */
    int32_t num_bands;                   /* effective bands */
    int32_t num_frames;
    int32_t band_y0, band_y1;
    int64_t device_bytes;                /* library-owned device memory */
    int64_t launches_last_composite;     /* kernels launched by the last ds_composite */
    float ms_last_composite;             /* device time of the last ds_composite (CUDA events) */
    int64_t algorithmic_bytes;           /* SURVEY.md §8(d) AB model for the uploaded frames */
    int64_t h2d_bytes_total;             /* frame bytes copied host -> device by this handle so far (a row-band
                                          * handle fed with DS_UPLOAD_ASYNC transfers only the source rows it reads) */
    int32_t reserved[6];
} ds_canvas_info;

/* ---- geometry helpers (host only, no device needed) ---- */

/* warper.warpRoi(src_size, K, R): out_xywh = {corner.x, corner.y, width, height}. PLANE_F32 only;
 * for the other kinds it echoes corner/size from the transform. */
DS_API int ds_warp_roi(const ds_transform* xf, int src_w, int src_h, int32_t out_xywh[4]);

/* 1 if a frame placed at out_xywh (from ds_warp_roi) contributes to the band [band_y0, band_y1) of a
 * canvas described by `desc` (footprint + multi-band gap + pyramid halo), else 0. */
DS_API int ds_frame_touches_band(const ds_canvas_desc* desc, const int32_t frame_xywh[4]);

/* Band count of the global stage's blender (src/stitch_global.cpp:632-635):
 *   auto = min(12, ceil(log2(max(canvas_w, canvas_h))) - 1);  bands = max(max(5, configured), auto).
 * (ds_create_canvas then crops it like MultiBandBlender::prepare does.) Returns the band count, or -1 for an empty canvas. */
DS_API int ds_global_blend_bands(int canvas_w, int canvas_h, int configured_bands);

/* Row-band edges for `n_bands` handles of one canvas (one per GPU): edges_out[0] = 0 <= ... <= edges_out[n_bands] =
 * padded height, interior edges at multiples of 2^bands (feather: of 32), chosen so that every band holds about the
 * same share of the work - the warped-bbox pixels of the frames (frames_xywh = n_frames x {x, y, w, h} from
 * ds_warp_roi, absolute coordinates) plus the canvas pixels themselves - instead of the same number of rows: flight
 * lines overlap, so rows under two lines cost twice the rows under one. */
DS_API int ds_plan_row_bands(const ds_canvas_desc* desc, const int32_t* frames_xywh, int n_frames, int n_bands, int32_t* edges_out);

/* ---- the four entry points ---- */

/* Replaces blender->prepare(corners, sizes) (stitch_global.cpp:636-638; inside composePanorama). */
DS_API int ds_create_canvas(const ds_canvas_desc* desc, ds_canvas** out);

/* Replaces, per frame, warper->warp(img) + warper->warp(mask) + convertTo(16S) + blender->feed()'s
 * input hand-off. bgr: 8UC3 interleaved, `stride` = cv::Mat::step. The host buffer is only borrowed
 * for the duration of the call (unless opts->flags has DS_UPLOAD_ASYNC). Feed order = frame_idx order.
 * Re-uploading an index replaces it. */
DS_API int ds_upload_frame(ds_canvas* c, int frame_idx, const uint8_t* bgr, int w, int h, size_t stride,
                           const ds_transform* xf, const ds_frame_opts* opts);

/* New optional inputs (seam masks, gains, DS_MASK_* flags) for a frame whose pixels are already uploaded: what the
 * global stage does between its warp loop (stitch_global.cpp:470-486) and its feed loop (:643-660), where exposure
 * gains and seam masks only exist after the strips were warped. No pixels are transferred; DS_UPLOAD_ASYNC is
 * ignored. opts == NULL removes every optional input. */
DS_API int ds_update_frame_opts(ds_canvas* c, int frame_idx, const ds_frame_opts* opts);

/* A frame's 8UC1 mask plane over its warped bbox (size from ds_debug_get_placement / ds_warp_roi):
 * which = 0: the mask the blender is fed with (nearest-warped 255s AND seam / content / soft mask),
 * which = 1: the content mask of DS_MASK_CONTENT (warped_masks[i] of the global stage). */
DS_API int ds_download_frame_mask(ds_canvas* c, int frame_idx, int which, uint8_t* out, size_t stride);

/* Same, but `dev_bgr` is a device pointer (frames already resident in HBM). The library reads it on its own streams:
 * whatever produced the pixels must have completed (synchronise the producing stream first). */
DS_API int ds_upload_frame_device(ds_canvas* c, int frame_idx, const void* dev_bgr, int w, int h, size_t stride,
                                  const ds_transform* xf, const ds_frame_opts* opts);

/* Warp + mask/weights + blend (+ collapse) over all uploaded frames. Synchronous on return. */
DS_API int ds_composite(ds_canvas* c);
/* Enqueue only (no host sync); results are valid after the canvas stream is synchronised. */
DS_API int ds_composite_async(ds_canvas* c);
DS_API int ds_synchronize(ds_canvas* c);

/* Copy a tile of the composited canvas to host memory. (x, y) relative to the canvas ROI origin.
 * out: BGR8 (3 B/px) or BGRA8 (4 B/px) per the canvas out_format. mask_out may be NULL. */
DS_API int ds_download_tile(ds_canvas* c, int x, int y, int w, int h, uint8_t* out, size_t stride,
                            uint8_t* mask_out, size_t mask_stride);

/* One frame warped on its own, no canvas: warper->warp(img, K, R, INTER_LINEAR, BORDER_REFLECT, img_warped) and
 * warper->warp(mask, K, R, INTER_NEAREST, BORDER_CONSTANT, mask_warped) as composePanorama runs them in its seam phase
 * on the seam-scale images (cameras scaled by seam_work_aspect; called from src/stitch_robust.cpp:256), or
 * cv::warpAffine of a strip (src/stitch_global.cpp:479-480) - whatever `xf` describes. The results feed the CPU-side
 * exposure compensator and seam finder. out_xywh = placement (ds_warp_roi); out_bgr (3 * w * h bytes, dense) and
 * out_mask (w * h bytes) may be NULL to query the placement only. */
DS_API int ds_warp_frame(int device, const uint8_t* bgr, int w, int h, size_t stride, const ds_transform* xf,
                         int32_t out_xywh[4], uint8_t* out_bgr, uint8_t* out_mask);

/* autoCropBlackBorder(pano) (src/stitch_common.cpp:4-27; called at src/stitch_app.cpp:213, :262) on the composited
 * canvas, without moving it to the host: BGR2GRAY > 1, the external contour of largest cv::contourArea, its
 * boundingRect -> out_xywh (relative to the canvas ROI origin; the whole canvas when nothing is brighter than 1, as the
 * reference returns the panorama untouched). The caller then downloads just that rectangle with ds_download_tile.
 * The device reduces the canvas to its per-row foreground runs; the host labels their 8-connected components and
 * picks the winner by bounds on contourArea that need no contour tracing: core pixels - 1 <= area <= (w-1)(h-1).
 * When those bounds cannot separate the two best candidates (or a row has more runs than the list holds) the call
 * returns DS_ERR_UNSUPPORTED and the caller keeps the reference's cv::findContours on the downloaded panorama.
 * Whole-canvas handles only (no row band). */
DS_API int ds_auto_crop_rect(ds_canvas* c, int32_t out_xywh[4]);

/* ---- NVLink peer-to-peer halo exchange between row-band handles (one handle per GPU, any process layout) ----
 * Without it a band recomputes the pyramid halo beyond its edges from the frames themselves: no communication,
 * about 9 % extra work per band at 5 bands. With it the level-0 feed (warp + first pyrDown, the bulk of the work)
 * runs over exactly the band's own rows, and the ~140 rows of every straddling frame's level-1 Gaussian / weight
 * level needed beyond each edge are pulled from the neighbouring handle's memory by one kernel of NVLink P2P loads
 * per composite; the levels above are rebuilt from them locally. No collective, no host round trip: neighbours
 * order themselves with a counter word in each other's memory (a one-thread kernel writes it, the reader's stream
 * waits for it with a stream memory operation).
 * Usage: upload every frame that touches the band (+ halo; ds_frame_touches_band), ds_p2p_export on every handle,
 * move the blobs between the processes by any means, ds_p2p_connect the neighbour above (side 0) and below
 * (side 1), then composite in lock step - every connected handle the same number of composites. Export again
 * and reconnect after frames were added or changed geometry. The exchange is used for composites over resident
 * frames; the pipelined schedule of DS_UPLOAD_ASYNC keeps the recomputed halo (its slices cannot wait for a
 * neighbour's last slice). Output is identical in all modes. */
DS_API int ds_p2p_export(ds_canvas* c, void* blob, size_t capacity, size_t* size);
DS_API int ds_p2p_connect(ds_canvas* c, int side, const void* blob, size_t size);
DS_API int ds_p2p_disconnect(ds_canvas* c);
/* ds_composite_async in two halves, for one thread driving several connected handles: stage 0 queues everything
 * up to the hand-over of this handle's level-1 rows, stage 1 the rest. Call stage 0 on every handle, then stage 1. */
DS_API int ds_composite_stage(ds_canvas* c, int stage);

DS_API void ds_destroy_canvas(ds_canvas* c);
DS_API const char* ds_last_error(void);
DS_API int ds_get_info(const ds_canvas* c, ds_canvas_info* info);
DS_API const char* ds_version(void);

/* ---- per-kernel timing (measurement only) ---- */

typedef struct ds_kernel_time {
    char name[32];             /* "mb_feed" (level 0), "mb_pyrdown", "mb_accum", "mb_collapse", "p2p_pull", "feather_mask", "feather_dist", "feather_blend" */
    int32_t level;             /* pyramid level the launch works on (-1 if not applicable) */
    float ms;                  /* device time of the launch (CUDA events on the canvas stream) */
    int64_t algorithmic_bytes; /* share of the SURVEY.md §8(d) model attributed to this launch */
} ds_kernel_time;

/* With profiling on, ds_composite records a CUDA event pair around every kernel it launches; entries
 * accumulate over successive composites until ds_set_profiling is called again. */
DS_API int ds_set_profiling(ds_canvas* c, int on);
/* After a profiled ds_composite: fills up to `cap` entries, *n = number of launches. */
DS_API int ds_get_kernel_times(ds_canvas* c, ds_kernel_time* out, int cap, int* n);

/* ---- debug taps for the parity tests ---- */

/* Placement of an uploaded frame: {corner.x, corner.y, width, height} (absolute canvas coords). */
DS_API int ds_debug_get_placement(ds_canvas* c, int frame_idx, int32_t out_xywh[4]);
/* The INTER_BITS fixed-point remap tables of a frame's warped bbox: xy = int16 pairs (w*h*2),
 * a = uint16 (w*h) — the same layout as cv::convertMaps(..., CV_16SC2). */
DS_API int ds_debug_get_maps(ds_canvas* c, int frame_idx, int16_t* xy, uint16_t* a);
/* The warped 8UC3 image and warped 8UC1 mask of a frame's bbox (dense, w*h*3 and w*h). */
DS_API int ds_debug_get_warped(ds_canvas* c, int frame_idx, uint8_t* bgr, uint8_t* mask);
/* MULTIBAND, after ds_composite: a frame's Gaussian level l >= 1 (16SC3 dense) and weight level
 * (f32 dense) over its aligned feed ROI; dims_out = {roi_x, roi_y, width, height} at level l
 * relative to the padded canvas origin. Either pointer may be NULL (query dims only). */
DS_API int ds_debug_get_frame_level(ds_canvas* c, int frame_idx, int level, int16_t* g, float* w,
                                    int32_t dims_out[4]);

#ifdef __cplusplus
}
#endif
#endif /* DRONESTITCH_H_ */
