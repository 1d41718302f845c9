"""ctypes binding of libdronestitch_cuda (include/dronestitch.h).

The product library is the CUDA build only. If it is missing this module raises — there is no CPU
fallback. `Library(path)` can bind any build of the same C ABI (tests/emu uses that to run the kernel
bodies on the CPU for logic checks); the package itself never does.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DS_LIB_VARIANT (development aid): a differently configured build of the same sources, tools/build_variant.py
CUDA_LIB_PATH = os.path.join(_HERE, "lib", "libdronestitch_cuda%s.so" % (("_" + os.environ["DS_LIB_VARIANT"]) if os.environ.get("DS_LIB_VARIANT") else ""))

DS_OK = 0
DS_ERR_BAD_ARG, DS_ERR_OOM, DS_ERR_CUDA, DS_ERR_P2P_UNAVAILABLE, DS_ERR_STATE, DS_ERR_NO_DEVICE, DS_ERR_UNSUPPORTED = range(1, 8)
DS_BLEND_FEATHER, DS_BLEND_MULTIBAND = 0, 1
DS_OUT_BGR8, DS_OUT_BGRA8 = 0, 1
DS_XF_PLANE_F32, DS_XF_AFFINE_F64, DS_XF_HOMOGRAPHY_F64 = 0, 1, 2
DS_BORDER_CONSTANT, DS_BORDER_REFLECT = 0, 1

EXPORTS = [
    "ds_warp_roi", "ds_frame_touches_band", "ds_create_canvas", "ds_upload_frame", "ds_upload_frame_device",
    "ds_composite", "ds_composite_async", "ds_synchronize", "ds_download_tile", "ds_destroy_canvas",
    "ds_p2p_export", "ds_p2p_connect", "ds_p2p_disconnect", "ds_composite_stage",
    "ds_last_error", "ds_get_info", "ds_version", "ds_debug_get_placement", "ds_debug_get_maps",
    "ds_debug_get_warped", "ds_debug_get_frame_level", "ds_set_profiling", "ds_get_kernel_times",
    "ds_update_frame_opts", "ds_download_frame_mask", "ds_auto_crop_rect", "ds_warp_frame",
    "ds_global_blend_bands", "ds_plan_row_bands",
]


class ds_transform(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("affine_warper", C.c_int32),
        ("K", C.c_float * 9), ("R", C.c_float * 9), ("scale", C.c_float),
        ("border", C.c_int32),
        ("M", C.c_double * 9),
        ("corner_x", C.c_int32), ("corner_y", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
    ]


DS_UPLOAD_ASYNC, DS_MASK_CONTENT, DS_SEAM_NEAREST, DS_MASK_SOFT = 1, 2, 4, 8


class ds_frame_opts(C.Structure):
    _fields_ = [("seam_mask", C.c_void_p), ("seam_mask_stride", C.c_size_t), ("channel_gain", C.POINTER(C.c_float)),
                ("seam_lowres", C.c_void_p), ("seam_lowres_w", C.c_int32), ("seam_lowres_h", C.c_int32),
                ("seam_lowres_stride", C.c_size_t),
                ("compensator_gain", C.POINTER(C.c_double)), ("gain_map", C.c_void_p), ("gain_map_stride", C.c_size_t),
                ("flags", C.c_uint32), ("soft_sigma", C.c_double),
                ("gain_blocks", C.c_void_p), ("gain_blocks_w", C.c_int32), ("gain_blocks_h", C.c_int32), ("gain_blocks_stride", C.c_size_t)]


class ds_canvas_desc(C.Structure):
    _fields_ = [
        ("x", C.c_int32), ("y", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("blend_mode", C.c_int32), ("num_bands", C.c_int32), ("sharpness", C.c_float),
        ("out_format", C.c_int32), ("device", C.c_int32),
        ("band_y0", C.c_int32), ("band_y1", C.c_int32),
        ("stream", C.c_void_p),
        ("pipeline_rows", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


class ds_canvas_info(C.Structure):
    _fields_ = [
        ("padded_width", C.c_int32), ("padded_height", C.c_int32), ("num_bands", C.c_int32), ("num_frames", C.c_int32),
        ("band_y0", C.c_int32), ("band_y1", C.c_int32),
        ("device_bytes", C.c_int64), ("launches_last_composite", C.c_int64),
        ("ms_last_composite", C.c_float), ("algorithmic_bytes", C.c_int64),
        ("h2d_bytes_total", C.c_int64),
        ("reserved", C.c_int32 * 6),
    ]


class ds_kernel_time(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("level", C.c_int32), ("ms", C.c_float), ("algorithmic_bytes", C.c_int64)]


class DroneStitchError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dronestitch error {code}: {msg}")
        self.code = code


class Library:
    """One loaded build of the C ABI."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). libdronestitch_cuda has no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        d = self.dll
        d.ds_last_error.restype = C.c_char_p
        d.ds_version.restype = C.c_char_p
        d.ds_create_canvas.argtypes = [C.POINTER(ds_canvas_desc), C.POINTER(C.c_void_p)]
        d.ds_upload_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t,
                                      C.POINTER(ds_transform), C.POINTER(ds_frame_opts)]
        d.ds_upload_frame_device.argtypes = d.ds_upload_frame.argtypes
        d.ds_composite.argtypes = [C.c_void_p]
        d.ds_composite_async.argtypes = [C.c_void_p]
        d.ds_synchronize.argtypes = [C.c_void_p]
        d.ds_download_tile.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                       C.c_void_p, C.c_size_t]
        d.ds_destroy_canvas.argtypes = [C.c_void_p]
        d.ds_destroy_canvas.restype = None
        d.ds_get_info.argtypes = [C.c_void_p, C.POINTER(ds_canvas_info)]
        d.ds_warp_roi.argtypes = [C.POINTER(ds_transform), C.c_int, C.c_int, C.POINTER(C.c_int32)]
        d.ds_frame_touches_band.argtypes = [C.POINTER(ds_canvas_desc), C.POINTER(C.c_int32)]
        d.ds_set_profiling.argtypes = [C.c_void_p, C.c_int]
        d.ds_get_kernel_times.argtypes = [C.c_void_p, C.POINTER(ds_kernel_time), C.c_int, C.POINTER(C.c_int)]
        d.ds_debug_get_placement.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]
        d.ds_debug_get_maps.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        d.ds_debug_get_warped.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        d.ds_debug_get_frame_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        d.ds_p2p_export.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        d.ds_p2p_connect.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        d.ds_p2p_disconnect.argtypes = [C.c_void_p]
        d.ds_composite_stage.argtypes = [C.c_void_p, C.c_int]
        d.ds_update_frame_opts.argtypes = [C.c_void_p, C.c_int, C.POINTER(ds_frame_opts)]
        d.ds_warp_frame.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.POINTER(ds_transform), C.POINTER(C.c_int32),
                                    C.c_void_p, C.c_void_p]
        d.ds_auto_crop_rect.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        d.ds_download_frame_mask.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        d.ds_global_blend_bands.argtypes = [C.c_int, C.c_int, C.c_int]
        d.ds_plan_row_bands.argtypes = [C.POINTER(ds_canvas_desc), C.POINTER(C.c_int32), C.c_int, C.c_int, C.POINTER(C.c_int32)]

    def check(self, rc):
        if rc != DS_OK:
            raise DroneStitchError(rc, (self.dll.ds_last_error() or b"").decode("utf-8", "replace"))

    def version(self):
        return self.dll.ds_version().decode()


_default = None


def default_library():
    """The CUDA product library. Raises if it has not been built."""
    global _default
    if _default is None:
        _default = Library(CUDA_LIB_PATH)
    return _default
