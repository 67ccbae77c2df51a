"""Host-side mirror of the reference's `Tracer` (include/tracer.hpp:26-88, src/tracer.cpp) over the
C ABI of libsrt_b200.so.  Same method names, argument meaning and frame protocol as the reference:

    tracer = Tracer(width, height, skybox)          # Tracer::Tracer, tracer.cpp:11-68
    tracer.options / tracer.scene_data              # public mutable RenderData / SceneData
    tracer.clear_canvas()                           # tracer.cpp:98-101
    tracer.update_scene(shapes, triangles, mats)    # tracer.cpp:70-96
    tracer.render(ticks_stopped, output)            # tracer.cpp:103-116 (render + average + read-back)

There is no CPU path: importing works anywhere, but constructing a Tracer without the CUDA library
or without a GPU raises.
"""
import ctypes
import os

import numpy as np

from .records import COUNTERS, MATERIAL, RENDER_DATA, SCENE_DATA, SHAPE, TRIANGLE, concat_records

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SRT_LIB") or os.path.join(_HERE, "libsrt_b200.so")  # SRT_LIB: developer builds
_lib = None


class SrtError(RuntimeError):
    """A C-ABI call returned non-zero (the reference throws boost::compute::opencl_error)."""


def load_library():
    """dlopen libsrt_b200.so and declare every entry point of include/srt.h.  Raises if the library
    has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SrtError(f"{LIB_PATH} is missing: the CUDA extension is not built and there is no fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, u32, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_size_t
    pp = ctypes.POINTER(ctypes.c_void_p)
    sigs = {
        "srt_create": [i32, i32, vp, i32, i32, i32, pp],
        "srt_upload_scene": [vp, vp, sz, vp, sz, vp, sz, vp],
        "srt_clear": [vp],
        "srt_render": [vp, vp],
        "srt_render_batch": [vp, vp, sz],
        "srt_reserve_batch": [vp, vp, sz],
        "srt_resolve": [vp, u32, vp],
        "srt_pin_output": [vp, vp, sz],
        "srt_unpin_output": [vp],
        "srt_render_frame": [vp, vp, u32, vp],
        "srt_set_row_bands": [vp, i32, i32, i32],
        "srt_set_accel": [vp, i32],
        "srt_set_sweep_filter": [vp, i32],
        "srt_set_schedule": [vp, i32],
        "srt_set_frame_pipeline": [vp, i32],
        "srt_read_canvas": [vp, vp],
        "srt_write_canvas": [vp, vp],
        "srt_canvas_device_ptr": [vp, pp, ctypes.POINTER(sz)],
        "srt_output_device_ptr": [vp, pp, ctypes.POINTER(sz)],
        "srt_resolve_device": [vp, u32],
        "srt_resolve_device_range": [vp, u32, sz, sz],
        "srt_read_output": [vp, vp],
        "srt_stream": [vp, pp],
        "srt_synchronize": [vp],
        "srt_debug_primary": [vp, vp, vp, vp],
        "srt_render_counted": [vp, vp, vp],
        "srt_debug_math": [vp, i32, vp, vp, vp, sz],
        "srt_measure_fp32_peak": [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)],
        "srt_render_time_ms": [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)],
        "srt_destroy": [vp],
        "srt_abi_version": [],
        "srt_load_stl": [ctypes.c_char_p, pp, ctypes.POINTER(sz)],
        "srt_load_obj": [ctypes.c_char_p, pp, ctypes.POINTER(sz)],
        "srt_save_ppm": [ctypes.c_char_p, vp, i32, i32],
        "srt_model_bounds": [vp, sz, vp],
        "srt_load_skybox_png": [ctypes.c_char_p, pp, ctypes.POINTER(i32), ctypes.POINTER(i32)],
    }
    for name, args in sigs.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = i32
    L.srt_last_error.argtypes = [vp]
    L.srt_last_error.restype = ctypes.c_char_p
    L.srt_free.argtypes = [vp]
    L.srt_free.restype = None
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class _CudaView:
    """__cuda_array_interface__ view of a device buffer owned by the tracer (for torch.as_tensor)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False),
                                         "version": 3, "strides": None}


class Tracer:
    MATH_OPS = {"log": 0, "cos": 1, "atan2pi": 2, "pow": 3, "sqrt": 4, "schlick": 5,
                "log_x2_lo": 6, "log_x2_hi": 7, "cos_x2_lo": 8, "cos_x2_hi": 9,
                "rcp_sqrt": 10, "rcp_of_sqrt": 11, "sqrt_x2_lo": 12, "sqrt_x2_hi": 13,
                "rcp_sqrt_all_patterns": 14, "sqrt_x2_all_patterns": 15}

    def __init__(self, width, height, skybox, device=-1):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        self.width, self.height = int(width), int(height)
        sky = np.ascontiguousarray(skybox, np.float32)
        if sky.ndim != 3 or sky.shape[2] != 4:
            raise ValueError("skybox must be (h, w, 4) float32 RGBA, row 0 = bottom")
        rc = self._lib.srt_create(self.width, self.height, _p(sky), sky.shape[1], sky.shape[0], device,
                                  ctypes.byref(self._h))
        if rc:
            raise SrtError(self._lib.srt_last_error(None).decode())
        # RenderData(width, height): num_samples 4, num_bounces 10 (tracer.hpp:61-66)
        self.options = np.zeros(1, RENDER_DATA)
        self.options["width"], self.options["height"] = self.width, self.height
        self.options["num_samples"], self.options["num_bounces"] = 4, 10
        self.scene_data = np.zeros(1, SCENE_DATA)
        if os.environ.get("SRT_SWEEP_FILTER"):  # test / tuning hook: "auto", "one", "two" (bit-identical results)
            self.set_sweep_filter(os.environ["SRT_SWEEP_FILTER"])
        if os.environ.get("SRT_FRAME_PIPELINE"):  # test / tuning hook: "auto", "separate" (bit-identical results)
            self.set_frame_pipeline(os.environ["SRT_FRAME_PIPELINE"])
        if os.environ.get("SRT_SCHEDULE"):  # test / tuning hook: "auto", "plain", "wavefront" (bit-identical results)
            self.set_schedule(os.environ["SRT_SCHEDULE"])

    # -- reference surface ----------------------------------------------------------------------
    def update_scene(self, shapes, triangles, materials):
        shapes = np.ascontiguousarray(shapes, SHAPE)
        triangles = np.ascontiguousarray(triangles, TRIANGLE)
        materials = np.ascontiguousarray(materials, MATERIAL)
        self.scene_data["num_shapes"] = len(shapes)  # tracer.cpp:94
        self._check(self._lib.srt_upload_scene(self._h, _p(shapes), len(shapes), _p(triangles), len(triangles),
                                               _p(materials), len(materials), _p(self.scene_data)))

    def clear_canvas(self):
        self._check(self._lib.srt_clear(self._h))

    def render(self, ticks_stopped, output):
        """output: writable uint8 buffer of width*height*4 bytes (A,R,G,B per pixel)."""
        out = np.frombuffer(output, np.uint8) if not isinstance(output, np.ndarray) else output
        if out.size != self.width * self.height * 4 or not out.flags.c_contiguous:
            raise ValueError("output must hold width*height*4 contiguous bytes (src/main.cpp:128)")
        self._check(self._lib.srt_render_frame(self._h, _p(self.options), int(ticks_stopped), _p(out)))

    def pin_output(self, output):
        """srt_pin_output: page-lock the ONE output buffer the caller hands to render() every frame
        (src/main.cpp:128,290) so the read-back lands in it directly.  The caller keeps `output` alive until
        unpin_output() / close(); the tracer holds a reference for as long as it is pinned."""
        out = np.frombuffer(output, np.uint8) if not isinstance(output, np.ndarray) else output
        self._check(self._lib.srt_pin_output(self._h, _p(out), out.nbytes))
        self._pinned = out

    def unpin_output(self):
        self._check(self._lib.srt_unpin_output(self._h))
        self._pinned = None

    # -- harness surface ------------------------------------------------------------------------
    def accumulate(self, render_data=None):
        """The `render` kernel launch alone (asynchronous): canvas += mean of num_samples paths."""
        rd = self.options if render_data is None else np.ascontiguousarray(render_data, RENDER_DATA)
        self._check(self._lib.srt_render(self._h, _p(rd)))

    def accumulate_batch(self, render_datas):
        """srt_render_batch: the launches of `render_datas` (a sequence of 1-element RenderData records or one
        record array) in order; runs that differ only in `time` share one persistent kernel."""
        from .records import concat_records
        if isinstance(render_datas, np.ndarray) and render_datas.dtype == RENDER_DATA:
            rds = np.ascontiguousarray(render_datas).reshape(-1)
        else:
            rds = concat_records(RENDER_DATA, *render_datas)
        self._check(self._lib.srt_render_batch(self._h, _p(rds), len(rds)))

    def reserve_batch(self, render_data, n):
        """srt_reserve_batch: allocate the scratch of an n-launch batch ahead of time."""
        self._check(self._lib.srt_reserve_batch(self._h, _p(np.ascontiguousarray(render_data, RENDER_DATA)), int(n)))

    def accumulate_counted(self, render_data=None, counters=None):
        rd = self.options if render_data is None else np.ascontiguousarray(render_data, RENDER_DATA)
        cnt = np.zeros(1, COUNTERS) if counters is None else counters
        self._check(self._lib.srt_render_counted(self._h, _p(rd), _p(cnt)))
        return cnt

    def resolve(self, num_steps, output=None):
        out = np.empty((self.height, self.width, 4), np.uint8) if output is None else output
        self._check(self._lib.srt_resolve(self._h, int(num_steps), _p(out)))
        return out

    def resolve_device(self, num_steps):
        self._check(self._lib.srt_resolve_device(self._h, int(num_steps)))

    def resolve_device_range(self, num_steps, first_pixel, count):
        self._check(self._lib.srt_resolve_device_range(self._h, int(num_steps), int(first_pixel), int(count)))

    def read_output(self, output=None):
        out = np.empty((self.height, self.width, 4), np.uint8) if output is None else output
        self._check(self._lib.srt_read_output(self._h, _p(out)))
        return out

    def set_row_bands(self, band_height, band_index, band_count):
        self._check(self._lib.srt_set_row_bands(self._h, band_height, band_index, band_count))

    def set_accel(self, accel):
        """srt_set_accel: "none" (default: the reference's brute-force triangle loop, bit-exact) or "bvh" (labelled
        extension outside the parity path)."""
        self._check(self._lib.srt_set_accel(self._h, {"none": 0, "bvh": 1}[accel]))

    def set_schedule(self, schedule):
        """srt_set_schedule: "auto" | "plain" | "wavefront" -- how scenes without large models schedule a warp's work."""
        self._check(self._lib.srt_set_schedule(self._h, {"auto": 0, "plain": 1, "wavefront": 2}[schedule]))

    def set_frame_pipeline(self, mode):
        """srt_set_frame_pipeline: "auto" | "separate" -- whether render() into the pinned vector may run accumulate +
        average + read-back as one epilogue kernel, or always takes the separate steps."""
        self._check(self._lib.srt_set_frame_pipeline(self._h, {"auto": 0, "separate": 1}[mode]))

    def set_sweep_filter(self, mode):
        """srt_set_sweep_filter: "auto" | "one" | "two" -- which conservative filter precedes the exact triangle test."""
        self._check(self._lib.srt_set_sweep_filter(self._h, {"auto": 0, "one": 1, "two": 2}[mode]))

    def read_canvas(self):
        out = np.empty((self.height, self.width, 4), np.float32)
        self._check(self._lib.srt_read_canvas(self._h, _p(out)))
        return out

    def write_canvas(self, canvas):
        c = np.ascontiguousarray(canvas, np.float32)
        assert c.size == self.width * self.height * 4
        self._check(self._lib.srt_write_canvas(self._h, _p(c)))

    def canvas_view(self):
        ptr, n = ctypes.c_void_p(), ctypes.c_size_t()
        self._check(self._lib.srt_canvas_device_ptr(self._h, ctypes.byref(ptr), ctypes.byref(n)))
        return _CudaView(ptr.value, (self.height, self.width, 4), "<f4", self)

    def output_view(self):
        ptr, n = ctypes.c_void_p(), ctypes.c_size_t()
        self._check(self._lib.srt_output_device_ptr(self._h, ctypes.byref(ptr), ctypes.byref(n)))
        return _CudaView(ptr.value, (self.height, self.width, 4), "|u1", self)

    def stream_handle(self):
        s = ctypes.c_void_p()
        self._check(self._lib.srt_stream(self._h, ctypes.byref(s)))
        return s.value

    def synchronize(self):
        self._check(self._lib.srt_synchronize(self._h))

    def debug_primary(self, render_data=None):
        rd = self.options if render_data is None else np.ascontiguousarray(render_data, RENDER_DATA)
        idx = np.empty((self.height, self.width), np.int32)
        t = np.empty((self.height, self.width), np.float32)
        self._check(self._lib.srt_debug_primary(self._h, _p(rd), _p(idx), _p(t)))
        return idx, t

    def debug_math(self, op, x, y=None):
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y if y is not None else np.zeros_like(x), np.float32)
        out = np.empty_like(x)
        self._check(self._lib.srt_debug_math(self._h, self.MATH_OPS[op], _p(x), _p(y), _p(out), x.size))
        return out

    def measure_fp32_peak(self):
        tf, mhz = ctypes.c_double(), ctypes.c_double()
        self._check(self._lib.srt_measure_fp32_peak(self._h, ctypes.byref(tf), ctypes.byref(mhz)))
        return tf.value, mhz.value

    def render_time_ms(self):
        ms, n = ctypes.c_double(), ctypes.c_uint64()
        self._check(self._lib.srt_render_time_ms(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.srt_destroy(self._h)  # unpins before the buffer reference is dropped
            self._h = ctypes.c_void_p()
        self._pinned = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise SrtError(self._lib.srt_last_error(self._h).decode() or f"srt error {rc}")


# -- mesh / image I/O (reference include/parser.hpp:14-28) --------------------------------------
def _load(fn_name, path):
    L = load_library()
    ptr, n = ctypes.c_void_p(), ctypes.c_size_t()
    rc = getattr(L, fn_name)(os.fsencode(path), ctypes.byref(ptr), ctypes.byref(n))
    if rc:
        return None  # the reference returns std::nullopt
    try:
        if n.value == 0:
            return np.zeros(0, TRIANGLE)
        buf = (ctypes.c_char * (n.value * TRIANGLE.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, TRIANGLE).copy()
    finally:
        L.srt_free(ptr)


def load_stl_model(path, triangles):
    """load_stl_model(filename, triangles) -> (first_index, count) or None; returns the new array too."""
    new = _load("srt_load_stl", path)
    if new is None:
        return None
    first = len(triangles)
    return (first, len(new)), concat_records(TRIANGLE, triangles, new)


def load_obj_model(path, triangles):
    new = _load("srt_load_obj", path)
    if new is None:
        return None
    first = len(triangles)
    return (first, len(new)), concat_records(TRIANGLE, triangles, new)


def load_skybox_png(path):
    """srt_load_skybox_png: the (h, w, 4) float32 array Tracer(width, height, skybox) takes, decoded from a PNG the way
    the reference's constructor does (src/tracer.cpp:42-52).  None if the file cannot be read."""
    L = load_library()
    ptr, w, h = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
    if L.srt_load_skybox_png(os.fsencode(path), ctypes.byref(ptr), ctypes.byref(w), ctypes.byref(h)):
        return None
    try:
        buf = (ctypes.c_float * (w.value * h.value * 4)).from_address(ptr.value)
        return np.frombuffer(buf, np.float32).reshape(h.value, w.value, 4).copy()
    finally:
        L.srt_free(ptr)


def save_ppm(path, pixels, width, height):
    px = np.ascontiguousarray(pixels, np.uint8)
    assert px.size == width * height * 4
    rc = load_library().srt_save_ppm(os.fsencode(path), _p(px), width, height)
    if rc:
        raise SrtError(f"cannot write {path}")


def model_bounds(shape_record, triangles):
    """Model::compute_bounding_box (src/shape.cpp:45-58) on a SHAPE record of type model, in place."""
    tris = np.ascontiguousarray(triangles, TRIANGLE)
    rec = np.ascontiguousarray(shape_record, SHAPE).reshape(1).copy()
    raw = rec.view(np.uint8)
    model = np.ascontiguousarray(raw[16:128])
    rc = load_library().srt_model_bounds(_p(tris), len(tris), _p(model))
    if rc:
        raise SrtError("model triangle range out of bounds")
    raw[16:128] = model
    return rec[0]
