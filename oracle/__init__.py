"""ctypes binding of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Three libraries, all built by oracle/Makefile:
  liboracle.so              oracle.c, the function-by-function restatement of render.cl (with debug outputs
                            and work counters the kernel does not have);
  liboracle_contract.so     the same with every a*b+c expression site fused (sensitivity variant);
  _ref/libref_render_cl.so  the reference's OWN kernel source /root/reference/src/render.cl compiled by g++
                            (ref_build/), i.e. the reference run on the CPU.  Built in the authoring container
                            (where /root/reference exists); the .so travels to the GPU box.

May be imported by tests/, __graft_entry__ and bench.py's cpu_baseline / --impl reference legs -- never by
the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_CONTRACT_PATH = os.path.join(_HERE, "liboracle_contract.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libref_render_cl.so")
REFERENCE_KERNEL = os.environ.get("SRT_REFERENCE_KERNEL", "/root/reference/src/render.cl")
_lib = None
_libs = {}
_ref = None

COUNTERS = np.dtype([("samples", "u8"), ("bounces", "u8"), ("tri_tests", "u8"),
                     ("aabb_pass", "u8"), ("hits", "u8"), ("sky", "u8")])


def _stale(target, sources):
    return not os.path.exists(target) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(target) for f in sources)


def build(force=False):
    """Compile oracle.c (both variants) with the committed Makefile (gcc only)."""
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle_math.h", "Makefile")]
    if force or _stale(_LIB_PATH, src) or _stale(_CONTRACT_PATH, src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so", "liboracle_contract.so"])
    return _LIB_PATH


def build_ref(force=False):
    """Compile the reference kernel source into oracle/_ref/ when it is present (authoring container).
    Returns the library path, or None where neither the source nor a prebuilt library exists."""
    if os.path.exists(REFERENCE_KERNEL):
        src = [os.path.join(_HERE, "ref_build", f) for f in ("ref_driver.cpp", "cl_shim.hpp", "rewrite_cl.py")]
        src += [os.path.join(_HERE, "oracle_math.h"), os.path.join(_HERE, "Makefile"), REFERENCE_KERNEL]
        if force or _stale(_REF_PATH, src):
            subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "ref",
                                   "REF_CL=" + REFERENCE_KERNEL])
    return _REF_PATH if os.path.exists(_REF_PATH) else None


def ref_available():
    return os.path.exists(_REF_PATH) or os.path.exists(REFERENCE_KERNEL)


def ref_lib():
    """The reference kernel compiled for the CPU (None-safe: raises FileNotFoundError if it cannot exist)."""
    global _ref
    if _ref is None:
        path = build_ref()
        if path is None:
            raise FileNotFoundError("oracle/_ref/libref_render_cl.so is missing and %s is not present" % REFERENCE_KERNEL)
        L = ctypes.CDLL(path)
        vp, i32, u32, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_size_t
        L.ref_render.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32]
        L.ref_render.restype = None
        L.ref_average.argtypes = [u32, vp, vp, sz]
        L.ref_average.restype = None
        L.ref_max_threads.restype = i32
        _ref = L
    return _ref


def lib(variant="default"):
    """variant: "default" (liboracle.so) or "contract" (liboracle_contract.so)."""
    global _lib
    if variant != "default":
        if variant not in _libs:
            build()
            L = ctypes.CDLL(_CONTRACT_PATH)
            vp, i32, u32, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_size_t
            L.oracle_render.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32,
                                        i32, i32, i32, i32, vp]
            L.oracle_render.restype = None
            L.oracle_average.argtypes = [u32, vp, vp, sz]
            L.oracle_average.restype = None
            _libs[variant] = L
        return _libs[variant]
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i32, u32, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_size_t
        L.oracle_render.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32,
                                    i32, i32, i32, i32, vp]
        L.oracle_render.restype = None
        L.oracle_primary.argtypes = [vp, vp, vp, vp, vp, vp, i32]
        L.oracle_primary.restype = None
        L.oracle_average.argtypes = [u32, vp, vp, sz]
        L.oracle_average.restype = None
        L.oracle_seed.argtypes = [u32, u32, u32, u32]
        L.oracle_seed.restype = u32
        L.oracle_random_float.argtypes = [ctypes.POINTER(u32), ctypes.POINTER(u32)]
        L.oracle_random_float.restype = ctypes.c_float
        L.oracle_math.argtypes = [i32, vp, vp, vp, sz]
        L.oracle_math.restype = None
        L.oracle_intersect.argtypes = [i32, vp, vp, vp, vp, vp, vp]
        L.oracle_intersect.restype = i32
        L.oracle_max_threads.restype = i32
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _c(a, dtype=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def max_threads():
    return int(lib().oracle_max_threads())


def render(render_data, scene_data, shapes, triangles, materials, sky, canvas=None, window=None,
           bands=None, threads=0, impl="oracle"):
    """One launch of kernel `render` (render.cl:483-523): canvas += mean of num_samples paths.

    render_data / scene_data: 1-element record arrays.  sky: (h, w, 4) float32.  canvas:
    (height, width, 4) float32, created zeroed when None.  window = (x0, y0, x1, y1) restricts
    the pixels rendered (global ids preserved).  bands = (band_h, band_i, band_n).
    impl: "oracle" (oracle.c), "contract" (its fused variant) or "ref" (the reference's render.cl itself,
    oracle/_ref; counters are all zero there because the kernel has none).
    Returns (canvas, counters).
    """
    rd = _c(render_data)
    sd = _c(scene_data).copy()
    sd["num_shapes"] = len(shapes)
    w, h = int(rd["width"].reshape(-1)[0]), int(rd["height"].reshape(-1)[0])
    if canvas is None:
        canvas = np.zeros((h, w, 4), np.float32)
    assert canvas.dtype == np.float32 and canvas.flags.c_contiguous and canvas.size == w * h * 4
    shapes, triangles, materials = _c(shapes), _c(triangles), _c(materials)
    sky = _c(sky, np.float32)
    x0, y0, x1, y1 = window if window is not None else (0, 0, w, h)
    bh, bi, bn = bands if bands is not None else (1, 0, 1)
    cnt = np.zeros(1, COUNTERS)
    if impl == "ref":
        ref_lib().ref_render(_p(rd), _p(sd), _p(canvas), _p(shapes), _p(triangles), _p(materials),
                             _p(sky), sky.shape[1], sky.shape[0], x0, y0, x1, y1, bh, bi, bn, threads)
    else:
        lib("default" if impl == "oracle" else impl).oracle_render(
            _p(rd), _p(sd), _p(canvas), _p(shapes), _p(triangles), _p(materials),
            _p(sky), sky.shape[1], sky.shape[0], x0, y0, x1, y1, bh, bi, bn, threads, _p(cnt))
    return canvas, cnt[0]


def primary(render_data, scene_data, shapes, triangles, threads=0):
    """Shape index (-1 = miss) and t of every pixel's sample-0 camera ray."""
    rd = _c(render_data)
    sd = _c(scene_data).copy()
    sd["num_shapes"] = len(shapes)
    w, h = int(rd["width"].reshape(-1)[0]), int(rd["height"].reshape(-1)[0])
    idx = np.empty((h, w), np.int32)
    t = np.empty((h, w), np.float32)
    shapes, triangles = _c(shapes), _c(triangles)
    lib().oracle_primary(_p(rd), _p(sd), _p(shapes), _p(triangles), _p(idx), _p(t), threads)
    return idx, t


def average(num_steps, canvas, impl="oracle"):
    """Kernel `average` (render.cl:525-535): returns (..., 4) uint8 in A,R,G,B byte order."""
    canvas = _c(canvas, np.float32)
    n = canvas.size // 4
    out = np.empty(canvas.shape[:-1] + (4,), np.uint8)
    if impl == "ref":
        ref_lib().ref_average(int(num_steps), _p(canvas), _p(out), n)
    else:
        lib("default" if impl == "oracle" else impl).oracle_average(int(num_steps), _p(canvas), _p(out), n)
    return out


def seed(sample, pixel_id, num_samples, time):
    return int(lib().oracle_seed(sample, pixel_id, num_samples, time))


def random_floats(seed0, n):
    """n successive (state, hash, float) triples of random_float (render.cl:143-148)."""
    s = ctypes.c_uint32(seed0)
    h = ctypes.c_uint32(0)
    out = []
    for _ in range(n):
        f = lib().oracle_random_float(ctypes.byref(s), ctypes.byref(h))
        out.append((s.value, h.value, float(f)))
    return out


MATH_OPS = {"log": 0, "cos": 1, "atan2pi": 2, "pow": 3, "sqrt": 4, "schlick": 5}


def math(op, x, y=None):
    x = _c(x, np.float32)
    y = _c(y if y is not None else np.zeros_like(x), np.float32)
    out = np.empty_like(x)
    lib().oracle_math(MATH_OPS[op], _p(x), _p(y), _p(out), x.size)
    return out


def intersect(kind, o, d, a, b, c=(0, 0, 0)):
    kinds = {"sphere": 0, "plane": 1, "triangle": 2, "aabb": 3}
    arrs = [np.asarray(v, np.float32) for v in (o, d, a, b, c)]
    arrs = [np.concatenate([v, np.zeros(3 - v.size, np.float32)]) if v.size < 3 else v for v in arrs]
    t = np.zeros(1, np.float32)
    hit = lib().oracle_intersect(kinds[kind], *[_p(v) for v in arrs], _p(t))
    return bool(hit), float(t[0])
