"""Seeded synthetic scenes for the five BASELINE.json configs.

The reference ships no scene files: its scenes are clicked together in the GUI
(/root/reference/src/interface.cpp:148-168) and its three showcase images
(/root/reference/readme/*.png) record no parameters.  These builders produce the same kinds
of scene as records byte-compatible with the reference structs (records.py), using the
reference's defaults: camera (0,0,5) looking down -Z, fov 90 deg, sun focus 25 / colour
0xffffd3 / intensity 1 / direction normalize(1,-1,0) (src/main.cpp:106-126).
"""
from dataclasses import dataclass, field

import numpy as np

from .records import (MATERIAL, RENDER_DATA, SCENE_DATA, SHAPE, SHAPE_MODEL, SHAPE_PLANE,
                      SHAPE_SPHERE, TRIANGLE, concat_records)  # noqa: F401

F = np.float32


# ----------------------------------------------------------------------------- helpers
def hex_color(v):
    """color::from_hex, /root/reference/include/color.hpp:12-14."""
    return np.array([((v >> 16) & 255) / F(255), ((v >> 8) & 255) / F(255), (v & 255) / F(255)], F)


def translate(p):
    m = np.eye(4, dtype=F)
    m[:3, 3] = p
    return m


def scale(s):
    s = np.broadcast_to(np.asarray(s, F), (3,))
    return np.diag(np.array([s[0], s[1], s[2], 1], F))


def rotate_y(a):
    c, s = F(np.cos(a)), F(np.sin(a))
    return np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], F)


def rotate_x(a):
    c, s = F(np.cos(a)), F(np.sin(a))
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], F)


def camera_matrix(position, yaw=0.0, pitch=0.0):
    """Camera::camera_matrix, /root/reference/include/helper.hpp:21-26:
    translate(position) * eulerAngleYXZ(yaw, pitch, 0).  Returned as a row-major 4x4 (math
    convention); use to_columns() for the column-major record field."""
    return (translate(position) @ rotate_y(yaw) @ rotate_x(pitch)).astype(F)


def to_columns(m):
    """math 4x4 -> glm column-major float[4][4] (element [c][r])."""
    return np.ascontiguousarray(np.asarray(m, F).T)


def material(color=(1, 1, 1), smoothness=0.0, metallic=0.0, specular=0.0, transmittance=0.0,
             refraction_index=1.0, emission=(0, 0, 0), emission_strength=0.0):
    """Material ctor defaults, /root/reference/include/material.hpp:23-27."""
    m = np.zeros((), MATERIAL)
    m["color"], m["emission"] = color, emission
    m["smoothness"], m["metallic"], m["specular"] = smoothness, metallic, specular
    m["transmittance"], m["refraction_index"] = transmittance, refraction_index
    m["emission_strength"] = emission_strength
    return m


def sphere(mat, position, radius):
    s = np.zeros((), SHAPE)
    s["type"], s["material"] = SHAPE_SPHERE, mat
    s["sphere_position"], s["sphere_radius"] = position, radius
    return s


def plane(mat, position, normal):
    s = np.zeros((), SHAPE)
    s["type"], s["material"] = SHAPE_PLANE, mat
    s["plane_position"], s["plane_normal"] = position, normal
    return s


def model(mat, triangles, first, count, transform=None):
    """Model(triangles, first, count) + a UI edit: transform = T*R*S and the AABB recomputed
    over the transformed vertices (/root/reference/src/shape.cpp:37-58, src/interface.cpp:98-101)."""
    s = np.zeros((), SHAPE)
    s["type"], s["material"] = SHAPE_MODEL, mat
    s["model_triangle_index"], s["model_num_triangles"] = first, count
    t = np.eye(4, dtype=F) if transform is None else np.asarray(transform, F)
    s["model_transform"] = to_columns(t)
    pos = triangles["v"]["pos"][first:first + count].reshape(-1, 3).astype(F)
    world = (pos @ t[:3, :3].T + t[:3, 3]).astype(F)
    s["model_bounding_min"] = world.min(axis=0)
    s["model_bounding_max"] = world.max(axis=0)
    return s


def triangles_from(positions, normals):
    """positions, normals: (n, 3, 3) -> TRIANGLE records."""
    t = np.zeros(len(positions), TRIANGLE)
    t["v"]["pos"] = np.asarray(positions, F)
    t["v"]["normal"] = np.asarray(normals, F)
    return t


def cube_triangles():
    """12 flat-shaded triangles of the cube [-1,1]^3 with outward normals: what
    Box::create_triangle provides (/root/reference/src/shape.cpp:91-119), built per face."""
    pos, nrm = [], []
    for axis in range(3):
        for sgn in (-1.0, 1.0):
            u, v = (axis + 1) % 3, (axis + 2) % 3
            c = np.zeros((4, 3), F)
            c[:, axis] = sgn
            c[:, u] = [-1, 1, 1, -1]
            c[:, v] = [-1, -1, 1, 1]
            n = np.zeros(3, F)
            n[axis] = sgn
            for tri in ((0, 1, 2), (0, 2, 3)):
                pos.append(c[list(tri)])
                nrm.append(np.tile(n, (3, 1)))
    return triangles_from(np.array(pos), np.array(nrm))


def reference_cube_triangles():
    """The same 12 triangles in the order Box::create_triangle emits them (/root/reference/src/shape.cpp:91-119;
    C++ twin: include/scene.hpp Box::create_triangle): corner i at (i&4 ? +1 : -1, i&1 ? +1 : -1, i&2 ? -1 : +1)."""
    faces = "120362746504602357132376754510640315"
    corner = [np.array([1 if i & 4 else -1, 1 if i & 1 else -1, -1 if i & 2 else 1], F) for i in range(8)]
    pos, nrm = [], []
    for k in range(12):
        v1, v2, v3 = (corner[int(c)] for c in faces[3 * k:3 * k + 3])
        a, b = v2 - v1, v3 - v1
        n = np.array([a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]], F)
        if not (v1[0] * n[0] + v1[1] * n[1] + v1[2] * n[2]) > 0:
            n = -n
        n = n * (F(1) / np.sqrt(F(n[0] * n[0] + n[1] * n[1] + n[2] * n[2])))
        pos.append([v1, v2, v3])
        nrm.append([n, n, n])
    return triangles_from(np.array(pos, F), np.array(nrm, F))


def box_model(mat, position, size=2.0):
    """Shape{mat, Box::model(position, size)}, /root/reference/src/shape.cpp:76-89: transform = translate(position)
    only, AABB = position -+ size / 2, triangles [0, 12)."""
    s = np.zeros((), SHAPE)
    s["type"], s["material"] = SHAPE_MODEL, mat
    s["model_triangle_index"], s["model_num_triangles"] = 0, 12
    p = np.asarray(position, F)
    half = np.broadcast_to(np.asarray(size, F), (3,)) * F(0.5)
    s["model_bounding_min"], s["model_bounding_max"] = p - half, p + half
    s["model_transform"] = to_columns(translate(p))
    return s


def icosphere(subdiv):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
                  [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                  [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
                  [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]])
    for _ in range(subdiv):
        cache, verts, nf = {}, list(v), []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[k] = len(verts) - 1
            return cache[k]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        v, f = np.array(verts), np.array(nf)
    return v, f


def smooth_normals(v, f):
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    n = np.zeros_like(v)
    for k in range(3):
        np.add.at(n, f[:, k], fn)
    return n / np.linalg.norm(n, axis=1, keepdims=True)


def noisy_icosphere(subdiv, seed, amplitude=0.12):
    """Closed mesh with smooth normals: icosphere with seeded low-frequency radial noise
    (subdiv 3 -> 1 280 triangles: the '~1k-triangle Suzanne-class' mesh of config 3)."""
    v, f = icosphere(subdiv)
    rng = np.random.default_rng(seed)
    k = rng.normal(size=(6, 3)) * 2.5
    ph = rng.uniform(0, 2 * np.pi, 6)
    r = 1.0 + amplitude * np.mean(np.sin(v @ k.T + ph), axis=1) * 2.0
    v = v * r[:, None]
    n = smooth_normals(v, f)
    return v.astype(F), n.astype(F), f


def displaced_torus(nu, nv, seed, R=2.0, r=0.8, amplitude=0.08):
    """nu*nv*2 triangles (224x224 -> 100 352): the ~100k-triangle stress mesh of config 5."""
    rng = np.random.default_rng(seed)
    u = np.arange(nu) * (2 * np.pi / nu)
    w = np.arange(nv) * (2 * np.pi / nv)
    U, W = np.meshgrid(u, w, indexing="ij")
    k = rng.integers(2, 9, size=(4, 2))
    ph = rng.uniform(0, 2 * np.pi, 4)
    d = sum(np.sin(k[i, 0] * U + k[i, 1] * W + ph[i]) for i in range(4)) / 4.0
    rr = r * (1.0 + amplitude * d)
    P = np.stack([(R + rr * np.cos(W)) * np.cos(U), rr * np.sin(W), (R + rr * np.cos(W)) * np.sin(U)], -1)
    v = P.reshape(-1, 3)
    idx = np.arange(nu * nv).reshape(nu, nv)
    a, b = idx, np.roll(idx, -1, 0)
    c, d2 = np.roll(idx, -1, 1), np.roll(np.roll(idx, -1, 0), -1, 1)
    f = np.concatenate([np.stack([a, c, b], -1).reshape(-1, 3), np.stack([b, c, d2], -1).reshape(-1, 3)])
    return v.astype(F), f


def mesh_triangles(v, f, n=None):
    """Indexed mesh -> TRIANGLE records; n=None gives flat facet normals (what an .stl carries,
    /root/reference/src/parser.cpp:46-50)."""
    pos = v[f]
    if n is None:
        fn = np.cross(pos[:, 1] - pos[:, 0], pos[:, 2] - pos[:, 0]).astype(np.float64)
        fn /= np.linalg.norm(fn, axis=1, keepdims=True)
        nrm = np.repeat(fn[:, None, :], 3, axis=1)
    else:
        nrm = n[f]
    return triangles_from(pos, nrm)


def procedural_skybox(width=2048, height=1024, seed=7):
    """RGBA float32 (height, width, 4), row 0 = v 0 = straight down.  Stand-in for
    assets/skybox.png as the reference uploads it (src/tracer.cpp:42-52): 8-bit texels pushed
    through stb_image's pow(x/255, 2.2) (lib/stb_image.h:1868), alpha 1, rows bottom-up."""
    rng = np.random.default_rng(seed)
    v = (np.arange(height) + 0.5) / height
    up = np.clip((v - 0.5) * 2.0, 0, 1)[:, None]
    down = np.clip((0.5 - v) * 2.0, 0, 1)[:, None]
    horizon = np.array([0.75, 0.82, 0.90])
    zenith = np.array([0.20, 0.42, 0.80])
    ground = np.array([0.35, 0.32, 0.30])
    sky = horizon * (1 - up ** 0.5) + zenith * up ** 0.5
    gnd = horizon * (1 - down ** 0.35) + ground * down ** 0.35
    base = np.where(v[:, None] >= 0.5, sky, gnd)[:, None, :].repeat(width, 1)
    # soft clouds: sum of bilinearly upsampled seeded value-noise octaves, periodic in u
    clouds = np.zeros((height, width))
    for o, amp in ((8, 0.5), (16, 0.3), (32, 0.2)):
        g = rng.random((o // 2 + 1, o))
        g = np.concatenate([g, g[:, :1]], 1)
        y = np.linspace(0, o // 2, height, endpoint=False)
        x = np.linspace(0, o, width, endpoint=False)
        y0, x0 = y.astype(int), x.astype(int)
        fy, fx = (y - y0)[:, None], (x - x0)[None, :]
        fy, fx = fy * fy * (3 - 2 * fy), fx * fx * (3 - 2 * fx)
        clouds += amp * ((g[y0][:, x0] * (1 - fx) + g[y0][:, x0 + 1] * fx) * (1 - fy) +
                         (g[y0 + 1][:, x0] * (1 - fx) + g[y0 + 1][:, x0 + 1] * fx) * fy)
    cover = np.clip((clouds - 0.52) * 3.0, 0, 1) * np.clip((v[:, None] - 0.52) * 6.0, 0, 1)
    img = base * (1 - cover[..., None]) + cover[..., None] * 0.97
    img8 = np.clip(np.rint(img * 255.0), 0, 255).astype(np.uint8)
    lut = np.power(np.arange(256, dtype=F) / F(255.0), F(2.2)).astype(F)
    out = np.ones((height, width, 4), F)
    out[..., :3] = lut[img8]
    return out


# ----------------------------------------------------------------------------- scene record
@dataclass
class Scene:
    name: str
    width: int
    height: int
    num_samples: int
    num_bounces: int
    launches: int
    shapes: np.ndarray
    triangles: np.ndarray
    materials: np.ndarray
    camera: np.ndarray  # math-convention 4x4
    scene_data: np.ndarray = field(default=None)
    fov_scale: float = 1.0  # tan(90deg / 2), src/main.cpp:111-112

    def __post_init__(self):
        if self.scene_data is None:
            self.scene_data = default_scene_data()
        self.scene_data["num_shapes"] = len(self.shapes)

    @property
    def spp(self):
        return self.num_samples * self.launches

    def render_data(self, k=0, show_normals=False, width=None, height=None, num_samples=None,
                    num_bounces=None):
        """RenderData for launch k: time_k = 1 000 003 + k (odd, non-zero), tick = k."""
        rd = np.zeros(1, RENDER_DATA)
        w, h = width or self.width, height or self.height
        rd["width"], rd["height"] = w, h
        rd["num_samples"] = num_samples or self.num_samples
        rd["num_bounces"] = num_bounces or self.num_bounces
        rd["aspect_ratio"] = F(w) / F(h)  # src/main.cpp:109
        rd["fov_scale"] = self.fov_scale
        rd["show_normals"] = 1 if show_normals else 0
        rd["camera_to_world"] = to_columns(self.camera)
        rd["time"], rd["tick"] = 1000003 + k, k
        return rd

    def resized(self, width, height, **kw):
        d = dict(self.__dict__)
        d.update(width=width, height=height, **kw)
        return Scene(**d)


def default_scene_data():
    """src/main.cpp:120-126."""
    sd = np.zeros(1, SCENE_DATA)
    sd["horizon_color"], sd["zenith_color"] = hex_color(0x374F62), hex_color(0x11334A)
    sd["ground_color"], sd["sun_color"] = hex_color(0x777777), hex_color(0xFFFFD3)
    sd["sun_focus"], sd["sun_intensity"] = 25.0, 1.0
    d = np.array([1.0, -1.0, 0.0])
    sd["sun_direction"] = (d / np.linalg.norm(d)).astype(F)
    return sd


def _stack(records, dtype):
    out = np.zeros(len(records), dtype)
    for i, r in enumerate(records):
        out[i] = r
    return out


# ----------------------------------------------------------------------------- configs
def config1(width=800, height=600):
    """C1: red/green-wall room, planes + spheres + emissive ceiling box (readme/red_green.png)."""
    tris = cube_triangles()
    mats = [material((0.9, 0.9, 0.9)),                                               # 0 white
            material((0.85, 0.1, 0.1)),                                              # 1 red
            material((0.1, 0.8, 0.15)),                                              # 2 green
            material((1, 1, 1), emission=(1.0, 0.93, 0.8), emission_strength=8.0),   # 3 light
            material((0.95, 0.95, 0.95), smoothness=1.0, metallic=1.0),              # 4 mirror
            material((0.9, 0.75, 0.3), smoothness=0.9, metallic=1.0),                # 5 gold
            material((1, 1, 1), smoothness=1.0, transmittance=1.0, refraction_index=1.5)]  # 6 glass
    shapes = [plane(1, (-4, 0, 0), (1, 0, 0)), plane(2, (4, 0, 0), (-1, 0, 0)),
              plane(0, (0, -3, 0), (0, 1, 0)), plane(0, (0, 3, 0), (0, -1, 0)),
              plane(0, (0, 0, -6), (0, 0, 1)), plane(0, (0, 0, 7), (0, 0, -1)),
              model(3, tris, 0, 12, translate((0, 2.95, -2)) @ scale((1.2, 0.05, 1.2))),
              sphere(4, (-2.2, -2.0, -3.0), 1.0), sphere(5, (0.2, -2.0, -4.2), 1.0),
              sphere(6, (2.1, -1.9, -1.8), 1.1)]
    return Scene("C1 red/green room 800x600 1spp 8 bounces", width, height, 1, 8, 1,
                 _stack(shapes, SHAPE), tris, _stack(mats, MATERIAL), camera_matrix((0, 0, 5)))


def config2(width=1920, height=1080, num_samples=4, launches=16):
    """C2: four-sphere material scene + skybox (readme/spheres.png)."""
    mats = [material((0.8, 0.8, 0.8)),                                               # 0 floor
            material((0.85, 0.1, 0.1)), material((0.1, 0.8, 0.15)),                  # 1 red 2 green
            material((0.9, 0.9, 0.85)),                                              # 3 diffuse
            material((1, 1, 1), smoothness=1.0, transmittance=1.0, refraction_index=1.5),  # 4 glass
            material((0.25, 0.4, 0.95), smoothness=0.95, metallic=1.0),              # 5 metal blue
            material((1, 0.2, 0.15), emission=(1.0, 0.15, 0.1), emission_strength=5.0)]  # 6 emissive
    shapes = [plane(0, (0, -2, 0), (0, 1, 0)), plane(1, (-5, 0, -7), (0.8, 0, 0.6)),
              plane(2, (5, 0, -7), (-0.6, 0, 0.8)),
              sphere(3, (-3.2, 0.0, -3.0), 2.0), sphere(4, (0.6, -0.8, -0.5), 1.2),
              sphere(5, (3.4, -0.6, -2.6), 1.4), sphere(6, (-0.4, -1.3, -4.6), 0.7)]
    return Scene("C2 four spheres + skybox", width, height, num_samples, 10, launches,
                 _stack(shapes, SHAPE), np.zeros(0, TRIANGLE), _stack(mats, MATERIAL),
                 camera_matrix((0, 0.5, 5.5), 0.0, -0.08))


def config3(width=1920, height=1080, num_samples=4, launches=64, subdiv=3):
    """C3: two instances of one ~1k-triangle smooth mesh, refractive + mildly metallic
    (readme/model.png)."""
    v, n, f = noisy_icosphere(subdiv, seed=3)
    tris = mesh_triangles(v, f, n)
    mats = [material((0.8, 0.8, 0.8)),
            material((0.3, 0.9, 0.4), smoothness=0.9, transmittance=0.9, refraction_index=1.3),
            material((0.55, 0.6, 0.75), smoothness=0.6, metallic=0.3)]
    shapes = [plane(0, (0, -1.6, 0), (0, 1, 0)),
              model(1, tris, 0, len(tris), translate((-1.7, 0, -1.5)) @ rotate_y(0.6) @ scale(1.5)),
              model(2, tris, 0, len(tris), translate((1.9, -0.2, -2.2)) @ rotate_y(-0.9) @ rotate_x(0.3) @ scale((1.3, 1.4, 1.3)))]
    return Scene("C3 two ~1k-tri meshes", width, height, num_samples, 10, launches,
                 _stack(shapes, SHAPE), tris, _stack(mats, MATERIAL), camera_matrix((0, 0.3, 3.5)))


def config4(width=3840, height=2160, num_samples=16, launches=64):
    """C4: C2's scene at 4K, 1 024 spp, sample-sharded across GPUs."""
    s = config2(width, height, num_samples, launches)
    s.name = "C4 four spheres 4K progressive"
    return s


def config5(width=1920, height=1080, num_samples=4, launches=16, nu=224, nv=224):
    """C5: ~100k-triangle flat-shaded (.stl-style) stress mesh, 16 bounces, tile-sharded."""
    v, f = displaced_torus(nu, nv, seed=5)
    tris = mesh_triangles(v, f, None)
    mats = [material((0.8, 0.8, 0.8)), material((0.9, 0.55, 0.25), smoothness=0.8, specular=0.25)]
    shapes = [plane(0, (0, -1.5, 0), (0, 1, 0)),
              model(1, tris, 0, len(tris), translate((0, 0.1, -2.5)) @ rotate_x(0.9) @ rotate_y(0.3))]
    return Scene("C5 ~100k-tri stress mesh", width, height, num_samples, 16, launches,
                 _stack(shapes, SHAPE), tris, _stack(mats, MATERIAL), camera_matrix((0, 0.4, 3.0)))


CONFIGS = {1: config1, 2: config2, 3: config3, 4: config4, 5: config5}
