import numpy as np
import pytest

from simple_raytracer_b200 import imageio


def test_skybox_png_follows_stb_rule(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(6, 9, 3), dtype=np.uint8)
    img[0, 0] = [0, 255, 128]
    p = tmp_path / "sky.png"
    Image.fromarray(img).save(p)
    sky = imageio.load_skybox_png(str(p))
    assert sky.shape == (6, 9, 4) and sky.dtype == np.float32
    assert (sky[..., 3] == 1.0).all()
    # vertical flip: image row 0 (top) is memory row h-1
    top = sky[5, 0]
    assert top[0] == 0.0 and top[1] == 1.0
    assert abs(float(top[2]) - (128 / 255.0) ** 2.2) < 1e-7
    want = (img.astype(np.float64) / 255.0) ** 2.2
    assert np.allclose(sky[::-1, :, :3], want, atol=1e-6)


def test_skybox_rgba_alpha_is_linear(tmp_path):
    from PIL import Image
    img = np.zeros((2, 2, 4), np.uint8)
    img[..., 3] = [[0, 51], [102, 255]]
    p = tmp_path / "a.png"
    Image.fromarray(img, "RGBA").save(p)
    sky = imageio.load_skybox_png(str(p))
    assert np.allclose(sky[::-1, :, 3], img[..., 3] / 255.0, atol=1e-7)


def test_save_png_round_trip(tmp_path):
    from PIL import Image
    px = np.arange(3 * 4 * 4, dtype=np.uint8).reshape(3, 4, 4)
    p = tmp_path / "o.png"
    imageio.save_png(str(p), px, 4, 3)
    assert np.array_equal(np.asarray(Image.open(p)), px[..., 1:])


@pytest.mark.parametrize("mode", ["RGB", "RGBA", "L", "LA", "P"])
def test_native_png_reader_matches_the_stb_rule(tmp_path, mode):
    """srt_load_skybox_png (C++, zlib inflate + PNG unfilter) against the Pillow-based reader: the same RGBA-f32 texels
    bit for bit -- rows bottom-up, colour through (float)pow(x / 255.0f, 2.2f), alpha linear (lib/stb_image.h:1868-1874)."""
    from PIL import Image
    from simple_raytracer_b200 import tracer
    rng = np.random.default_rng(5)
    w, h = 67, 41  # odd sizes; smooth + noisy content so that the encoder uses several filter types
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 3 + yy) % 256, (xx + yy * 5) % 256, (xx * yy) % 256, 255 - (xx * 2) % 256], -1).astype(np.uint8)
    base[::3] = rng.integers(0, 256, base[::3].shape, dtype=np.uint8)
    img = {"RGB": Image.fromarray(base[..., :3], "RGB"), "RGBA": Image.fromarray(base, "RGBA"),
           "L": Image.fromarray(base[..., 0], "L"), "LA": Image.fromarray(base[..., [0, 3]], "LA"),
           "P": Image.fromarray(base[..., :3], "RGB").quantize(64)}[mode]
    path = tmp_path / f"sky_{mode}.png"
    img.save(path)
    got = tracer.load_skybox_png(str(path))
    want = imageio.load_skybox_png(str(path))
    assert got is not None and got.shape == (h, w, 4) and got.dtype == np.float32
    assert got.tobytes() == want.tobytes()
    assert tracer.load_skybox_png(str(tmp_path / "missing.png")) is None
    bad = tmp_path / "bad.png"
    bad.write_bytes(path.read_bytes()[:60])
    assert tracer.load_skybox_png(str(bad)) is None
