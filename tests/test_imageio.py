import numpy as np

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
