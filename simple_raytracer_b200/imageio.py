"""Sky-box image input the way the reference prepares it (src/tracer.cpp:42-52): an 8-bit image is decoded,
flipped vertically (`stbi_set_flip_vertically_on_load(1)`: memory row 0 = image bottom), expanded to 4 channels
and converted to float by stb_image's `stbi__ldr_to_hdr` rule (lib/stb_image.h:1868):
colour = (float)pow(x / 255.0f, 2.2f), alpha = x / 255.0f.  Pillow does the PNG decode here (stb_image is the
reference's vendored third-party code and is not copied)."""
import numpy as np


def ldr_to_hdr_lut():
    x = np.arange(256, dtype=np.float32) / np.float32(255.0)
    return np.power(x.astype(np.float64), float(np.float32(2.2))).astype(np.float32)


def skybox_from_rgb8(img8):
    """img8: (h, w, 3|4) uint8 in image order (row 0 = top).  Returns (h, w, 4) float32, row 0 = bottom."""
    img8 = np.asarray(img8, np.uint8)
    if img8.ndim != 3 or img8.shape[2] not in (3, 4):
        raise ValueError("expected an (h, w, 3|4) uint8 image")
    h, w, c = img8.shape
    out = np.ones((h, w, 4), np.float32)
    out[..., :3] = ldr_to_hdr_lut()[img8[..., :3]]
    if c == 4:
        out[..., 3] = img8[..., 3].astype(np.float32) / np.float32(255.0)
    return np.ascontiguousarray(out[::-1])


def load_skybox_png(path):
    """The array Tracer(width, height, skybox) expects, from an image file such as assets/skybox.png."""
    from PIL import Image
    with Image.open(path) as im:
        im = im.convert("RGBA" if im.mode in ("RGBA", "LA", "PA") else "RGB")
        return skybox_from_rgb8(np.asarray(im))


def save_png(path, argb, width, height):
    """A,R,G,B bytes (the layout Tracer.render fills) -> RGB PNG."""
    from PIL import Image
    px = np.asarray(argb, np.uint8).reshape(height, width, 4)
    Image.fromarray(np.ascontiguousarray(px[..., 1:])).save(path)
