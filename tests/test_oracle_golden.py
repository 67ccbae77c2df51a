"""The oracle against the committed golden outputs of the REFERENCE KERNEL (tests/golden/make_golden.py ran
/root/reference/src/render.cl through oracle/_ref and stored inputs + outputs).  Runs anywhere: no reference
needed at test time."""
import numpy as np
import pytest

import golden_util
from conftest import assert_bit_equal


@pytest.mark.parametrize("cfg", golden_util.CONFIGS)
def test_oracle_reproduces_golden(oracle_lib, cfg):
    g = golden_util.load(cfg)
    rd = g["rd"]
    canvas = None
    for k in range(len(rd)):
        canvas, _ = oracle_lib.render(rd[k:k + 1], g["scene_data"], g["shapes"], g["triangles"], g["materials"],
                                      g["sky"], canvas)
    assert_bit_equal(g["canvas"], canvas, f"C{cfg}")
    assert np.array_equal(g["argb"], oracle_lib.average(len(rd), canvas))
    idx, t = oracle_lib.primary(g["primary_rd"], g["scene_data"], g["shapes"], g["triangles"])
    assert np.array_equal(idx, g["primary_idx"])
    assert_bit_equal(t, g["primary_t"], "primary t")


@pytest.mark.parametrize("key", ["config1", "config2"])
def test_oracle_reproduces_the_reference_kernel_at_full_baseline_size(oracle_lib, sky, key):
    """tests/golden/fullsize_hashes.json holds SHA-256 digests of full-size canvases rendered by the reference kernel
    (make_fullsize_hashes.py); oracle.c must hit them from the same seeded scene (no reference needed at test time)."""
    import hashlib
    import fullsize_util
    h = fullsize_util.load()
    e = h[key]
    if hashlib.sha256(sky.tobytes()).hexdigest() != h["sky_sha256"]:
        pytest.skip("procedural sky differs on this platform")
    sc, same = fullsize_util.scene_for(e)
    if not same:
        pytest.skip("scene builder produced different bytes on this platform")
    canvas = None
    for k in range(e["launches"]):
        canvas, _ = oracle_lib.render(sc.render_data(k, num_samples=e["num_samples"]), sc.scene_data, sc.shapes,
                                      sc.triangles, sc.materials, sky, canvas)
    assert fullsize_util.canvas_digest(canvas) == e["canvas_sha256"]
    assert hashlib.sha256(oracle_lib.average(e["launches"], canvas).tobytes()).hexdigest() == e["argb_sha256"]
