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
