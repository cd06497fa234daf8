"""The known-answer facts the reference's own test project holds for the code either side of the hot path, asserted
through the front end's probes: the four `Triangle.slice` facts (FuncTracer.Tests/Geometry/Triangle.Tests.fs:12-54;
every clipped mesh triangle of the BSP workloads comes out of that function) and the three colour-parser facts
(FuncTracer.Tests/Parser/Colour.fs:17-30).  CPU only."""
import numpy as np
import pytest

from functracer_b200 import frontend

# Triangle.Tests.fs:13-16
A = (-1.0, 0.0, 0.0)
B = (1.0, 2.0, 0.0)
C_ = (1.0, 0.0, 0.0)
PLANE_P0 = (0.0, 0.0, 0.0)
PLANE_N = (-1.0, 0.0, 0.0)


def _tri(*pts):
    return np.array(pts, dtype=np.float64)


def _slice_example(a, b, c):  # Triangle.Tests.fs:20-28
    ab_intercept = (0.0, 1.0, 0.0)
    ac_intercept = (0.0, 0.0, 0.0)
    above, below = frontend.slice_triangle(PLANE_P0, PLANE_N, _tri(a, b, c))
    assert len(above) >= 1 and len(below) >= 2
    t1, t2, t3 = above[0], below[0], below[1]
    # Assert.Equal on Triangle records = exact structural equality of the three points
    assert np.array_equal(t1, _tri(A, ab_intercept, ac_intercept))
    assert np.array_equal(t2, _tri(ab_intercept, B, C_))
    assert np.array_equal(t3, _tri(C_, ac_intercept, ab_intercept))


def test_ref_slice_bisected_triangle_is_split_into_three():  # Triangle.Tests.fs:30-32
    _slice_example(A, B, C_)


def test_ref_slice_same_result_for_every_rotation_of_the_points():  # Triangle.Tests.fs:34-38
    _slice_example(A, B, C_)
    _slice_example(C_, A, B)
    _slice_example(B, C_, A)


def test_ref_slice_triangle_above_is_returned_unchanged_in_fst():  # Triangle.Tests.fs:40-46
    t = _tri((-1.0, 0.0, 0.0), (-2.0, 1.0, 0.0), (-1.0, 1.0, 0.0))
    above, _ = frontend.slice_triangle(PLANE_P0, PLANE_N, t)
    assert np.array_equal(above[0], t)


def test_ref_slice_triangle_below_is_returned_unchanged_in_snd():  # Triangle.Tests.fs:48-54
    t = _tri((1.0, 0.0, 0.0), (2.0, 1.0, 0.0), (1.0, 1.0, 0.0))
    _, below = frontend.slice_triangle(PLANE_P0, PLANE_N, t)
    assert np.array_equal(below[0], t)


def test_ref_colour_triple_is_rgb():  # Parser/Colour.fs:17-20
    assert frontend.parse_colour("(1,0,0)") == (1.0, 0.0, 0.0)


def test_ref_colour_single_float_is_grey():  # Parser/Colour.fs:22-25
    assert frontend.parse_colour("1") == (1.0, 1.0, 1.0)


def test_ref_colour_hex_notation():  # Parser/Colour.fs:27-30
    assert frontend.parse_colour("#ff0000") == (1.0, 0.0, 0.0)


def test_colour_parser_rejects_garbage():
    with pytest.raises(frontend.SceneParseError):
        frontend.parse_colour("red")
