"""FP32-specific restatements in the product kernels, replayed in numpy float32 against the oracle's literal double forms.
CPU only; the GPU parity tests measure the same thing end to end on frames."""
import numpy as np

from oracle import ftb_oracle as orc

F = np.float32


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _norm32(v):
    """CommonTypes.normalise in float32 (separate roundings)."""
    l = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]).astype(F)).astype(F)
    return np.where((l < F(1e-7))[:, None], v, (v * (F(1) / l)[:, None]).astype(F)).astype(F)


def _dot32(a, b):
    return (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2]).astype(F)


def _rough_diffuse_fp32(n, ld, vd, colour, roughness):
    """render.cuh roughDiffuse, FP32 branch: Oren-Nayar without forming the angles (Shading.fs:50-63 in real arithmetic):
    cos(lightAngle) = cl, sin(alpha) = sqrt((1 - ca)(1 + ca)) with ca the smaller cosine, tan(beta) = sqrt((1 - cb)(1 + cb)) / cb."""
    rough = (roughness * roughness).astype(F)
    A = (F(1) - F(0.5) * rough / (rough + F(0.33))).astype(F)
    B = (F(0.45) * rough / (rough + F(0.09))).astype(F)
    nn = _norm32(n)
    ml, mv = (-ld).astype(F), (-vd).astype(F)
    tl = _norm32((ml - _dot32(ml, nn)[:, None] * nn).astype(F))
    tr = _norm32((mv - _dot32(mv, nn)[:, None] * nn).astype(F))
    cr = np.clip(_dot32(nn, _norm32(mv)), F(-1), F(1))
    cl = np.clip(_dot32(nn, _norm32(ml)), F(-1), F(1))
    ca, cb = np.minimum(cr, cl), np.maximum(cr, cl)
    cb = np.where(np.abs(cb) < F(4e-8), np.where(cb < 0, F(-4e-8), F(4e-8)), cb).astype(F)
    sin_a = np.sqrt(((F(1) - ca) * (F(1) + ca)).astype(F)).astype(F)
    tan_b = (np.sqrt(((F(1) - cb) * (F(1) + cb)).astype(F)) / cb).astype(F)
    inten = (cl * (A + (B * np.maximum(F(0), _dot32(tl, tr)) * sin_a * tan_b))).astype(F)
    return (inten[:, None] * colour).astype(F)


def test_angle_free_oren_nayar_is_the_literal_formula():
    """Against the oracle's literal Shading.roughDiffuse in double on 20 000 random fragments: the FP32 form agrees to FP32 rounding
    wherever the literal form is well conditioned (tan(beta) amplifies any input error by 1 / cos^2(beta); beta is the SMALLER of
    the two angles, so that only happens when view and light both graze the surface), and stays finite everywhere."""
    rng = np.random.default_rng(5)
    n_cases = 20_000
    n = _unit(rng.normal(size=(n_cases, 3)))
    vd = -_unit(n + 1.5 * rng.normal(size=(n_cases, 3)))  # viewRay.d: mostly towards the surface, some from behind
    ld = -_unit(n + 1.5 * rng.normal(size=(n_cases, 3)))
    vd *= rng.uniform(0.3, 3.0, size=(n_cases, 1))          # the reference never normalises the view direction
    colour = rng.uniform(0.05, 1.0, size=(n_cases, 3))
    rough = rng.uniform(0.05, 0.9, size=n_cases)
    got = _rough_diffuse_fp32(n.astype(F), ld.astype(F), vd.astype(F), colour.astype(F), rough.astype(F)).astype(np.float64)
    ref = np.array([orc.rough_diffuse(n[i], ld[i], vd[i], colour[i], rough[i]) for i in range(n_cases)])
    assert np.isfinite(got).all()
    cr = np.einsum("ij,ij->i", n, _unit(-vd))
    cl = np.einsum("ij,ij->i", n, _unit(-ld))
    cb = np.maximum(cr, cl)
    well = np.abs(cb) > 0.05  # beta at least 3 degrees off grazing
    err = np.abs(got - ref).max(axis=1)
    print("well conditioned: %d of %d, max abs error %.3g; all: median %.3g" % (well.sum(), n_cases, err[well].max(), np.median(err)))
    assert well.mean() > 0.9
    assert err[well].max() < 2e-5          # measured 2e-6: amplification <= 1 / 0.05^2 = 400 on ~1e-7 of input rounding, times the colour
    assert np.median(err) < 2e-6
