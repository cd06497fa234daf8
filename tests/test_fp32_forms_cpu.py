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


def _roots64(kind, o, d):
    """Exact-enough reference: the canonical quadric x^2 + wy y^2 + z^2 = k along o + t d in float64, far ("+") root first like
    Math.quadratic (Math.fs:4-10)."""
    wy = {0: 1.0, 1: 0.0, 2: -1.0}[kind]
    k = 0.0 if kind == 2 else 1.0
    a = d[:, 0] ** 2 + d[:, 2] ** 2 + wy * d[:, 1] ** 2
    b = 2 * (o[:, 0] * d[:, 0] + o[:, 2] * d[:, 2] + wy * o[:, 1] * d[:, 1])
    c = o[:, 0] ** 2 + o[:, 2] ** 2 + wy * o[:, 1] ** 2 - k
    disc = b * b - 4 * a * c
    sq = np.sqrt(np.maximum(disc, 0))
    return disc, (-b + sq) / (2 * a), (-b - sq) / (2 * a)


def _roots32(kind, o, d, recentre):
    """render.cuh quadricRoots in float32: the literal a, b, c (recentre = False) or the product build's re-centred form
    (origin slid to the point of closest approach to the model origin, t = t' + ts)."""
    wy = F({0: 1.0, 1: 0.0, 2: -1.0}[kind])
    k = F(0.0 if kind == 2 else 1.0)
    ts = np.zeros(o.shape[0], dtype=F)
    if recentre:
        dd = _dot32(d, d)
        ts = (-_dot32(o, d) / dd).astype(F)
        o = (o + ts[:, None] * d).astype(F)
    a = (d[:, 0] * d[:, 0] + d[:, 2] * d[:, 2] + wy * d[:, 1] * d[:, 1]).astype(F)
    b = (F(2) * (o[:, 0] * d[:, 0] + o[:, 2] * d[:, 2] + wy * o[:, 1] * d[:, 1]).astype(F)).astype(F)
    c = ((o[:, 0] * o[:, 0] + o[:, 2] * o[:, 2] + wy * o[:, 1] * o[:, 1]).astype(F) - k).astype(F)
    disc = (b * b - F(4) * a * c).astype(F)
    sq = np.sqrt(np.maximum(disc, F(0))).astype(F)
    two_a = (F(2) * a).astype(F)
    return disc, ((-b + sq) / two_a + ts).astype(F), ((-b - sq) / two_a + ts).astype(F)


def test_recentred_quadric_roots_keep_fp32_accurate_far_from_the_surface():
    """Why the FP32 kernels do not evaluate Math.quadratic on the literal coefficients: with the origin far
    from the surface b^2 and 4ac agree in their leading digits and float32 loses the discriminant - the error in t exceeds the
    reference's fixed 1e-4 shadow-ray offset (Shading.fs:111).  The re-centred form has the same real roots and keeps t to a few
    ulp of its own size at every distance."""
    rng = np.random.default_rng(9)
    n = 100_000
    for kind in (0, 1):  # sphere, cylinder (the cone's apex form has k = 0 and no cancellation of this kind)
        dist = 10.0 ** rng.uniform(0.3, 3, size=n)  # 2 .. 1000 model units from the axis / centre
        o = _unit(rng.normal(size=(n, 3))) * dist[:, None]
        if kind == 1:
            o[:, 1] = rng.uniform(-0.5, 1.5, size=n)
        aim = rng.uniform(-0.6, 0.6, size=(n, 3))  # through the body of the unit quadric
        d = _unit(aim - o) * rng.uniform(0.3, 3.0, size=(n, 1))
        o32, d32 = o.astype(F), d.astype(F)
        disc, t0, t1 = _roots64(kind, o32.astype(np.float64), d32.astype(np.float64))
        hit = disc > 1e-3 * (d32.astype(np.float64) ** 2).sum(axis=1)  # clear of the silhouette, where the count itself is ill conditioned
        assert hit.mean() > 0.5
        _, r0, r1 = _roots32(kind, o32, d32, recentre=True)
        _, l0, l1 = _roots32(kind, o32, d32, recentre=False)
        scale = np.maximum(1.0, np.abs(t1))
        err_re = np.maximum(np.abs(r0 - t0), np.abs(r1 - t1))[hit] / scale[hit]
        err_lit = np.maximum(np.abs(l0 - t0), np.abs(l1 - t1))[hit] / scale[hit]
        far = (dist > 30)[hit]
        print("kind %d: re-centred max rel error %.2e; literal float32 max %.2e (origins beyond 30 units: %.2e)" %
              (kind, err_re.max(), np.nanmax(err_lit), np.nanmax(err_lit[far])))
        assert err_re.max() < 1e-5  # a few ulp of t (the kernel forms oc with FMAs, this replay with separate roundings)
        assert np.nanmax(err_lit[far]) > 20 * err_re.max()  # what the literal form would cost
