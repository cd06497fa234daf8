"""The common-origin bound table of the render kernel (render.cuh) must never cull an item a ray can hit.

The kernel answers "can this ray touch item j at all?" for primary rays and point-light shadow rays with

    row.xyz . unit(d) + (slack - row.w)  >=  0,    row = (centre - origin, sqrt(|oc|^2 (1 - 8e-6) - w)),  w = inflated r^2

in FP32.  This test replays that arithmetic in numpy float32 (separate roundings instead of FMAs, the reciprocal square
root perturbed by +-2 ulp like MUFU.RSQ) on rays aimed at the silhouettes of randomly placed bounding spheres, over the
scales the scenes use and well beyond, and checks it against the exact geometry in float64: every ray whose line meets
the true sphere at an admissible distance must come out as a candidate.  It pins the rounding budget stated next to the
table (2e-4 |camera| for primary rays, 4e-4 tmax for shadow rays); no GPU needed.
"""
import numpy as np
import pytest

F = np.float32
N = 400_000


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _rand_dirs(rng, n):
    return _unit(rng.normal(size=(n, 3)))


def _perp(rng, u):
    p = np.cross(u, rng.normal(size=u.shape))
    return _unit(p)


def _inflated_r2(r):  # api.cu uploadScene: radius inflated by 0.2 % + 1e-5, squared, stored in the working precision
    ri = r * 1.002 + 1e-5
    return (ri * ri).astype(F)


def _row(centre32, origin32, w32, sign, rng=None):
    """The table row the kernel's prologue builds (float32); with rng the square root is off by up to 2 ulp like MUFU.SQRT."""
    v = (F(sign) * (centre32 - origin32)).astype(F)
    vv = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]).astype(F)
    k = (vv * F(1.0 - 8e-6) - w32).astype(F)
    s = np.where(~(k > 0), F(-np.inf), np.sqrt(np.maximum(k, F(0))).astype(F)).astype(F)  # w = +inf (unbounded item) gives k = -inf
    if rng is not None:
        s = (s * (F(1) + rng.uniform(-2.4e-7, 2.4e-7, size=s.shape).astype(F))).astype(F)
    return v, s


def _unit32(d32, rng):
    """unit(d) as traceScene forms it: d * (1 / sqrt(d.d)) with a 2-ulp reciprocal square root."""
    dd = (d32[:, 0] * d32[:, 0] + d32[:, 1] * d32[:, 1] + d32[:, 2] * d32[:, 2]).astype(F)
    inv = (F(1) / np.sqrt(dd).astype(F)).astype(F)
    inv = (inv * (F(1) + rng.uniform(-2.4e-7, 2.4e-7, size=inv.shape).astype(F))).astype(F)
    return (d32 * inv[:, None]).astype(F)


def _candidate(v, s, du, slack32):
    """traceScene's table loop: the sign of b - w with b - w = v.x du.x + (v.y du.y + (v.z du.z + (slack - w)))"""
    b = (slack32 - s).astype(F)
    b = (v[:, 2] * du[:, 2] + b).astype(F)
    b = (v[:, 1] * du[:, 1] + b).astype(F)
    b = (v[:, 0] * du[:, 0] + b).astype(F)
    return ~np.signbit(b) | np.isnan(b)


def _eps(rng, n):
    """relative miss distance of the aim point: 70 % at the silhouette (+-3e-3: inside, grazing, just past the inflated
    bound), 30 % clear misses by 5 % .. 300 % of the radius"""
    return np.where(rng.random(n) < 0.7, rng.uniform(-3e-3, 3e-3, size=n), 10.0 ** rng.uniform(np.log10(0.05), np.log10(3.0), size=n))


def _line_distance(o, d, c):
    """exact (float64): distance along the ray to the point nearest c, and the squared distance of c from the line"""
    u = _unit(d)
    oc = c - o
    b = np.einsum("ij,ij->i", oc, u)
    return b, np.einsum("ij,ij->i", oc, oc) - b * b


def _true_hit(o, d, c, r, tmin, tmax):
    """exact (float64) ray / sphere: a crossing with tmin <= t*|d| <= tmax (distances along the ray)"""
    b, perp2 = _line_distance(o, d, c)
    hit = perp2 <= r * r
    h = np.sqrt(np.maximum(r * r - perp2, 0.0))
    near, far = b - h, b + h
    return hit & (far >= tmin) & (near <= tmax)


def test_primary_rays_never_lose_a_hit():
    rng = np.random.default_rng(7)
    cam = (_rand_dirs(rng, N) * 10.0 ** rng.uniform(-1, 3, size=(N, 1))).astype(F)  # |camera| 0.1 .. 1000
    dist = 10.0 ** rng.uniform(-2, 4, size=N)                                        # camera-to-centre 0.01 .. 10 000
    r = dist * 10.0 ** rng.uniform(-3, -0.05, size=N)                                # outside the bound
    # a third of the cases: bounds so small that their inflation is of the order of the origin's rounding, seen grazing
    adv = rng.random(N) < 0.33
    g = 1e-7 * np.linalg.norm(cam.astype(np.float64), axis=1)
    r_adv = np.maximum((g * 10.0 ** rng.uniform(-1, 1, size=N) - 1e-5) / 0.002, 1e-6)
    r = np.where(adv & (r_adv < 0.5 * dist), r_adv, r)
    to_c = _rand_dirs(rng, N)
    centre = (cam.astype(np.float64) + to_c * dist[:, None]).astype(F)
    eps = _eps(rng, N)  # aim point: perpendicular offset r (1 + eps) from the centre
    eps = np.where(adv, rng.uniform(-2e-3, 2e-3, size=N), eps)
    aim = centre.astype(np.float64) + _perp(rng, to_c) * (r * (1 + eps))[:, None]
    d = (_unit(aim - cam.astype(np.float64)) * rng.uniform(0.5, 2.0, size=(N, 1))).astype(F)  # not normalised (Image.fs:83-89)
    o = (cam + F(1e-4) * d).astype(F)                                                # slightOffset (Shading.fs:129)
    w = _inflated_r2(r)
    v, s = _row(centre, cam, w, +1, rng)
    slack = (F(2e-4) * np.linalg.norm(cam.astype(np.float64), axis=1)).astype(F)
    cand = _candidate(v, s, _unit32(d, rng), slack)
    truth = _true_hit(o.astype(np.float64), d.astype(np.float64), centre.astype(np.float64), r * 1.001 + 5e-6, 0.0, np.inf)
    assert truth.mean() > 0.3  # the generator does produce hits
    lost = truth & ~cand
    assert not lost.any(), "%d of %d true hits culled" % (lost.sum(), truth.sum())
    # and it still culls: clear misses at the scales of the bundled scenes (camera within 30 of the origin, r >= 0.1, objects
    # no further than 50 radii: the 8e-6 |oc|^2 slack is 2 % of r^2 there and grows with the square of the distance)
    perp2 = _line_distance(o.astype(np.float64), d.astype(np.float64), centre.astype(np.float64))[1]
    clear = (perp2 > (1.2 * r) ** 2) & (np.linalg.norm(cam.astype(np.float64), axis=1) < 30) & (r > 0.1) & (dist < 50 * r)
    assert clear.sum() > 1000 and (~cand[clear]).mean() > 0.9, (~cand[clear]).mean()


def test_point_light_shadow_rays_never_lose_a_hit():
    rng = np.random.default_rng(11)
    light = (_rand_dirs(rng, N) * 10.0 ** rng.uniform(-1, 3, size=(N, 1))).astype(F)
    tmax_true = 10.0 ** rng.uniform(-2, 4, size=N)                                   # fragment-to-light 0.01 .. 10 000
    u = _rand_dirs(rng, N)
    frag = (light.astype(np.float64) - u * tmax_true[:, None]).astype(F)             # ray origin (f.p + 1e-4 n, any FP32 point)
    # the ray as the kernel forms it (shadowLightIntensity, Shading.fs:33-42)
    dvec = (light - frag).astype(F)
    tmax = np.sqrt((dvec[:, 0] * dvec[:, 0] + dvec[:, 1] * dvec[:, 1] + dvec[:, 2] * dvec[:, 2]).astype(F)).astype(F)
    d = _unit32(dvec, rng)
    # an occluder between fragment and light, often hugging the light (|oc| << tmax: the case the per-ray slack is for)
    a = tmax_true * 10.0 ** rng.uniform(-4, 0, size=N)                               # distance from the light along the ray
    r = a * 10.0 ** rng.uniform(-3, -0.05, size=N)                                   # the light is outside the bound
    eps = _eps(rng, N)
    # a third of the cases: bounds so small that their inflation is of the order of the direction's rounding, seen grazing
    adv = rng.random(N) < 0.33
    g = 1.7e-7 * tmax_true
    r_adv = np.maximum((g * 10.0 ** rng.uniform(-1, 1, size=N) - 1e-5) / 0.002, 1e-6)
    use = adv & (r_adv < 0.5 * a)
    r = np.where(use, r_adv, r)
    eps = np.where(use, rng.uniform(-2e-3, 2e-3, size=N), eps)
    ud = _unit(d.astype(np.float64))
    centre = (light.astype(np.float64) - ud * a[:, None] + _perp(rng, ud) * (r * (1 + eps))[:, None]).astype(F)
    w = _inflated_r2(r)
    v, s = _row(centre, light, w, -1, rng)
    slack = (F(4e-4) * tmax).astype(F)
    cand = _candidate(v, s, _unit32(d, rng), slack)
    truth = _true_hit(frag.astype(np.float64), d.astype(np.float64), centre.astype(np.float64), r * 1.001 + 5e-6, 0.0, tmax.astype(np.float64) * (1 + 1e-6))
    assert truth.mean() > 0.2
    lost = truth & ~cand
    assert not lost.any(), "%d of %d true hits culled" % (lost.sum(), truth.sum())
    perp2 = _line_distance(frag.astype(np.float64), d.astype(np.float64), centre.astype(np.float64))[1]
    # (occluders hugging the light are kept more often: the per-ray slack is a fraction of tmax, not of their distance)
    clear = (perp2 > (1.2 * r) ** 2) & (tmax_true < 100) & (r > 0.1) & (a < 50 * r) & (a > 0.1 * tmax_true)
    assert clear.sum() > 1000 and (~cand[clear]).mean() > 0.9, (~cand[clear]).mean()


def test_origin_inside_or_unbounded_is_always_a_candidate():
    rng = np.random.default_rng(3)
    n = 10_000
    cam = (_rand_dirs(rng, n) * 10.0).astype(F)
    r = rng.uniform(0.5, 5.0, size=n)
    centre = (cam.astype(np.float64) + _rand_dirs(rng, n) * (r * rng.uniform(0, 0.999, size=n))[:, None]).astype(F)  # camera inside
    v, s = _row(centre, cam, _inflated_r2(r), +1)
    assert np.isneginf(s).all()
    v, s = _row(centre, cam, np.full(n, np.inf, dtype=F), +1)  # unbounded items carry w = +inf
    assert np.isneginf(s).all()
    du = _unit32(_rand_dirs(rng, n).astype(F), rng)
    assert _candidate(v, s, du, np.zeros(n, dtype=F)).all()


# ---- the general bound loop (rays without a common origin: bounce rays, shadow rays of directional lights) -----------------------
def _general_candidate(centre32, w32, o32, du, packed):
    """traceScene's general loop in float32 with separate roundings: oc = c - o, b = oc . du, oc2 = oc . oc;
    scalar form:  miss = (w + 1e-6 oc2) - (oc2 - b b);   packed form (FP32 house-family kernels):  miss = b b + ((w + 1e-6 oc2) - oc2);
    outside = w - oc2;  candidate <=> !(miss < 0 or (b < 0 and outside < 0))"""
    oc = (centre32 - o32).astype(F)
    b = (oc[:, 0] * du[:, 0]).astype(F)
    b = (oc[:, 1] * du[:, 1] + b).astype(F)
    b = (oc[:, 2] * du[:, 2] + b).astype(F)
    oc2 = (oc[:, 0] * oc[:, 0]).astype(F)
    oc2 = (oc[:, 1] * oc[:, 1] + oc2).astype(F)
    oc2 = (oc[:, 2] * oc[:, 2] + oc2).astype(F)
    t1 = (w32 + (F(1e-6) * oc2).astype(F)).astype(F)
    bb = (b * b).astype(F)
    if packed:
        miss = (bb + (t1 - oc2).astype(F)).astype(F)
    else:
        miss = (t1 - (oc2 - bb).astype(F)).astype(F)
    outside = (w32 - oc2).astype(F)
    return ~((miss < 0) | ((b < 0) & (outside < 0)))


@pytest.mark.parametrize("packed", [False, True])
def test_general_bound_loop_never_loses_a_hit(packed):
    """Both associations of the line-miss test keep every ray that can touch the true bound at t >= 0, from origins at all scales,
    outside, inside and grazing; and both still cull clear misses and bounds behind the origin."""
    rng = np.random.default_rng(23 + int(packed))
    o = (_rand_dirs(rng, N) * 10.0 ** rng.uniform(-1, 3, size=(N, 1))).astype(F)
    dist = 10.0 ** rng.uniform(-2, 4, size=N)
    r = dist * 10.0 ** rng.uniform(-3, 0.3, size=N)  # up to 2 x the distance: the origin is inside some bounds
    to_c = _rand_dirs(rng, N)
    centre = (o.astype(np.float64) + to_c * dist[:, None]).astype(F)
    eps = _eps(rng, N)
    aim = centre.astype(np.float64) + _perp(rng, to_c) * (r * (1 + eps))[:, None]
    flip = np.where(rng.random(N) < 0.15, -1.0, 1.0)  # some rays point away from the bound
    d = (flip[:, None] * _unit(aim - o.astype(np.float64)) * rng.uniform(0.2, 3.0, size=(N, 1))).astype(F)
    w = _inflated_r2(r)
    cand = _general_candidate(centre, w, o, _unit32(d, rng), packed)
    truth = _true_hit(o.astype(np.float64), d.astype(np.float64), centre.astype(np.float64), r * 1.001 + 5e-6, 0.0, np.inf)
    assert truth.mean() > 0.3
    lost = truth & ~cand
    assert not lost.any(), "%d of %d true hits culled" % (lost.sum(), truth.sum())
    b, perp2 = _line_distance(o.astype(np.float64), d.astype(np.float64), centre.astype(np.float64))
    clear = (perp2 > (1.2 * r) ** 2) & (r > 0.1) & (dist < 50 * r)
    assert clear.sum() > 1000 and (~cand[clear]).mean() > 0.9, (~cand[clear]).mean()
    behind = (b < -1.01 * r * 1.002 - 1e-4) & (dist > 1.01 * r)
    assert behind.sum() > 1000 and (~cand[behind]).all()
    unbounded = _general_candidate(centre, np.full(N, np.inf, dtype=F), o, _unit32(d, rng), packed)
    assert unbounded.all()  # w = +inf: neither test can fire
