"""Parity bookkeeping shared by tests/ and bench.py's cpu_baseline leg (TEST INFRASTRUCTURE, like everything under
oracle/: the product package never imports it).

The bar is BASELINE.json's north_star: final colour within 1/255 per channel on >= 99.9 % of pixels with the largest
error stated, and primary hit / primitive-id maps equal except at grazing or tie cases.  A mismatching sample counts as
a grazing / silhouette case only if the kernel's answer is a primitive (or the background) that the ORACLE's own id map
shows within one pixel of that sample - i.e. the sample sits on the boundary between the two, where the last bits of t
decide (jitter offsets reach one pixel, Image.fs:101-110).  One more documented case exists for meshes (`prim_edge_leak`):
Moller-Trumbore in FP32 is not watertight - a ray within ~1e-7 (relative) of an edge shared by two triangles can be
rejected by both and leak through to whatever lies behind; such a sample is recognised by the oracle's sub-id plane showing
two different triangles of the same mesh within one pixel of it while the kernel reports the background.  Anything else -
an object missing, a wrong occluder - is `unexplained` and fails the tests.
"""
import numpy as np


def stripe_windows(W, H, n, width, margin=0):
    """n evenly spaced full-height stripes of `width` columns (+ margin columns of context on both sides)."""
    out = []
    for k in range(n):
        x = int((k + 0.5) * W / n - width / 2)
        x = min(max(x, margin), max(margin, W - width - margin))
        out.append((x - margin, 0, min(W, x + width + margin), H))
    return out


def sample_windows(W, H, spp, target_primary, width=8, min_stripes=64, margin=1):
    """The whole frame when it has at most target_primary primary samples, else evenly spaced full-height stripes adding
    up to about target_primary samples: min_stripes stripes of `width` px (each with `margin` px of context either side for
    the silhouette classification) if that fits the budget, else narrower (>= 2 px) and then fewer (>= 8) stripes.
    Returns (windows, margin)."""
    if W * H * spp <= target_primary:
        return [(0, 0, W, H)], 0
    cols = max(1, int(target_primary / float(H * spp)))
    n, w = max(1, min(min_stripes, W // (width + 2 * margin))), width
    while w > 2 and n * (w + 2 * margin) > cols:
        w -= 1
    if n * (w + 2 * margin) > cols:
        n = max(min(8, n), cols // (w + 2 * margin))
    return stripe_windows(W, H, n, w, margin), margin


def compare_window(ref_rgb, ref_prim, got_rgb, got_prim, margin=0, ref_sub=None, got_sub=None):
    """One window.  ref_* from the oracle, got_* from the kernel, same shapes: rgb [h, w, 3], prim / sub [h, w, spp].
    Columns within `margin` of the window's left / right edge are context only (not counted, but used as neighbours).
    Returns counts; see module docstring for `unexplained`."""
    h, w = ref_rgb.shape[:2]
    inner = slice(margin, w - margin) if margin else slice(0, w)
    d = np.abs(np.asarray(got_rgb, dtype=np.float64) - ref_rgb).max(axis=-1)[:, inner]
    finite = np.isfinite(d)
    out = dict(pixels=int(d.size), within=int((d[finite] <= 1.0 / 255.0).sum()), max_err=float(d[finite].max()) if finite.any() else 0.0,
               nonfinite=int((~finite).sum()))
    if ref_prim is not None and got_prim is not None:
        mism = (got_prim != ref_prim)
        mism[:, :margin] = False
        if margin:
            mism[:, w - margin:] = False
        out["samples"] = int(ref_prim[:, inner].size)
        out["prim_mismatch"] = int(mism.sum())
        unexplained = leaks = 0
        if mism.any():
            ys, xs, ss = np.nonzero(mism)
            for y, x, s in zip(ys.tolist(), xs.tolist(), ss.tolist()):
                nb = ref_prim[max(0, y - 1):y + 2, max(0, x - 1):x + 2]
                if (nb == got_prim[y, x, s]).any():
                    continue
                if ref_sub is not None and got_prim[y, x, s] == -1:  # leak through a shared edge of the oracle's mesh?
                    same = nb == ref_prim[y, x, s]
                    subs = ref_sub[max(0, y - 1):y + 2, max(0, x - 1):x + 2][same]
                    if np.unique(subs).size >= 2:
                        leaks += 1
                        continue
                unexplained += 1
                if len(out.setdefault("unexplained_samples", [])) < 8:
                    out["unexplained_samples"].append(dict(x=x, y=y, s=s, ref_prim=int(ref_prim[y, x, s]), got_prim=int(got_prim[y, x, s]),
                                                           ref_neighbourhood=sorted(set(int(v) for v in nb.ravel()))))
        out["prim_unexplained"] = unexplained
        out["prim_edge_leak"] = leaks
        if ref_sub is not None and got_sub is not None:
            same = (got_prim == ref_prim)
            sm = same & (got_sub != ref_sub)
            sm[:, :margin] = False
            if margin:
                sm[:, w - margin:] = False
            out["sub_mismatch"] = int(sm.sum())
    return out


def merge(results):
    tot = {}
    for r in results:
        for k, v in r.items():
            if k == "unexplained_samples":
                tot.setdefault(k, []).extend(v)
                continue
            tot[k] = max(tot.get(k, 0.0), v) if k == "max_err" else tot.get(k, 0) + v
    if tot.get("pixels"):
        tot["frac_within_1_255"] = tot["within"] / float(tot["pixels"])
    return tot
