"""Shared helpers for the parity tests."""
import numpy as np

from functracer_b200 import abi, frontend, scenes
from oracle import ftb_oracle as orc


def parse(text, assets=None):
    return frontend.ParsedScene(text, assets or scenes.asset_dir())


def one_object_scene(obj, lights="", camera="camera pos (0,0,-5) lookat (0,0,0) up (0,1,0) fov 60 ratio 1"):
    return camera + "\nsamples 1\n\n" + obj + "\n\n" + lights + ("\n" if lights else "")


def oracle_render(sc, width=None, height=None, spp=None, seed=1, rng_seed=1234, sampling=None, **kw):
    width = width or sc.width
    height = height or sc.height
    spp = spp or sc.spp
    sampling = sc.sampling if sampling is None else sampling
    jit = frontend.jitter_pattern(seed, spp)
    p = orc.make_params(width, height, spp, jit, sampling=sampling, seed=rng_seed, **kw)
    return orc.render(sc, p), p


def colour_stats(a, b):
    """Per-pixel max-channel error statistics in 1/255 units."""
    d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max(axis=-1)
    return dict(max=float(d.max()), frac_within=float((d <= 1.0 / 255.0).mean()))
