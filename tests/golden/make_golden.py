#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/: small frames of every bundled scene rendered by the CPU
oracle (oracle/ftb_oracle.cpp, the line-by-line restatement of the F# render loop), with the primary hit maps.

These are NOT outputs of the F# reference (which cannot run in this image: no .NET); they pin the oracle and
the CUDA path against silent drift.  Re-generate with:  python tests/golden/make_golden.py
Stored per scene: rgb (float64, H x W x 3), prim / sub (int32, per sample), jitter (the host-drawn pattern),
and the scene text itself, so the fixture is self-contained."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from functracer_b200 import frontend, scenes  # noqa: E402
from oracle import ftb_oracle as orc  # noqa: E402

CASES = {
    "sample": lambda: scenes.sample(res=(64, 48), spp=2),
    "hollow_sphere": lambda: scenes.hollow_sphere(res=(64, 36), spp=2),
    "house": lambda: scenes.house(res=(64, 36), spp=2),
    "night_house": lambda: scenes.night_house(res=(64, 36), spp=2),
    "repeat": lambda: scenes.repeat(res=(64, 36), spp=2),
    "moon": lambda: scenes.moon(res=(48, 48), spp=2),
    "bunny_d0": lambda: scenes.bunny(res=(48, 27), spp=1, depth=0, mesh="bunny_tiny.ply"),
    "bunny_d3": lambda: scenes.bunny(res=(48, 27), spp=2, depth=3, mesh="bunny_tiny.ply"),
}
RNG_SEED = 1234
JITTER_SEED = 9


def main():
    for name, build in CASES.items():
        text = build()
        sc = frontend.ParsedScene(text, scenes.asset_dir())
        jit = frontend.jitter_pattern(JITTER_SEED, sc.spp)
        r = orc.render(sc, orc.make_params(sc.width, sc.height, sc.spp, jit, seed=RNG_SEED))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rgb=r["rgb"], prim=r["prim"], sub=r["sub"], jitter=jit,
                            text=np.array(text), width=sc.width, height=sc.height, spp=sc.spp, rng_seed=RNG_SEED)
        print(name, r["rgb"].shape, "mean", float(np.nanmean(r["rgb"])))


if __name__ == "__main__":
    main()
