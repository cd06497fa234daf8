#!/usr/bin/env python
"""Runs fuzz seeds (tests/test_gpu_fuzz.py) and prints where the CUDA path and the oracle disagree."""
import importlib.util, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
spec = importlib.util.spec_from_file_location("fz", os.path.join(ROOT, "tests", "test_gpu_fuzz.py")); fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
from functracer_b200 import abi, api, frontend
from oracle import ftb_oracle as orc
from util import parse
for seed in [int(x) for x in sys.argv[1:]]:
    text = fz._scene(1000 + seed)
    print("=" * 30, seed); print(text)
    sc = parse(text)
    jit = frontend.jitter_pattern(seed + 1, sc.spp)
    ref = orc.render(sc, orc.make_params(sc.width, sc.height, sc.spp, jit, seed=77, sampling=sc.sampling))
    with api.Scene(sc) as scene:
        for prec in (abi.PRECISION_FP64_VERIFY, abi.PRECISION_FP32):
            try:
                g = scene.render(sc.width, sc.height, sc.spp, jit, seed=77, precision=prec, debug=True, sampling=sc.sampling)
            except api.FtbError as e:
                print("prec", prec, "ERROR", e); continue
            mm = np.nonzero(g["prim"] != ref["prim"])[0]
            d = np.abs(g["rgb"] - ref["rgb"]).max(axis=-1)
            bad = np.argwhere(~(d <= (1e-6 if prec else 1 / 255)) & np.isfinite(ref["rgb"]).all(axis=-1))
            print("prec", prec, "id mismatches", len(mm), "bad colour px", len(bad), "nan ref px", int((~np.isfinite(ref["rgb"]).all(axis=-1)).sum()), "nan gpu px", int((~np.isfinite(g["rgb"]).all(axis=-1)).sum()))
            for i in mm[:6]:
                print("   sample", i, "px", (i // sc.spp) % sc.width, (i // sc.spp) // sc.width, "ref prim/sub/t", ref["prim"][i], ref["sub"][i], ref["t"][i], "gpu", g["prim"][i], g["sub"][i], g["t"][i])
            for (y, x) in bad[:6]:
                print("   px", x, y, "ref", ref["rgb"][y, x], "gpu", g["rgb"][y, x], "prims", ref["prim"][(y * sc.width + x) * sc.spp:(y * sc.width + x + 1) * sc.spp], g["prim"][(y * sc.width + x) * sc.spp:(y * sc.width + x + 1) * sc.spp])
