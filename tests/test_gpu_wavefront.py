"""The wavefront arm of the render loop (csrc/cuda/wavefront.cuh, FTB_WAVEFRONT=1: one kernel per stage, rays and path state in
structure-of-arrays records in HBM) is the measured-and-retired alternative to the persistent megakernel (BENCH.md §4).  It
stays in the tree as an A/B arm, so it stays correct: the same scene through both arms gives the same primary hits and the
same frame up to the last bits (the stages are the same device functions, but the compiler contracts multiply-adds
differently in different kernels)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from functracer_b200 import abi, api, frontend, scenes
from util import parse

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from functracer_b200 import abi, api, frontend, scenes
from util import parse
sc = parse(getattr(scenes, %(scene)r)(res=(%(w)d, %(h)d), spp=%(spp)d))
jit = frontend.jitter_pattern(3, %(spp)d)
with api.Scene(sc) as scene:
    out = scene.render(%(w)d, %(h)d, %(spp)d, jit, seed=77, out_format=abi.OUT_RGB_F32, debug=True)
np.savez(%(path)r, rgb=out["rgb"], prim=out["prim"], sub=out["sub"])
"""


@pytest.mark.parametrize("scene,w,h,spp", [("house", 200, 120, 3), ("night_house", 160, 96, 2), ("hollow_sphere", 160, 96, 2), ("moon", 96, 96, 5)])
def test_wavefront_arm_matches_the_megakernel(tmp_path, scene, w, h, spp):
    sc = parse(getattr(scenes, scene)(res=(w, h), spp=spp))
    jit = frontend.jitter_pattern(3, spp)
    with api.Scene(sc) as s:
        mk = s.render(w, h, spp, jit, seed=77, out_format=abi.OUT_RGB_F32, debug=True)
    path = str(tmp_path / "wf.npz")
    env = dict(os.environ, FTB_WAVEFRONT="1")
    r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT, scene=scene, w=w, h=h, spp=spp, path=path)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    wf = np.load(path)
    assert (wf["prim"] == mk["prim"]).mean() >= 0.9999 and (wf["sub"] == mk["sub"]).mean() >= 0.9999
    d = np.abs(wf["rgb"].astype(np.float64) - mk["rgb"]).max(axis=-1)
    print("%s: wavefront vs megakernel max |diff| %.3g, pixels off by > 1e-4: %d" % (scene, d.max(), int((d > 1e-4).sum())))
    assert float((d <= 1e-4).mean()) >= 0.9995
