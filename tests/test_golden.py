"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the CPU oracle).
CPU: the oracle still reproduces them (guards the oracle and the front end against drift).
GPU (-m gpu): the CUDA path reproduces them through the C ABI."""
import glob
import os

import numpy as np
import pytest

from functracer_b200 import abi, api, frontend, scenes
from oracle import ftb_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _load(path):
    z = np.load(path)
    sc = frontend.ParsedScene(str(z["text"]), scenes.asset_dir())
    return z, sc


def test_fixtures_exist():
    assert len(FIXTURES) >= 8


@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_reproduces_golden(path):
    z, sc = _load(path)
    assert (sc.width, sc.height, sc.spp) == (int(z["width"]), int(z["height"]), int(z["spp"]))
    r = orc.render(sc, orc.make_params(sc.width, sc.height, sc.spp, z["jitter"], seed=int(z["rng_seed"])))
    assert (r["prim"] == z["prim"]).all() and (r["sub"] == z["sub"]).all()
    assert np.allclose(r["rgb"], z["rgb"], rtol=1e-9, atol=1e-12, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_reproduces_golden(path):
    z, sc = _load(path)
    with api.Scene(sc) as scene:
        g64 = scene.render(sc.width, sc.height, sc.spp, z["jitter"], seed=int(z["rng_seed"]), precision=abi.PRECISION_FP64_VERIFY, debug=True)
        g32 = scene.render(sc.width, sc.height, sc.spp, z["jitter"], seed=int(z["rng_seed"]), precision=abi.PRECISION_FP32, debug=True)
    assert float((g64["prim"] != z["prim"]).mean()) <= 1e-3
    d64 = np.abs(g64["rgb"] - z["rgb"]).max(axis=-1)
    assert float((d64 <= 1e-6).mean()) >= 0.999
    d32 = np.abs(g32["rgb"] - z["rgb"]).max(axis=-1)
    assert float((d32 <= 1.0 / 255.0).mean()) >= 0.999
    assert float((g32["prim"] != z["prim"]).mean()) <= 1e-2
