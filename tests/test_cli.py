"""The CLI drop-in (functracer_b200/ftb-render = Program.fs with the render loop replaced): argument / exit-code
contract on CPU, a rendered PNG against the oracle on GPU."""
import os
import subprocess
import zlib

import numpy as np
import pytest

from functracer_b200 import frontend, scenes
from oracle import ftb_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "functracer_b200", "ftb-render")


def _read_png_rgba(data):
    """Minimal decoder for the 8-bit RGBA, non-interlaced PNGs the front end writes (filter types 0-4)."""
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        n = int.from_bytes(data[pos:pos + 4], "big")
        typ = data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        if typ == b"IHDR":
            w, h = int.from_bytes(body[:4], "big"), int.from_bytes(body[4:8], "big")
            assert body[8] == 8 and body[9] == 6 and body[12] == 0
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    raw = zlib.decompress(idat)
    out = np.zeros((h, w * 4), dtype=np.uint8)
    stride = w * 4
    for y in range(h):
        f = raw[y * (stride + 1)]
        line = np.frombuffer(raw[y * (stride + 1) + 1:(y + 1) * (stride + 1)], dtype=np.uint8).astype(np.int32)
        prev = out[y - 1].astype(np.int32) if y else np.zeros(stride, dtype=np.int32)
        cur = np.zeros(stride, dtype=np.int32)
        for x in range(stride):
            a = cur[x - 4] if x >= 4 else 0
            b = prev[x]
            c = prev[x - 4] if x >= 4 else 0
            if f == 0: p = 0
            elif f == 1: p = a
            elif f == 2: p = b
            elif f == 3: p = (a + b) // 2
            else:
                pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            cur[x] = (line[x] + p) & 255
        out[y] = cur
    return out.reshape(h, w, 4)


def test_cli_parse_error_exits_1_with_message(tmp_path):
    """Program.readScene (Program.fs:10-16): message on stdout, exit code 1."""
    bad = tmp_path / "bad.scene"
    bad.write_text("camera pos (0,0,0) lookat (0,0,1) up (0,1,0) fov 60 ratio 1\n\nspheer\n")
    r = subprocess.run([CLI, str(bad), str(tmp_path / "o.png")], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.strip() != ""
    assert "Using input file" in r.stderr


@pytest.mark.gpu
def test_cli_renders_the_same_image_as_the_oracle(tmp_path):
    text = scenes.hollow_sphere(res=(96, 54), spp=2)
    scene = tmp_path / "hs.scene"
    scene.write_text(text)
    out = tmp_path / "o.png"
    env = dict(os.environ, FTB_SEED="5", FTB_ASSETS=scenes.asset_dir())
    r = subprocess.run([CLI, str(scene), str(out)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    for phase in ("Parsed input", "Generated rays", "Geometry created", "Shaded scene", "Writing output", "Elapsed Time"):
        assert phase in r.stderr  # the stderr phase stamps of Program.runTracer (Program.fs:53-67)
    img = _read_png_rgba(out.read_bytes())
    sc = frontend.ParsedScene(text, scenes.asset_dir())
    jit = frontend.jitter_pattern(5, 2)
    ref = orc.quantise(orc.render(sc, orc.make_params(96, 54, 2, jit, seed=5))["rgb"])
    assert img.shape == (54, 96, 4) and (img[..., 3] == 255).all()
    d = np.abs(img[..., :3].astype(int) - ref.astype(int)).max(axis=-1)
    assert float((d <= 1).mean()) >= 0.999
    # stdout mode: two-argument form only writes a file; one argument streams the PNG to stdout
    r2 = subprocess.run([CLI, str(scene)], capture_output=True, env=env)
    assert r2.returncode == 0 and r2.stdout[:8] == b"\x89PNG\r\n\x1a\n"
