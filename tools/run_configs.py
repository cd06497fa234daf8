#!/usr/bin/env python
"""Renders every BASELINE.json config at full size on one GPU (device-resident timing, CUDA events),
counts rays with the STATS kernel, and checks a window of each frame against the CPU oracle.
Prints one JSON line per config; `python tools/run_configs.py [name ...] > gpurun_out/configs.jsonl`."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from functracer_b200 import abi, api, frontend, scenes
from oracle import ftb_oracle as orc

SEED = 1234


def run(name, steps=3, kw=None, check=True):
    cfg = scenes.CONFIGS[name.split("@")[0]]
    text = scenes.config_text(name.split("@")[0], **(kw or {}))
    parsed = frontend.ParsedScene(text, scenes.asset_dir())
    W, H, spp = parsed.width, parsed.height, parsed.spp
    jit = frontend.jitter_pattern(cfg["seed"], spp)
    stream = torch.cuda.current_stream().cuda_stream
    with api.Scene(parsed) as scene:
        p = api.make_params(W, H, spp, jit, seed=SEED, out_format=abi.OUT_RGB_F32)
        tiles = torch.empty(api.tile_buffer_bytes(p), dtype=torch.uint8, device="cuda")
        frame = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
        ps = api.make_params(W, H, spp, jit, seed=SEED, out_format=abi.OUT_RGB_F32, collect_stats=1)
        st = scene.render_tiles_device(ps, tiles.data_ptr(), stream=stream, stats=True)
        rays = st.primary_rays + st.shadow_rays + st.reflection_rays
        for _ in range(2):
            scene.render_tiles_device(p, tiles.data_ptr(), stream=stream)
        torch.cuda.synchronize()
        ms = []
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            scene.render_tiles_device(p, tiles.data_ptr(), stream=stream)
            api.assemble_device(p, [tiles.data_ptr()], frame.data_ptr(), stream=stream)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        best = min(ms)
        out = {"config": name, "width": W, "height": H, "spp": spp, "ms_per_frame": best, "mrays_per_s": rays / best / 1e3,
               "primary": st.primary_rays, "shadow": st.shadow_rays, "reflection": st.reflection_rays,
               "algorithmic_gflop": st.flops / 1e9, "fp32_tflops": st.flops / best / 1e9, "stats_kernel_ms": st.kernel_ms}
        if check:
            # a 96 x 64 window around the brightest pixel of the frame against the oracle (same jitter, same RNG seed)
            lum = frame.sum(dim=-1)
            idx = int(torch.argmax(lum).item())
            x0, y0 = min(max(0, idx % W - 48), max(0, W - 96)), min(max(0, idx // W - 32), max(0, H - 64))
            x1, y1 = min(W, x0 + 96), min(H, y0 + 64)
            po = orc.make_params(W, H, spp, jit, seed=SEED)
            t0 = time.perf_counter()
            ref = orc.render(parsed, po, window=(x0, y0, x1, y1), debug=False)
            cpu_s = time.perf_counter() - t0
            got = frame[y0:y1, x0:x1].cpu().numpy().astype(np.float64)
            d = np.abs(got - ref["rgb"][y0:y1, x0:x1]).max(axis=-1)
            rr = ref["stats"]
            out.update({"window": [x0, y0, x1, y1], "window_frac_within_1_255": float((d <= 1 / 255).mean()), "window_max_err": float(d.max()),
                        "cpu_oracle_mrays_per_s": (rr.primary_rays + rr.shadow_rays + rr.reflection_rays) / cpu_s / 1e6, "cpu_threads": os.cpu_count()})
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or ["cfg1-sample", "cfg2-hollow-sphere", "cfg3-house", "cfg3-night-house", "cfg4-bunny", "cfg4-bunny-d12", "cfg4-bunny-full-d14", "cfg5-moon", "cfg5-repeat"]
    for n in names:
        kw = None
        try:
            run(n, kw=kw)
        except Exception as e:  # keep going: one failing config must not hide the others
            print(json.dumps({"config": n, "error": str(e)}), flush=True)
