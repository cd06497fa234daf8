#!/usr/bin/env python
"""Times shard 0 of K (K = 1, 2, 4, 8) of a workload on ONE GPU: what each rank of a K-GPU run executes.
Separates the kernel's own strong-scaling loss (ramp, tail, granularity) from the gather cost."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from functracer_b200 import abi, api, frontend, scenes

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2-hollow-sphere"
cfg = scenes.CONFIGS[name]
parsed = frontend.ParsedScene(scenes.config_text(name), scenes.asset_dir())
W, H, spp = parsed.width, parsed.height, parsed.spp
jit = frontend.jitter_pattern(cfg["seed"], spp)
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with api.Scene(parsed) as scene:
    for K in (1, 2, 4, 8):
        p = api.make_params(W, H, spp, jit, shard_index=0, shard_count=K, out_format=abi.OUT_RGB_F32)
        tiles = torch.empty(api.tile_buffer_bytes(p), dtype=torch.uint8, device="cuda")
        for _ in range(3):
            scene.render_tiles_device(p, tiles.data_ptr(), stream=stream)
        ms = []
        for _ in range(10):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            scene.render_tiles_device(p, tiles.data_ptr(), stream=stream)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        print(json.dumps({"workload": name, "shards": K, "ms_median": ms[len(ms) // 2], "ms_min": ms[0], "ideal_ms": None}), flush=True)
