#!/usr/bin/env python
"""A/B of the two formulations of the render loop on one GPU: the persistent megakernel (default) and the wavefront kernels
(FTB_WAVEFRONT=1, csrc/cuda/wavefront.cuh).  Each arm runs in its own process (the switch is read once); the frames must be
bit-identical, the timings are CUDA events over device-resident frames with the L2 flushed in between.
usage: python tools/wavefront_ab.py [workload ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def arm(names):
    import numpy as np
    import torch
    from functracer_b200 import abi, api, frontend, scenes
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for name in names:
        cfg = scenes.CONFIGS[name]
        sc = frontend.ParsedScene(scenes.config_text(name), scenes.asset_dir())
        W, H, spp = sc.width, sc.height, sc.spp
        jit = frontend.jitter_pattern(cfg["seed"], spp)
        with api.Scene(sc) as scene:
            p = api.make_params(W, H, spp, jit, seed=1234, out_format=abi.OUT_RGB_F32)
            tiles = torch.zeros(api.tile_buffer_bytes(p), dtype=torch.uint8, device="cuda")
            frame = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
            ms = []
            for it in range(6):
                flush.zero_()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                scene.render_tiles_device(p, tiles.data_ptr(), stream=stream)
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    ms.append(e0.elapsed_time(e1))
            api.assemble_device(p, [tiles.data_ptr()], frame.data_ptr(), stream=stream)
            torch.cuda.synchronize()
            scene.check_overflow(stream=stream)
            f = frame.cpu().numpy()
            import hashlib
            path = os.path.join(os.environ.get("FTB_AB_DIR", "/tmp"), "wfab_%s_%s.npy" % (name, "wf" if os.environ.get("FTB_WAVEFRONT") else "mk"))
            np.save(path, f)
            out[name] = dict(ms=sum(ms) / len(ms), sha=hashlib.sha256(f.tobytes()).hexdigest(), path=path)
    print("ARM " + json.dumps(out), flush=True)


if __name__ == "__main__":
    if os.environ.get("FTB_AB_ARM"):
        arm(sys.argv[1:])
        sys.exit(0)
    names = sys.argv[1:] or ["cfg3-house", "cfg3-night-house", "cfg5-repeat", "cfg2-hollow-sphere", "cfg5-moon"]
    res = {}
    for label, env in (("megakernel", {}), ("wavefront", {"FTB_WAVEFRONT": "1"})):
        e = dict(os.environ, FTB_AB_ARM="1", **env)
        e.pop("FTB_WAVEFRONT", None) if not env else None
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + names, env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        line = [l for l in r.stdout.splitlines() if l.startswith("ARM ")]
        if not line:
            print(label, "FAILED", r.stderr[-2000:])
            continue
        res[label] = json.loads(line[-1][4:])
    import numpy as np
    for n in names:
        a, b = res.get("megakernel", {}).get(n), res.get("wavefront", {}).get(n)
        if a and b:
            fa, fb = np.load(a["path"]), np.load(b["path"])
            d = np.abs(fa.astype(np.float64) - fb)
            same = "bit-identical" if a["sha"] == b["sha"] else "max |diff| %.3g, %.4f %% of pixels differ at all, %.6f %% by more than 1/255" % (
                np.nanmax(d), 100.0 * float((d.max(axis=-1) > 0).mean()), 100.0 * float((d.max(axis=-1) > 1 / 255.0).mean()))
            print("%-22s megakernel %9.3f ms   wavefront %9.3f ms   (%+.1f %%)   frames: %s" % (n, a["ms"], b["ms"], 100.0 * (b["ms"] / a["ms"] - 1.0), same))
            os.remove(a["path"]); os.remove(b["path"])
