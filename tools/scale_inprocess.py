#!/usr/bin/env python
"""Every BASELINE.json config at 1 / 2 / 4 / 8 GPUs through the call the CLI and the F# shim make: ONE process,
ftb_render(n_gpus = N) into a pageable RGBA8 buffer (tiles dealt to the devices, devices 1..N-1 store their finished
tiles into device 0's memory over NVLink, device 0 assembles and downloads band by band behind the rendering).
No torch, no torchrun: the whole sweep is one short process, so it fits one multi-GPU gpurun call.

Per (config, N): 2 warm-up frames, K timed ones (wall clock around the synchronous call = end to end, download included),
the frame compared byte for byte with the N = 1 frame; rays from one counting launch at N = 1.
usage: python tools/scale_inprocess.py [--steps K] [--gpus 1,2,4,8] [workload ...]  > gpurun_out/scale.jsonl"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    from functracer_b200 import abi, api, frontend, scenes
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("workloads", nargs="*")
    args = ap.parse_args()
    have = api.device_count()
    ns = [n for n in (int(x) for x in args.gpus.split(",")) if n <= have]
    names = args.workloads or ["cfg1-sample", "cfg2-hollow-sphere", "cfg3-house", "cfg3-night-house", "cfg4-bunny", "cfg4-bunny-d12",
                               "cfg4-bunny-full-d14", "cfg5-moon", "cfg5-repeat"]
    rows = []
    for name in names:
        cfg = scenes.CONFIGS[name]
        sc = frontend.ParsedScene(scenes.config_text(name), scenes.asset_dir())
        W, H, spp = sc.width, sc.height, sc.spp
        jit = frontend.jitter_pattern(cfg["seed"], spp)
        with api.Scene(sc) as scene:
            st = scene.render_params(api.make_params(W, H, spp, jit, seed=1234, out_format=abi.OUT_RGBA8, collect_stats=1), stats=True)["stats"]
            rays = float(st.primary_rays + st.shadow_rays + st.reflection_rays)
            first = None
            t1 = None
            for n in ns:
                p = api.make_params(W, H, spp, jit, seed=1234, out_format=abi.OUT_RGBA8, n_gpus=n)
                out = np.empty((H, W, 4), dtype=np.uint8)
                try:
                    for _ in range(2):
                        scene.render_params(p, out=out)
                    t0 = time.perf_counter()
                    for _ in range(args.steps):
                        scene.render_params(p, out=out)
                    ms = 1e3 * (time.perf_counter() - t0) / args.steps
                    kst = scene.render_params(p, out=out, stats=True)["stats"]  # one more frame with the library's own kernel timing (unbanded)
                except api.FtbError as e:
                    rows.append(dict(workload=name, n_gpus=n, error=str(e)))
                    print(json.dumps(rows[-1]), flush=True)
                    continue
                if first is None:
                    first, t1 = out.copy(), ms
                row = dict(workload=name, width=W, height=H, spp=spp, n_gpus=n, e2e_ms=ms, mrays_s=rays / ms / 1e3, kernel_ms_max_over_devices=kst.kernel_ms,
                           speedup=t1 / ms, efficiency=t1 / ms / n, frame="bit-exact vs 1 GPU" if bool((out == first).all()) else "MISMATCH vs 1 GPU",
                           d2h_bytes=int(out.nbytes), steps=args.steps)
                rows.append(row)
                print(json.dumps(row), flush=True)
    print("\n| config | W×H×spp | " + " | ".join("%d GPU%s: ms (Mrays/s, eff.)" % (n, "" if n == 1 else "s") for n in ns) + " | frames |", file=sys.stderr)
    print("|---|---|" + "---|" * (len(ns) + 1), file=sys.stderr)
    for name in names:
        rs = [r for r in rows if r["workload"] == name and "error" not in r]
        if not rs:
            continue
        cells = ["%.3f (%.0f, %.2f)" % (r["e2e_ms"], r["mrays_s"], r["efficiency"]) for r in rs]
        ok = all(r["frame"].startswith("bit-exact") for r in rs)
        print("| %s | %d×%d×%d | %s | %s |" % (name, rs[0]["width"], rs[0]["height"], rs[0]["spp"], " | ".join(cells), "bit-exact" if ok else "MISMATCH"), file=sys.stderr)


if __name__ == "__main__":
    main()
