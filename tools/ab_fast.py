#!/usr/bin/env python
"""Same-box A/B of builds of libfunctracer_b200.so without importing torch (a process costs ~1 s instead of ~10 s, so one
gpurun call can compare many builds on many workloads).  Every (workload, library) pair runs in its own process (FTB_LIB is
read once): ftb_render into a pageable RGBA8 buffer with a stats block, which makes the library time the render kernel
with its own CUDA events (ftb_stats.kernel_ms) - 2 warm-up frames, then the mean of K.  The RGBA8 frames of the builds are
compared byte for byte against the first library's.

usage: python tools/ab_fast.py "cfg5-repeat cfg3-house" "tree NAME ..." [K]     (NAME = ab/libftb_NAME.so, tools/ab_build.sh)"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def arm(name, k):
    import numpy as np
    from functracer_b200 import abi, api, frontend, scenes
    cfg = scenes.CONFIGS[name]
    sc = frontend.ParsedScene(scenes.config_text(name), scenes.asset_dir())
    W, H, spp = sc.width, sc.height, sc.spp
    jit = frontend.jitter_pattern(cfg["seed"], spp)
    out = np.empty((H, W, 4), dtype=np.uint8)
    with api.Scene(sc) as scene:
        p = api.make_params(W, H, spp, jit, seed=1234, out_format=abi.OUT_RGBA8)
        ms = []
        for it in range(2 + k):
            st = scene.render_params(p, out=out, stats=True)["stats"]
            if it >= 2:
                ms.append(st.kernel_ms)
    path = os.path.join(os.environ.get("FTB_AB_DIR", "/tmp"), "abfast_%s_%s.npy" % (name, os.environ.get("FTB_AB_TAG", "tree")))
    np.save(path, out)
    print("ARM " + json.dumps(dict(ms=sum(ms) / len(ms), lo=min(ms), sha=hashlib.sha256(out.tobytes()).hexdigest(), path=path)), flush=True)


if __name__ == "__main__":
    if os.environ.get("FTB_AB_ARM"):
        arm(sys.argv[1], int(sys.argv[2]))
        sys.exit(0)
    import numpy as np
    workloads = sys.argv[1].split()
    libs = sys.argv[2].split()
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    for w in workloads:
        base = None
        for lib in libs:
            name, _, extra = lib.partition("@")  # NAME@VAR=value[,VAR=value]: run-time switches of the library
            env = dict(os.environ, FTB_AB_ARM="1", FTB_AB_TAG=lib.replace("@", "_").replace("=", "_").replace(",", "_"))
            env.pop("FTB_LIB", None)
            if name != "tree":
                env["FTB_LIB"] = os.path.join(ROOT, "ab", "libftb_%s.so" % name)
            for kv in filter(None, extra.split(",")):
                env[kv.split("=")[0]] = kv.split("=")[1]
            r = subprocess.run([sys.executable, os.path.abspath(__file__), w, str(k)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
            line = [l for l in r.stdout.splitlines() if l.startswith("ARM ")]
            if not line:
                print("%-22s %-12s FAILED rc=%d %s" % (w, lib, r.returncode, r.stderr[-400:].replace("\n", " | ")), flush=True)
                continue
            d = json.loads(line[-1][4:])
            frame = np.load(d["path"])
            os.remove(d["path"])
            if base is None:
                base = (d["ms"], frame)
                same = "(base)"
            else:
                diff = frame.astype(np.int16) - base[1]
                same = "frame identical" if not diff.any() else "frame: %d of %d bytes differ, max |diff| %d" % (int((diff != 0).sum()), diff.size, int(np.abs(diff).max()))
            print("%-22s %-12s kernel %9.4f ms (min %9.4f)  %+6.2f %%   %s" % (w, lib, d["ms"], d["lo"], 100.0 * (d["ms"] / base[0] - 1.0), same), flush=True)
