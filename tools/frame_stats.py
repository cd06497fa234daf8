#!/usr/bin/env python
"""Work counters of one frame (counting kernel): where the algorithmic work goes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from functracer_b200 import abi, api, frontend, scenes
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2-hollow-sphere"
cfg = scenes.CONFIGS[name]
sc = frontend.ParsedScene(scenes.config_text(name), scenes.asset_dir())
jit = frontend.jitter_pattern(cfg["seed"], sc.spp)
with api.Scene(sc) as scene:
    r = scene.render(sc.width, sc.height, sc.spp, jit, stats=True, out_format=abi.OUT_RGBA8)
d = r["stats"].as_dict()
rays = d["primary_rays"] + d["shadow_rays"] + d["reflection_rays"]
d["per_ray"] = {"bound_tests": d["bound_tests"] / rays, "csg_ops": d["csg_ops"] / rays, "leaf_tests": sum(d["leaf_tests"]) / rays, "bvh_nodes": d["bsp_nodes_visited"] / rays}
print(json.dumps(d))
