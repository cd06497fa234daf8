#!/usr/bin/env python
"""bench.py — the headline benchmark of functracer_b200.

Metric (BASELINE.json): Mrays/s, primary + secondary (shadow + unique reflection rays), for one
frame of a bundled scene.  A "step" is one frame: generateRays -> shade -> blendPixels
(Program.fs:54-64).  Default workload = BASELINE.json configs[1]: hollow-sphere.scene at
1920x1080, 4 jittered spp (the deterministic CSG + reflection-depth-8 scene).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL): the frame's 16x16 tiles are dealt
round-robin to the ranks, each rank renders its tiles from its own atomic tile queue, and the
tile buffers are gathered to rank 0 over NVLink (NCCL gather) where the frame is assembled.
Total work is fixed as N grows => "scaling": "strong".

`value`  : inputs resident (scene on the GPU), device-to-device: tiles -> (gather) -> assembled frame.
`e2e`    : the same frame through the host-buffer C-ABI call the F# shim makes (ftb_render at N = 1;
           ftb_render_tiles_device + gather + ftb_assemble_device + D2H at N > 1), host buffers
           pinned, H2D of the step's inputs and D2H of the frame inside the timed region.
`roofline`: the render kernel against the FP32 pipe (this path is not HBM- or tensor-bound:
           the scene is KBs and rays live in registers).
`cpu_baseline`: the CPU oracle (a C++ port of the reference's algorithm; the F# original cannot
           run in this image) on all host cores, on a bounded sample of the same frame.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DEFAULT_WORKLOAD = "cfg2-hollow-sphere"
METRIC = "Mrays/sec (primary+secondary)"
UNIT = "Mrays/s"
RNG_SEED = 1234


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# per-ray intersection cost of the reference's brute-force algorithm on the bundled scenes (flop table)
_BRUTE_FORCE_FLOPS_PER_RAY = {"sample": 366, "hollow_sphere": 2972, "house": 610, "night_house": 1285, "repeat": 1214, "moon": 244,
                              "bunny": 960 * 45 + 33}


def _fp32_peak_tflops():
    """FP32 pipe peak.  Prefers the on-box FMA-saturation measurement committed under profiles/
    (tools/fp32_peak.cu); else 148 SM x 128 lanes x 2 x clocks.max.sm from MEASURED_PEAKS.json."""
    p = os.path.join(ROOT, "profiles", "fp32_peak.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["fp32_tflops"]), "measured FMA-saturation microbenchmark (profiles/fp32_peak.json)"
        except Exception:
            pass
    pk, how = _peaks()
    mhz = float(pk.get("sm_max_mhz", 1965.0))
    return 148 * 128 * 2 * mhz * 1e6 / 1e12, "derived 148 SM x 128 lanes x 2 x %g MHz (%s MEASURED_PEAKS.json sm_max_mhz)" % (mhz, how)


def _traffic_per_launch():
    p = os.path.join(ROOT, "profiles", "render_kernel_dram.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:
            pass
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _workload(name):
    from functracer_b200 import frontend, scenes
    cfg = scenes.CONFIGS[name]
    text = scenes.config_text(name)
    parsed = frontend.ParsedScene(text, scenes.asset_dir())
    jit = frontend.jitter_pattern(cfg["seed"], parsed.spp)
    return cfg, parsed, jit


def _sample_windows(W, H, n_stripes, stripe_w):
    """Evenly spaced full-height vertical stripes: a bounded, representative sample of the frame."""
    xs = [int((k + 0.5) * W / n_stripes - stripe_w / 2) for k in range(n_stripes)]
    return [(max(0, x), 0, min(W, max(0, x) + stripe_w), H) for x in xs]


def cpu_sample(parsed, jit, windows, threads=0):
    """Times the CPU oracle on the given windows of the frame.  Returns (rays, seconds, cores)."""
    from oracle import ftb_oracle as orc
    import numpy as np
    p = orc.make_params(parsed.width, parsed.height, parsed.spp, jit, seed=RNG_SEED)
    L = orc.lib()
    rgb = np.zeros((parsed.height, parsed.width, 3))
    from functracer_b200 import abi
    rays, secs = 0, 0.0
    for (x0, y0, x1, y1) in windows:
        st = abi.Stats()
        t0 = time.perf_counter()
        rc = L.ftbo_render_window(parsed.desc_ptr, parsed.camera_ptr, C.byref(p), x0, y0, x1, y1,
                                  rgb.ctypes.data_as(C.POINTER(C.c_double)), None, C.byref(st), threads)
        secs += time.perf_counter() - t0
        assert rc == 0
        rays += st.primary_rays + st.shadow_rays + st.reflection_rays
    cores = threads if threads > 0 else (os.cpu_count() or 1)
    return rays, secs, cores


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the F# binary cannot run in this
    image: no dotnet) on all host cores, each step = the bounded sample of the same frame."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, parsed, jit = _workload(args.workload)
    W, H = parsed.width, parsed.height
    windows = _sample_windows(W, H, 8, 8)
    for _ in range(args.warmup):
        cpu_sample(parsed, jit, windows[:1])
    rays = secs = 0
    for _ in range(args.steps):
        r, s, cores = cpu_sample(parsed, jit, windows)
        rays += r; secs += s
    v = rays / secs / 1e6
    sample = "%d full-height stripes of %d px (%.2f%% of the %dx%d x %d spp frame) per step" % (len(windows), 8, 100.0 * len(windows) * 8 / W, W, H, parsed.spp)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": args.workload, "scene": cfg["build"].__name__, "width": W, "height": H, "spp": parsed.spp,
                                         "note": "CPU oracle (C++ port of the F# algorithm, -O2, all host threads); F# reference not runnable here (no dotnet)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from functracer_b200 import abi, api
    from functracer_b200 import dist as fdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: functracer_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream().cuda_stream

    cfg, parsed, jit = _workload(args.workload)
    W, H, spp = parsed.width, parsed.height, parsed.spp
    t_create = time.perf_counter()
    scene = api.Scene(parsed)
    create_ms = 1e3 * (time.perf_counter() - t_create)

    def params(**kw):
        return api.make_params(W, H, spp, jit, seed=RNG_SEED, shard_index=rank if world > 1 else 0, shard_count=world, **kw)

    # ---- ray accounting: one counting pass of this rank's shard (outside every timed region) ----------
    p_f32 = params(out_format=abi.OUT_RGB_F32)
    tile_bytes = api.tile_buffer_bytes(p_f32)
    tiles = torch.empty(tile_bytes, dtype=torch.uint8, device=dev)
    p_stats = params(out_format=abi.OUT_RGB_F32, collect_stats=1)
    st = scene.render_tiles_device(p_stats, tiles.data_ptr(), stream=stream, stats=True)
    counts = torch.tensor([st.primary_rays, st.shadow_rays, st.reflection_rays, st.flops], dtype=torch.float64, device=dev)
    local_flops = float(st.flops)
    if world > 1:
        dist.all_reduce(counts)
    n_primary, n_shadow, n_refl, flops_total = [float(x) for x in counts.tolist()]
    rays_per_frame = n_primary + n_shadow + n_refl

    # gather buffers (rank 0) + assembled frame
    sizes = []
    for k in range(world):
        pk = api.make_params(W, H, spp, jit, shard_index=k, shard_count=world, out_format=abi.OUT_RGB_F32)
        sizes.append(api.tile_buffer_bytes(pk))
    max_bytes = max(sizes)
    arena = None
    if world > 1:
        tiles = torch.empty(max_bytes, dtype=torch.uint8, device=dev)  # equal-sized for gather
        gather_list = [torch.empty(max_bytes, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
        if not os.environ.get("FTB_NO_P2P"):
            try:  # the render kernel writes its tiles straight into rank 0's memory over NVLink (functracer_b200/dist.py)
                arena = fdist.PeerArena(max_bytes)
            except Exception as e:  # no IPC / peer access: NCCL gather
                arena = None
                if rank == 0:
                    print("peer arena unavailable (%s): falling back to the NCCL gather" % e, file=sys.stderr)
    sync_flag = torch.zeros(1, dtype=torch.int32, device=dev)
    frame = torch.empty((H, W, 3), dtype=torch.float32, device=dev) if rank == 0 else None
    # L2 flush between timed iterations: 256 MB > the 126 MB L2
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def device_step(ev=None):
        if ev:
            ev[0].record()
        scene.render_tiles_device(p_f32, arena.ptr(rank) if arena else tiles.data_ptr(), stream=stream)
        if ev:
            ev[1].record()
        if world > 1 and arena:
            dist.all_reduce(sync_flag)  # the only collective: every rank's stores have landed before rank 0 assembles
            if rank == 0:
                api.assemble_device(p_f32, [arena.ptr(k) for k in range(world)], frame.data_ptr(), stream=stream)
        elif world > 1:
            fdist.gather_tiles(tiles, max_bytes, gather_list)
            if rank == 0:
                api.assemble_device(p_f32, [g.data_ptr() for g in gather_list], frame.data_ptr(), stream=stream)
        else:
            api.assemble_device(p_f32, [tiles.data_ptr()], frame.data_ptr(), stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident ------------------------------------------------------------------------------
    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    step_ms, kern_ms = [], []
    for _ in range(args.steps):
        flush.zero_()  # evicts L2 (not timed)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        e0.record()
        device_step(kev)
        e1.record()
        barrier()
        step_ms.append(e0.elapsed_time(e1))
        kern_ms.append(kev[0].elapsed_time(kev[1]))
    t = torch.tensor([sum(step_ms), sum(kern_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_total_ms = [float(x) for x in t.tolist()]
    ms_per_step = total_ms / args.steps
    value = rays_per_frame / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host buffers through the C ABI --------------------------------------------------------------------
    out_host = torch.empty((H, W, 3), dtype=torch.float64).pin_memory() if rank == 0 else None
    out_np = out_host.numpy() if rank == 0 else None
    p_host = api.make_params(W, H, spp, jit, seed=RNG_SEED, out_format=abi.OUT_RGB_F64)
    frame64 = torch.empty((H, W, 3), dtype=torch.float64, device=dev) if (rank == 0 and world > 1) else None
    p_f64out = params(out_format=abi.OUT_RGB_F64)

    def e2e_step():
        if world == 1:
            scene.render_params(p_host, out=out_np)  # H2D (jitter, frame constants) + kernels + D2H, synchronous
        else:
            scene.render_tiles_device(p_f32, arena.ptr(rank) if arena else tiles.data_ptr(), stream=stream)
            if arena:
                dist.all_reduce(sync_flag)
            else:
                fdist.gather_tiles(tiles, max_bytes, gather_list)
            if rank == 0:
                srcs = [arena.ptr(k) for k in range(world)] if arena else [g.data_ptr() for g in gather_list]
                api.assemble_device(p_f64out, srcs, frame64.data_ptr(), stream=stream)
                out_host.copy_(frame64, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = rays_per_frame * args.steps / e2e_s / 1e6
    clocks = sampler.stop() if sampler else None  # sampled over both timed regions

    frame_check = None
    if world > 1:  # outside every timed region: the sharded, peer-written frame equals rank 0's own unsharded render, bit for bit
        device_step()
        barrier()
        if rank == 0:
            p_one = api.make_params(W, H, spp, jit, seed=RNG_SEED, out_format=abi.OUT_RGB_F32)
            t1 = torch.empty(api.tile_buffer_bytes(p_one), dtype=torch.uint8, device=dev)
            f1 = torch.empty_like(frame)
            scene.render_tiles_device(p_one, t1.data_ptr(), stream=stream)
            api.assemble_device(p_one, [t1.data_ptr()], f1.data_ptr(), stream=stream)
            torch.cuda.synchronize()
            frame_check = "bit-exact vs the 1-GPU render" if bool((f1 == frame).all()) else "MISMATCH vs the 1-GPU render"
        barrier()
    if rank == 0:
        peak, peak_how = _fp32_peak_tflops()
        kern_avg_ms = kern_total_ms / args.steps
        achieved = local_flops / (kern_avg_ms * 1e-3) / 1e12 if world == 1 else (flops_total / world) / (kern_avg_ms * 1e-3) / 1e12
        pk, pk_how = _peaks()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "scene": cfg["build"].__name__, "width": W, "height": H, "spp": spp,
                       "recursion_limit": 8, "rays_per_frame": {"primary": n_primary, "shadow": n_shadow, "reflection": n_refl},
                       "tile": "16x16 round-robin over ranks", "l2": "flushed between timed iterations (256 MB write)",
                       "parallelism": "tiles%d" % world, "gather": ("p2p-stores" if arena else "nccl-gather") if world > 1 else "none"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "h2d_bytes_per_step": int(16 * spp + 512), "d2h_bytes_per_step": int(W * H * 24),
                    "call": "ftb_render (host RGB f64 frame)" if world == 1 else ("ftb_render_tiles_device (tiles stored into rank 0's memory over NVLink) + barrier + ftb_assemble_device + D2H" if arena else "ftb_render_tiles_device + NCCL gather + ftb_assemble_device + D2H"),
                    "scene_create_ms_first_call_incl_cuda_init": create_ms},
            "gpu_launches": int(args.steps * (2 if world == 1 else (2 if rank == 0 else 1))),
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": _traffic_per_launch(), "kernel": "ftb::render_kernel<float, FEAT, false> (the scene's feature-specialised variant)", "kernel_ms": kern_avg_ms,
                         "algorithmic_flops_per_launch": local_flops if world == 1 else flops_total / world, "peak_source": peak_how,
                         "hbm_note": "scene is KBs and rays never leave registers; algorithmic HBM bytes = framebuffer only (%d B/launch) vs %s %.0f GB/s"
                                     % (W * H * 12 // world, pk_how, pk.get("hbm_gbs", 0.0))},
            "clocks": clocks,
        }
        # What the same rays cost with the reference's own algorithm (every leaf tested for every ray):
        # not the roofline figure -- that counts the tests this kernel performs -- but the size of the algorithmic win.
        brute = _BRUTE_FORCE_FLOPS_PER_RAY.get(cfg["build"].__name__) if not cfg.get("kw") else None
        if brute and world == 1:
            line["roofline"]["reference_algorithm"] = {
                "intersection_flops_per_ray": brute, "source": "SURVEY.md 8(d): every leaf tested per ray, no culling",
                "equivalent_tflops": rays_per_frame * brute / (kern_avg_ms * 1e-3) / 1e12}
        if frame_check:
            line["frame_check"] = frame_check
        if not args.no_cpu_baseline:
            windows = _sample_windows(W, H, 8, 8)
            rays, secs, cores = cpu_sample(parsed, jit, windows)
            if secs < 8.0:  # scale the sample towards ~10 s of wall time on all cores, at most the whole frame
                n = int(min(W // 8, max(8, 8 * 10.0 / max(secs, 1e-3))))
                windows = _sample_windows(W, H, n, 8) if n < W // 8 else [(0, 0, W, H)]
                rays, secs, cores = cpu_sample(parsed, jit, windows)
            frac = sum((x1 - x0) * (y1 - y0) for x0, y0, x1, y1 in windows) / float(W * H)
            line["cpu_baseline"] = {"value": rays / secs / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d full-height stripe(s) = %.1f%% of the %dx%d x %d spp frame, %.1f s of wall time" % (len(windows), 100.0 * frac, W, H, spp, secs)}
        print(json.dumps(line))
    scene.close()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        if arena:
            arena.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
