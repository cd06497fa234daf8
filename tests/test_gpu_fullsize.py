"""BASELINE.json's configs at their FULL sizes on the GPU, checked through properties that do not need a full CPU
render: (1) a window of the frame against the CPU oracle (same jitter pattern, same RNG seed) at the >= 99.9 %
within-1/255 bar, (2) tile sharding is invisible (2 shards assembled == 1 shard, bit for bit), (3) the ray counts
of the counting kernel satisfy the scene's invariants (one primary ray per sample; shadow rays <= lights x
soft-samples x shaded hits; reflection rays <= depth limit x primaries)."""
import numpy as np
import pytest

from functracer_b200 import abi, api, frontend, scenes
from oracle import ftb_oracle as orc

pytestmark = pytest.mark.gpu
SEED = 1234

CASES = ["cfg1-sample", "cfg2-hollow-sphere", "cfg3-house", "cfg3-night-house", "cfg4-bunny", "cfg4-bunny-d12", "cfg5-moon", "cfg5-repeat"]


@pytest.mark.parametrize("name", CASES)
def test_full_size_config(name):
    import torch
    cfg = scenes.CONFIGS[name]
    sc = frontend.ParsedScene(scenes.config_text(name), scenes.asset_dir())
    W, H, spp = sc.width, sc.height, sc.spp
    assert (W, H, spp) == (cfg["res"][0], cfg["res"][1], cfg["spp"])
    jit = frontend.jitter_pattern(cfg["seed"], spp)
    stream = torch.cuda.current_stream().cuda_stream
    with api.Scene(sc) as scene:
        p = api.make_params(W, H, spp, jit, seed=SEED, out_format=abi.OUT_RGB_F32)
        tiles = torch.empty(api.tile_buffer_bytes(p), dtype=torch.uint8, device="cuda")
        ps = api.make_params(W, H, spp, jit, seed=SEED, out_format=abi.OUT_RGB_F32, collect_stats=1)
        st = scene.render_tiles_device(ps, tiles.data_ptr(), stream=stream, stats=True)
        scene.render_tiles_device(p, tiles.data_ptr(), stream=stream)
        frame = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
        api.assemble_device(p, [tiles.data_ptr()], frame.data_ptr(), stream=stream)
        torch.cuda.synchronize()
        # (3) ray accounting
        n_lights = sc.desc.n_lights
        max_soft = max([sc.desc.lights[i].samples for i in range(n_lights) if sc.desc.lights[i].kind == abi.LIGHT_SOFT_DIRECTIONAL] + [1])
        assert st.primary_rays == W * H * spp
        assert st.shaded_hits <= st.primary_rays + st.reflection_rays
        assert st.shadow_rays <= st.shaded_hits * n_lights * max_soft
        assert st.reflection_rays <= 8 * st.primary_rays
        # (2) sharding invisibility on the smaller configs (two more full renders)
        if W * H * spp <= 40_000_000:
            bufs = []
            for k in range(2):
                pk = api.make_params(W, H, spp, jit, seed=SEED, out_format=abi.OUT_RGB_F32, shard_index=k, shard_count=2)
                b = torch.empty(api.tile_buffer_bytes(pk), dtype=torch.uint8, device="cuda")
                scene.render_tiles_device(pk, b.data_ptr(), stream=stream)
                bufs.append(b)
            frame2 = torch.empty_like(frame)
            api.assemble_device(pk, [b.data_ptr() for b in bufs], frame2.data_ptr(), stream=stream)
            torch.cuda.synchronize()
            assert bool((frame2 == frame).all())
        # (1) window parity around the brightest pixel
        idx = int(torch.argmax(torch.nan_to_num(frame.sum(dim=-1), nan=0.0)).item())
        ww, wh = (96, 64) if spp <= 16 else (48, 32)
        x0, y0 = min(max(0, idx % W - ww // 2), max(0, W - ww)), min(max(0, idx // W - wh // 2), max(0, H - wh))
        x1, y1 = min(W, x0 + ww), min(H, y0 + wh)
        ref = orc.render(sc, orc.make_params(W, H, spp, jit, seed=SEED), window=(x0, y0, x1, y1), debug=False)
        got = frame[y0:y1, x0:x1].cpu().numpy().astype(np.float64)
        d = np.abs(got - ref["rgb"][y0:y1, x0:x1]).max(axis=-1)
        frac = float((d <= 1.0 / 255.0).mean())
        print("%s: window %s within 1/255 on %.5f, max err %.3g, rays %d + %d + %d" % (name, (x0, y0, x1, y1), frac, np.nanmax(d), st.primary_rays, st.shadow_rays, st.reflection_rays))
        assert frac >= 0.999
