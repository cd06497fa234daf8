"""The N > 1 plumbing on CPU: two gloo ranks each own the tiles `t % 2 == rank` of a frame, "render" them band by band
(the band geometry comes from the library's own ftb_band_rows / ftb_tile_buffer_bytes, which need no device), gather their
tile-major buffers to rank 0 (functracer_b200.dist.gather_tiles) and rank 0 assembles the frame with the host twin of
ftb_assemble_device (tests/tile_twin.py).  The per-shard pixels come from the CPU oracle (there is no GPU here);
what is under test is the partition / layout / gather contract of the C ABI."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import tile_twin as tiles
from functracer_b200 import abi, api, frontend, scenes
from oracle import ftb_oracle as orc

W, H, SPP = 70, 37, 2  # deliberately not multiples of the 16 x 16 tile


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _frame():
    sc = frontend.ParsedScene(scenes.hollow_sphere(res=(W, H), spp=SPP), scenes.asset_dir())
    jit = frontend.jitter_pattern(2, SPP)
    return orc.render(sc, orc.make_params(W, H, SPP, jit), threads=2, debug=False)["rgb"]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from functracer_b200 import dist as fdist
        frame = _frame()  # every rank "renders" with the oracle and keeps only its own tiles
        mine = tiles.pack(frame, rank, world)
        # ... band by band, as bench.py's multi-process e2e path does: the bands of the library partition this shard's tiles
        bands = 3
        pb = api.make_params(W, H, SPP, [0.0] * (2 * SPP), shard_index=rank, shard_count=world)
        rows = [api.band_rows(pb, c, bands) for c in range(bands)]
        assert rows[0][0] == 0 and rows[-1][1] == H and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        banded = np.zeros_like(mine)
        mine_tiles = tiles.local_tiles(W, H, rank, world)
        n = abi.TILE_PIXELS * 3
        for c in range(bands):
            for k, t in enumerate(mine_tiles):
                if t is not None and tiles.band_of_tile(t, W, H, bands, rows) == c:
                    banded[k * n:(k + 1) * n] = mine[k * n:(k + 1) * n]
        assert (banded == mine).all()
        # the C ABI agrees on the buffer size of this shard (argument-only call: no device needed)
        p = api.make_params(W, H, SPP, [0.0] * (2 * SPP), shard_index=rank, shard_count=world, precision=abi.PRECISION_FP64_VERIFY)
        assert api.tile_buffer_bytes(p) == mine.size * 8
        max_elems = max(tiles.tile_buffer_elems(W, H, k, world) for k in range(world))
        buf = torch.zeros(max_elems, dtype=torch.float64)
        buf[:mine.size] = torch.from_numpy(mine)
        got = fdist.gather_tiles(buf, max_elems)
        if rank == 0:
            out = tiles.assemble([g.numpy() for g in got], W, H)
            q.put(bool((out == frame).all()))
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


def test_two_rank_tile_gather_reassembles_the_frame():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_partition_is_exact():
    for n in (1, 2, 3, 4, 8):
        seen = sorted(t for k in range(n) for t in tiles.local_tiles(W, H, k, n) if t is not None)
        tx, ty = tiles.grid(W, H)
        assert seen == list(range(tx * ty))
    rng = np.random.default_rng(0)
    frame = rng.random((H, W, 3))
    for n in (1, 2, 5):
        assert (tiles.assemble([tiles.pack(frame, k, n) for k in range(n)], W, H) == frame).all()
