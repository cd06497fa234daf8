"""TEST INFRASTRUCTURE (not part of the product package): the tile sharding contract of the C ABI
(include/functracer_b200.h: FTB_TILE_W/H, shard_index/shard_count) restated on the host for the CPU tests of the
multi-process plumbing: which tiles a shard owns, how its tile-major buffer is laid out, and how N shard buffers
become the row-major frame (what ftb_assemble_device does on the GPU).

Pixels are independent in the reference (Shading.fs:141-147 shades 1000-ray chunks independently), so any
partition is legal; this one is 16x16 tiles dealt in groups of N consecutive tiles: group g gives one tile to every
shard (which one is rotated by a hash of g, device_scene.h tileOfLocal) and is every shard's local tile g.
"""
import numpy as np

from functracer_b200 import abi


def grid(width, height):
    tx = (width + abi.TILE_W - 1) // abi.TILE_W
    ty = (height + abi.TILE_H - 1) // abi.TILE_H
    return tx, ty


def shard_rot(group, n):
    """device_scene.h shardRot: which tile of group `group` shard 0 gets."""
    return ((((group * 2654435761) & 0xffffffff) >> 10) % n) if n > 1 else 0


def tile_of_local(ltile, shard, n):
    return ltile * n + (shard - shard_rot(ltile, n)) % n


def shard_of_tile(tile, n):
    g = tile // n
    return (tile - g * n + shard_rot(g, n)) % n


def local_tiles(width, height, shard_index, shard_count):
    """Global tile index of every local tile of a shard (None: the last group has no tile for this shard)."""
    tx, ty = grid(width, height)
    n_local = (tx * ty + shard_count - 1) // shard_count
    tiles = [tile_of_local(l, shard_index, shard_count) for l in range(n_local)]
    return [t if t < tx * ty else None for t in tiles]


def tile_buffer_elems(width, height, shard_index, shard_count):
    return len(local_tiles(width, height, shard_index, shard_count)) * abi.TILE_PIXELS * 3


def pack(frame, shard_index, shard_count):
    """Row-major frame [H, W, 3] -> the tile-major buffer a shard would have rendered (pixels of other
    shards are not touched; padding pixels of edge tiles are zero)."""
    H, W, _ = frame.shape
    tx, _ = grid(W, H)
    tiles = local_tiles(W, H, shard_index, shard_count)
    buf = np.zeros((len(tiles), abi.TILE_H, abi.TILE_W, 3), dtype=frame.dtype)
    for k, t in enumerate(tiles):
        if t is None:
            continue
        x0, y0 = (t % tx) * abi.TILE_W, (t // tx) * abi.TILE_H
        blk = frame[y0:y0 + abi.TILE_H, x0:x0 + abi.TILE_W]
        buf[k, :blk.shape[0], :blk.shape[1]] = blk
    return buf.reshape(-1)


def assemble(buffers, width, height):
    """N tile-major shard buffers -> row-major frame [H, W, 3] (host twin of ftb_assemble_device)."""
    n = len(buffers)
    tx, ty = grid(width, height)
    out = np.zeros((height, width, 3), dtype=np.asarray(buffers[0]).dtype)
    for t in range(tx * ty):
        shard, local = shard_of_tile(t, n), t // n
        blk = np.asarray(buffers[shard])[local * abi.TILE_PIXELS * 3:(local + 1) * abi.TILE_PIXELS * 3].reshape(abi.TILE_H, abi.TILE_W, 3)
        x0, y0 = (t % tx) * abi.TILE_W, (t // tx) * abi.TILE_H
        h, w = min(abi.TILE_H, height - y0), min(abi.TILE_W, width - x0)
        out[y0:y0 + h, x0:x0 + w] = blk[:h, :w]
    return out


def band_of_tile(t, width, height, band_count, rows_of_band):
    """The band (ftb_band_rows) a global tile belongs to: bands are whole tile rows."""
    tx, _ = grid(width, height)
    y0 = (t // tx) * abi.TILE_H
    for c in range(band_count):
        a, b = rows_of_band[c]
        if a <= y0 < b:
            return c
    raise ValueError("tile row outside every band")
