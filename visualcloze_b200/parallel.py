"""Multi-GPU plumbing: one process per GPU, batched grids sharded by sample, one all-gather of decoded tiles.

The reference's inference is single-GPU (sample.py:258); independent grid samples have no cross-sample term, so the
path shards by sample with NO per-step collective.  The only exchange is the all-gather of the
decoded query-row tiles at the end of a batch, over NCCL (NVLink 5 / NVSwitch) on GPUs, gloo in the CPU tests.
A sample's result must not depend on the world size or the rank that computed it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_samples(n_samples: int, rank: int, world: int) -> list[int]:
    """round-robin: rank r takes samples r, r + W, r + 2W, ..."""
    return list(range(rank, n_samples, world))


def gather_tiles(tiles: list[torch.Tensor], n_samples: int, group=None) -> list[torch.Tensor]:
    """All-gather per-sample image tiles ([C, h, w], any integer/float dtype, possibly different sizes) so that every rank
    ends with the list of all ``n_samples`` tiles in sample order.  Shapes are exchanged first; tiles are padded to the
    batch maximum for one fixed-size all_gather_into_tensor, then cropped."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        assert len(tiles) == n_samples
        return tiles
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = shard_samples(n_samples, rank, world)
    assert len(tiles) == len(mine), f"rank {rank} computed {len(tiles)} tiles, expected {len(mine)}"
    per_rank = (n_samples + world - 1) // world
    dev = tiles[0].device if tiles else torch.device("cpu")
    dtype = tiles[0].dtype if tiles else torch.uint8
    shp = torch.zeros(per_rank, 3, dtype=torch.int64, device=dev)
    for i, t in enumerate(tiles):
        shp[i] = torch.tensor(t.shape, dtype=torch.int64)
    all_shp = torch.empty(world * per_rank, 3, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_shp, shp, group=group)
    cmax, hmax, wmax = (int(v) for v in all_shp.max(dim=0).values)
    buf = torch.zeros(per_rank, cmax, hmax, wmax, dtype=dtype, device=dev)
    for i, t in enumerate(tiles):
        buf[i, : t.shape[0], : t.shape[1], : t.shape[2]] = t
    allbuf = torch.empty(world * per_rank, cmax, hmax, wmax, dtype=dtype, device=dev)
    dist.all_gather_into_tensor(allbuf, buf, group=group)
    out: list[torch.Tensor] = [None] * n_samples  # type: ignore[list-item]
    for r in range(world):
        for i, s in enumerate(shard_samples(n_samples, r, world)):
            c, h, w = (int(v) for v in all_shp[r * per_rank + i])
            out[s] = allbuf[r * per_rank + i, :c, :h, :w].clone()
    return out
