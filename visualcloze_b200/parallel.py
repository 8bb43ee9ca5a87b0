"""Multi-GPU plumbing: one process per GPU.

(1) Throughput: batched grids sharded by sample, one all-gather of decoded tiles (below).
(2) Latency: ONE sample sharded by token rows over the ranks -- ``SequenceParallel`` -- with the attention
    all-to-alls fused into the QKV-GEMM / attention epilogues as NVLink peer stores (include/vcb200.h, "sequence
    parallelism"); torch.distributed only carries the IPC handles at set-up and the final gather of the trajectory.

The reference's inference is single-GPU (sample.py:258); independent grid samples have no cross-sample term, so the
path shards by sample with NO per-step collective.  The only exchange is the all-gather of the
decoded query-row tiles at the end of a batch, over NCCL (NVLink 5 / NVSwitch) on GPUs, gloo in the CPU tests.
A sample's result must not depend on the world size or the rank that computed it.
"""
from __future__ import annotations

import ctypes as C

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


# ------------------------------------------------------------------------------------------------------
# sequence parallelism of one sample (token rows over ranks)
# ------------------------------------------------------------------------------------------------------
def sp_row_slice(n_rows: int, rank: int, world: int) -> slice:
    """rows [rank * n/W, (rank + 1) * n/W) of a stream of ``n_rows`` tokens; every rank must get the same count"""
    if n_rows % world:
        raise ValueError(f"sequence parallelism needs token counts divisible by the world size ({n_rows} % {world} != 0)")
    per = n_rows // world
    return slice(rank * per, (rank + 1) * per)


def sp_shard_rows(t: torch.Tensor, rank: int, world: int, dim: int = 1) -> torch.Tensor:
    sl = [slice(None)] * t.ndim
    sl[dim] = sp_row_slice(t.shape[dim], rank, world)
    return t[tuple(sl)]


def sp_gather_rows(local: torch.Tensor, dim: int = 1, group=None) -> torch.Tensor:
    """inverse of sp_shard_rows: all-gather the ranks' row blocks (equal sizes) along ``dim``; every rank gets the whole"""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    loc = local.contiguous()
    if loc.ndim == 0:
        raise ValueError("sp_gather_rows needs at least one dimension")
    buf = torch.empty((world * loc.shape[0],) + tuple(loc.shape[1:]), dtype=loc.dtype, device=loc.device)
    dist.all_gather_into_tensor(buf, loc, group=group)        # concatenated along dim 0 (the form gloo and NCCL both take)
    return torch.cat(list(buf.reshape((world,) + tuple(loc.shape)).unbind(0)), dim=dim)


class PeerBuffer:
    """One zero-filled device allocation per rank, mapped into every rank of the group through CUDA IPC
    (vcb_peer_alloc / vcb_peer_open).  ``ptrs[r]`` is rank r's allocation as addressable from THIS process."""

    def __init__(self, nbytes: int, group=None):
        from . import _lib
        self.lib = _lib.lib()
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.nbytes = int(nbytes)
        local = C.c_void_p()
        handle = C.create_string_buffer(64)
        _lib.check(self.lib.vcb_peer_alloc(self.nbytes, C.byref(local), handle), "vcb_peer_alloc")
        self.local = local.value
        handles: list = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)      # also: every rank has allocated + zeroed
        self.ptrs: list[int] = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.local)
                continue
            p = C.c_void_p()
            _lib.check(self.lib.vcb_peer_open(C.create_string_buffer(h, 64), C.byref(p)), f"vcb_peer_open(rank {r})")
            self.ptrs.append(p.value)

    def array(self):
        arr = (C.c_void_p * self.world)(*self.ptrs)
        return arr

    def close(self):
        """collective: unmap the peers' allocations, then free ours (all ranks must have finished using it)"""
        if self.local is None:
            return
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                self.lib.vcb_peer_close(p)
        dist.barrier(group=self.group)
        self.lib.vcb_peer_free(self.local)
        self.local = None
        self.ptrs = []


class SequenceParallel:
    """Shares ONE sample across the ranks of ``group``: rank r holds txt rows and img rows ``sp_row_slice`` and attends the
    heads [r * heads / W, (r + 1) * heads / W) of all rows.  ``attach`` (collective) gives an engine the peer buffers for
    a local shape; the per-step path then has no host-side communication at all."""

    def __init__(self, group=None, timeout_ms: int = 10000):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("SequenceParallel needs an initialised torch.distributed process group (one process per GPU)")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        from . import _lib
        if self.world > _lib.SP_MAX:
            raise ValueError(f"at most {_lib.SP_MAX} ranks per sequence-parallel group")
        self.timeout_ms = int(timeout_ms)
        self._bufs: tuple | None = None
        self._shape = None
        self.err = torch.zeros(1, dtype=torch.int32, device="cuda")

    def row_slice(self, n_rows: int) -> slice:
        return sp_row_slice(n_rows, self.rank, self.world)

    def shard(self, t: torch.Tensor, dim: int = 1) -> torch.Tensor:
        return sp_shard_rows(t, self.rank, self.world, dim)

    def gather(self, local: torch.Tensor, dim: int = 1) -> torch.Tensor:
        return sp_gather_rows(local, dim, self.group)

    def attach(self, engine, li_local: int, lt_local: int) -> None:
        """collective; (re)allocates the shared qkv / cat / flag buffers when the local shape changes"""
        from . import _lib
        key = (engine, li_local, lt_local)            # the engine object itself (not id()): a recycled id must not alias
        if self._shape is not None and self._shape[0] is engine and self._shape[1:] == key[1:]:
            return
        self.release()
        qb, cb = C.c_int64(), C.c_int64()
        _lib.check(engine.lib.vcb_flux_sp_shared_bytes(engine._h, li_local, lt_local, C.byref(qb), C.byref(cb)), "vcb_flux_sp_shared_bytes")
        qkv, cat, flags = PeerBuffer(qb.value, self.group), PeerBuffer(cb.value, self.group), PeerBuffer(256, self.group)
        self._bufs = (qkv, cat, flags)
        self.err.zero_()
        _lib.check(engine.lib.vcb_flux_sp_attach(engine._h, self.world, self.rank, qkv.array(), cat.array(), flags.array(),
                                                 self.err.data_ptr(), self.timeout_ms), "vcb_flux_sp_attach")
        self._shape = key

    def check(self) -> None:
        """after a synchronisation point: raise if a phase barrier timed out (a peer died or fell behind)"""
        e = int(self.err.item())
        if e:
            raise RuntimeError(f"sequence-parallel barrier {e} timed out on rank {self.rank}: a peer did not arrive within {self.timeout_ms} ms")

    def release(self) -> None:
        if self._bufs:
            for b in self._bufs:
                b.close()
        self._bufs, self._shape = None, None
