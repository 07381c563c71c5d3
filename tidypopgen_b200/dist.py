"""Sharding layer: one process per GPU, `torch.distributed` for the plumbing.

The reference has no multi-process path at all (single R process; the only decomposition is
its serial loop over loci blocks, R/snp_king.R:63-77).  Two exact partitions exist because
every accumulated quantity is an integer sum ("Multi-GPU"):

* ``loci``  — rank r takes the r-th contiguous slice of ``ind_col`` (the reduction dimension),
  accumulates full n x n int32 count matrices with the staged C-ABI, then ONE all-reduce
  (NCCL over NVLink on GPUs) sums them; any rank can finish.  For moderate n.
* ``tiles`` — rank r owns every W-th 256 x 256 block of the upper triangle, streams ALL loci,
  and finishes only its own blocks: no data-path collective.  For large n, where the count
  matrices (12 n^2 bytes) are the big object.

The engine (what accumulates / finishes) is injectable so the partition + combine logic is
testable on CPU with gloo; the product engine is `CudaEngine` (libtpg_b200.so), and nothing
else ships.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from . import snp


def init_process_group(backend: str | None = None):
    """Rendezvous from the torchrun environment (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).
    Returns (rank, world, local_rank).  A single process needs no group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def loci_slice(m: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) positions of ind_col that rank owns — identical to the C side
    (make_plan in csrc/api.cu): m*rank/world .. m*(rank+1)/world."""
    return m * rank // world, m * (rank + 1) // world


def shard_region_blocks(nblocks: int, world: int) -> int:
    """Side (in 256-blocks) of the square regions tile ownership is dealt in — identical to
    shard_region_blocks() in csrc/internal.h: 15 (3840 x 3840: whole 256-row and 240- / 480-column tiles of
    the 4-bit Gram kernels) while a rank gets at least 8 regions, else 4 (1024 x 1024), 2, 1."""
    nr = (nblocks + 14) // 15
    if nr * (nr + 1) // 2 >= 8 * world:
        return 15
    r = 4
    while r > 1:
        nr = (nblocks + r - 1) // r
        if nr * (nr + 1) // 2 >= 8 * world:
            break
        r >>= 1
    return r


_owner_cache: dict = {}


def owner_table(nblocks: int, world: int):
    """Region -> rank for the large (15 x 15 block) regions — identical to build_owner_table() in csrc/api.cu:
    regions sorted by weight (2 h w blocks off the diagonal, h w on it), largest first, each given to the
    least loaded rank (ties: lower region index, lower rank).  None where round-robin applies (smaller regions)."""
    key = (nblocks, world)
    if key in _owner_cache:
        return _owner_cache[key]
    r = shard_region_blocks(nblocks, world)
    table = None
    if world > 1 and r == 15:
        nr = (nblocks + r - 1) // r
        regs = []
        for ri in range(nr):
            for rj in range(ri, nr):
                h, w = min(r, nblocks - ri * r), min(r, nblocks - rj * r)
                regs.append((-(h * w if ri == rj else 2 * h * w), ri * nr - ri * (ri - 1) // 2 + (rj - ri)))
        regs.sort()
        load = [0] * world
        table = np.zeros(nr * (nr + 1) // 2, np.int64)
        for negw, idx in regs:
            best = min(range(world), key=lambda k: (load[k], k))
            load[best] -= negw
            table[idx] = best
    _owner_cache[key] = table
    return table


def block_owner(bi: int, bj: int, nblocks: int, world: int) -> int:
    """Owner rank of output block (bi <= bj) of the 256 x 256 block grid — identical to
    block_owned() in csrc/internal.h."""
    r = shard_region_blocks(nblocks, world)
    ri, rj, nr = bi // r, bj // r, (nblocks + r - 1) // r
    idx = ri * nr - ri * (ri - 1) // 2 + (rj - ri)
    table = owner_table(nblocks, world)
    return int(table[idx]) if table is not None else idx % world


def owned_mask(n: int, rank: int, world: int) -> np.ndarray:
    """Boolean n x n mask of the entries rank finishes in tiles mode (block + mirror)."""
    nb = (n + 255) // 256
    r = shard_region_blocks(nb, world)
    nr = (nb + r - 1) // r
    b = np.arange(n) // 256 // r
    bi, bj = np.minimum.outer(b, b), np.maximum.outer(b, b)
    idx = bi * nr - bi * (bi - 1) // 2 + (bj - bi)
    table = owner_table(nb, world)
    return (table[idx] if table is not None else idx % world) == rank


def choose_mode(n: int, world: int) -> str:
    """`auto`: loci sharding while three int32 count matrices stay small next to HBM."""
    return "loci" if world == 1 or n <= 30000 else "tiles"


def needed_matrices(stat: int) -> list[int]:
    """Count matrices a statistic reads (prod_mask_for in csrc/api.cu): KING DD, NN, HN; IBS DD, NN, HH;
    allele sharing / GRM DD, NN."""
    if stat == L.STAT_KING:
        return [L.MAT_DD, L.MAT_NN, L.MAT_HN]
    if stat in (L.STAT_IBS_RAW, L.STAT_IBS_PROPORTION, L.STAT_IBS_ADJUSTED):
        return [L.MAT_DD, L.MAT_NN, L.MAT_HH]
    return [L.MAT_DD, L.MAT_NN]


def all_reduce_counts(counts: torch.Tensor, stat: int, n: int, group=None):
    """The one exchange step of a loci-sharded job: integer SUM of the count matrices the statistic
    needs — contiguous runs of them, so KING (DD, NN | HN) costs two collectives over 3/4 of the
    buffer and allele sharing one over half of it, instead of one over all four matrices."""
    ld = int(L.lib().tpg_counts_ld(n))
    ms = ld * ld
    mats = needed_matrices(stat)
    k = 0
    while k < len(mats):
        j = k
        while j + 1 < len(mats) and mats[j + 1] == mats[j] + 1:
            j += 1
        dist.all_reduce(counts[mats[k] * ms:(mats[j] + 1) * ms], op=dist.ReduceOp.SUM, group=group)
        k = j + 1


class CudaEngine:
    """The product engine: staged C-ABI calls on this rank's GPU."""

    def __init__(self, device: int):
        self.device = device
        self.ctx = L.Context(device)
        self.stream = torch.cuda.ExternalStream(L.lib().tpg_ctx_stream(self.ctx.handle), device=device)
        self.last_stats = None

    def new_counts(self, n: int) -> torch.Tensor:
        return torch.zeros(int(L.lib().tpg_counts_elems(n)), dtype=torch.int32, device=f"cuda:{self.device}")

    def _order_after_torch(self):
        # buffers handed to the library are filled (zeroed) by torch on ITS current stream; the
        # library's kernels run on the context's stream, so order that stream after torch's
        self.stream.wait_stream(torch.cuda.current_stream(self.device))

    def accumulate(self, job, counts: torch.Tensor):
        self._order_after_torch()
        st = L.Stats()
        eb = L.errbuf()
        rc = L.lib().tpg_accumulate(self.ctx.handle, C.byref(job), counts.data_ptr(), C.byref(st), eb, len(eb))
        L.check(rc, eb)
        self.last_stats = st.as_dict()
        return self.last_stats

    def finish(self, job, counts: torch.Tensor, n: int, want_out1: bool, on_device: bool = False):
        res = L.Result()
        if on_device:
            out0 = torch.zeros((n, n), dtype=torch.float64, device=counts.device)
            out1 = torch.zeros((n, n), dtype=torch.float64, device=counts.device) if want_out1 else None
            res.out0 = out0.data_ptr()
            res.out1 = out1.data_ptr() if want_out1 else None
        else:
            out0 = np.zeros((n, n), np.float64, order="F")
            out1 = np.zeros((n, n), np.float64, order="F") if want_out1 else None
            res.out0 = out0.ctypes.data
            res.out1 = out1.ctypes.data if want_out1 else None
        res.out_on_device = 1 if on_device else 0
        if on_device:
            self._order_after_torch()
        eb = L.errbuf()
        rc = L.lib().tpg_finish(self.ctx.handle, C.byref(job), counts.data_ptr(), C.byref(res), eb, len(eb))
        L.check(rc, eb)
        if on_device:  # column-major storage: expose as [i, j]
            out0 = out0.t()
            out1 = out1.t() if want_out1 else None
        return out0, out1, res.stats.as_dict()


def pairwise_sharded(engine, X, ind_row=None, ind_col=None, stat: int = L.STAT_KING, block_size: int | None = None,
                     mode: str = "auto", rank: int = 0, world: int = 1, group=None, src=None,
                     pad_quirk: bool = True, gpu_chunk_loci: int = 0, flags: int = 0, on_device: bool = False):
    """Runs one job across `world` ranks.  Every rank calls this with the same arguments.

    loci mode: returns (out0, out1) complete on every rank.
    tiles mode: returns this rank's blocks (zeros elsewhere); `owned_mask` says which.
    """
    n = X.nrow if ind_row is None else len(np.asarray(ind_row).reshape(-1))
    if block_size is None:
        block_size = snp.block_size(X.nrow) * (4 if stat == L.STAT_KING else 1)
    if mode == "auto":
        mode = choose_mode(n, world)
    if world == 1:
        mode = "none"
    shard = {"none": L.SHARD_NONE, "loci": L.SHARD_LOCI, "tiles": L.SHARD_TILES}[mode]
    if stat == L.STAT_GRM and mode == "tiles":
        raise ValueError("GRM needs the whole allele-sharing matrix: use mode='loci', or compute AS "
                         "tile-sharded, gather it, and call grm_from_allele_sharing")
    job, keep = snp.make_job(X, ind_row, ind_col, stat, block_size, device=getattr(engine, "device", 0),
                             pad_quirk=pad_quirk, src=src, gpu_chunk_loci=gpu_chunk_loci, flags=flags,
                             shard_mode=shard, shard_rank=rank, shard_world=world)
    counts = engine.new_counts(n)
    stream = getattr(engine, "stream", None)
    if mode == "loci" and stream is not None:
        job.flags |= L.FLAG_ASYNC  # accumulate -> all-reduce -> finish queue up on one stream, one wait at the end
    engine.accumulate(job, counts)
    if mode == "loci":
        # the one exchange step of the path: integer sum of the count matrices
        if stream is not None:
            with torch.cuda.stream(stream):
                all_reduce_counts(counts, stat, n, group)
        else:
            all_reduce_counts(counts, stat, n, group)
        job.shard_mode = L.SHARD_NONE  # finish sees the whole job (m = all loci, pad constant once)
        job.shard_world = 1
        job.shard_rank = 0
    out0, out1, _ = engine.finish(job, counts, n, stat == L.STAT_IBS_RAW, on_device=on_device)
    del keep
    return out0, out1
