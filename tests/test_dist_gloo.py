"""The N>1 host logic on CPU: world_size-2 gloo runs of tidypopgen_b200.dist with a numpy
stand-in engine (the product engine needs a GPU).  Checks the loci partition + integer
all-reduce and the tile ownership partition against the oracle — and, on the way, the count
identities the CUDA finish kernels rely on (K = DD - NN + HN + HN', ibs = NN + HH + DD, ...)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from tidypopgen_b200 import _lib as L
from tidypopgen_b200 import dist as tdist
from tidypopgen_b200.fbm import FBMCode256


class NumpyEngine:
    """Test stand-in with the CudaEngine interface; same buffer layout (4 column-major
    int32 matrices, ld = n rounded up to 256, symmetric ones upper-only)."""

    device = 0

    def __init__(self, raw, nrow):
        self.raw = raw.reshape(-1, nrow)  # [locus, individual]

    def new_counts(self, n):
        return torch.zeros(int(L.lib().tpg_counts_elems(n)), dtype=torch.int32)

    @staticmethod
    def _idx(ptr, k, total):
        if not ptr:
            return np.arange(total)
        return np.ctypeslib.as_array((C.c_int32 * k).from_address(ptr)).astype(np.int64) - 1

    def accumulate(self, job, counts):
        rows = self._idx(job.ind_row, job.n, job.nrow_total)
        cols = self._idx(job.ind_col, job.m, job.ncol_total)
        n, m = len(rows), len(cols)
        rank, world = job.shard_rank, job.shard_world
        if job.shard_mode == L.SHARD_LOCI:
            lo, hi = tdist.loci_slice(m, rank, world)
            cols = cols[lo:hi]
        g = self.raw[np.ix_(cols, rows)].T.astype(np.int64)  # [n, loci]
        d = (g == 2).astype(np.int64) - (g == 0)
        nm = (g <= 2).astype(np.int64)
        h1 = (g == 1).astype(np.int64)
        ld = int(L.lib().tpg_counts_ld(n))
        buf = counts.numpy().reshape(4, ld, ld)  # [mat][j][i]
        mats = {L.MAT_DD: d @ d.T, L.MAT_NN: nm @ nm.T, L.MAT_HH: h1 @ h1.T, L.MAT_HN: h1 @ nm.T}
        own = tdist.owned_mask(n, rank, world) if job.shard_mode == L.SHARD_TILES else np.ones((n, n), bool)
        for k, v in mats.items():
            keep = own if k == L.MAT_HN else own & np.triu(np.ones((n, n), bool))
            buf[k, :n, :n] += np.where(keep, v, 0).T.astype(np.int32)

    def finish(self, job, counts, n, want_out1, on_device=False):
        ld = int(L.lib().tpg_counts_ld(n))
        buf = counts.numpy().reshape(4, ld, ld)[:, :n, :n].transpose(0, 2, 1).astype(np.float64)
        sym = lambda a: np.triu(a) + np.triu(a, 1).T  # noqa: E731
        DD, NN, HH, HN = sym(buf[0]), sym(buf[1]), sym(buf[2]), buf[3]
        m = job.m if job.ind_col else job.ncol_total
        with np.errstate(all="ignore"):
            if job.stat == L.STAT_KING:
                K, Ni, Nj = DD - NN + HN + HN.T, HN, HN.T
                pm = np.minimum(Ni, Nj)
                out = K / (2 * pm) + 0.5 - 0.25 * (Ni + Nj) / pm
            elif job.stat == L.STAT_IBS_RAW:
                return NN + HH + DD, 2 * NN, {}
            elif job.stat == L.STAT_AS:
                ns = L.lib().tpg_as_pad_count(m, job.block_size) if job.pad_quirk else 0
                out = 0.5 * (1 + (DD + ns) / NN)
            else:
                raise NotImplementedError
        if job.shard_mode == L.SHARD_TILES:
            out = np.where(tdist.owned_mask(n, job.shard_rank, job.shard_world), out, 0.0)
        return out, None, {}


def _worker(rank, world, port, raw, nrow, ncol, rows, cols, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X = FBMCode256.__new__(FBMCode256)
        X.backingfile, X.nrow, X.ncol, X.code256 = "<mem>", nrow, ncol, None
        eng = NumpyEngine(raw, nrow)
        res = {}
        for name, stat, bs in (("king", L.STAT_KING, 64), ("ibs", L.STAT_IBS_RAW, 64), ("as", L.STAT_AS, 50)):
            o0, o1 = tdist.pairwise_sharded(eng, X, rows, cols, stat, bs, mode="loci", rank=rank, world=world)
            res[name] = (o0, o1)
        t0, _ = tdist.pairwise_sharded(eng, X, rows, cols, L.STAT_KING, 64, mode="tiles", rank=rank, world=world)
        res["king_tiles"] = (t0, None)
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(180)
def test_world2_gloo_loci_and_tiles():
    rng = np.random.default_rng(4)
    nrow, ncol = 300, 401
    raw = rng.choice(np.array([0, 1, 2, 3, 5], np.uint8), size=ncol * nrow, p=[.4, .3, .2, .05, .05])
    rows = rng.permutation(nrow)[:270].astype(np.int32) + 1   # n = 270 -> 2 x 2 blocks of 256
    cols = rng.permutation(ncol)[:333].astype(np.int32) + 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, raw, nrow, ncol, rows, cols, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=150) for _ in procs)
    [p.join(30) for p in procs]
    ref_k = O.snp_king(raw, nrow, ncol, rows, cols, block_size_=64)["king"]
    ref_i = O.snp_ibs(raw, nrow, ncol, rows, cols, type="raw_counts", block_size_=64)
    ref_a = O.snp_allele_sharing(raw, nrow, ncol, rows, cols, block_size_=50)["as"]
    assert O.n_short(333, 50) > 0
    eq = lambda a, b: bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))  # noqa: E731
    for r in range(2):  # loci mode: complete and identical on every rank
        assert eq(got[r]["king"][0], ref_k)
        assert np.array_equal(got[r]["ibs"][0], ref_i["ibs"]) and np.array_equal(got[r]["ibs"][1], ref_i["valid_n"])
        assert eq(got[r]["as"][0], ref_a)
    # tiles mode: the ranks' blocks are disjoint and tile the matrix
    m0, m1 = tdist.owned_mask(270, 0, 2), tdist.owned_mask(270, 1, 2)
    assert not (m0 & m1).any() and (m0 | m1).all() and np.array_equal(m0, m0.T)
    merged = np.where(m0, got[0]["king_tiles"][0], got[1]["king_tiles"][0])
    assert eq(merged, ref_k)


def test_partitions_cover_exactly():
    for m, w in ((650000, 8), (7, 8), (1000001, 4), (5, 2)):
        sl = [tdist.loci_slice(m, r, w) for r in range(w)]
        assert sl[0][0] == 0 and sl[-1][1] == m and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
    for n, w in ((929, 2), (2504, 8), (5000, 3)):
        tot = sum(tdist.owned_mask(n, r, w).astype(int) for r in range(w))
        assert (tot == 1).all()
        nb = (n + 255) // 256
        r = tdist.shard_region_blocks(nb, w)          # ownership is dealt in r x r block regions
        nr = (nb + r - 1) // r
        cnt = np.bincount([tdist.block_owner(i * r, j * r, nb, w) for i in range(nr) for j in range(i, nr)], minlength=w)
        assert cnt.max() - cnt.min() <= 1  # round-robin balance over the regions
        assert nr * (nr + 1) // 2 >= 8 * w or r == 1
    # the host mirror and the library agree on every block's owner, round-robin and weighted regimes alike
    lib = L.lib()
    for n, w in ((929, 2), (5000, 3), (20000, 8), (50000, 8), (50000, 4), (50000, 2), (61000, 7)):
        nb = (n + 255) // 256
        blocks = [(i, j) for i in range(0, nb, max(1, nb // 23)) for j in range(i, nb, max(1, nb // 19))]
        for bi, bj in blocks:
            assert lib.tpg_tile_owner(n, bi, bj, w) == tdist.block_owner(bi, bj, nb, w), (n, w, bi, bj)
    # 3840 x 3840 regions are few and unequal: the weighted assignment keeps every rank within 4 % of the mean
    for w in (2, 4, 8):
        nb, r = 196, 15
        nr = (nb + r - 1) // r
        load = np.zeros(w)
        for ri in range(nr):
            for rj in range(ri, nr):
                h, wd = min(r, nb - ri * r), min(r, nb - rj * r)
                load[tdist.block_owner(ri * r, rj * r, nb, w)] += h * wd if ri == rj else 2 * h * wd
        assert load.max() / load.mean() < 1.04
    # 50,000 individuals on 8 ranks: 3840 x 3840 regions (whole 256-row and 240- / 480-column tiles);
    # 20,000 on 8: too few of those, 1024 x 1024; small matrices 512, 256
    assert tdist.shard_region_blocks(196, 8) == 15 and tdist.shard_region_blocks(79, 8) == 4
    assert tdist.shard_region_blocks(24, 8) == 2 and tdist.shard_region_blocks(4, 2) == 1
    assert tdist.choose_mode(929, 8) == "loci" and tdist.choose_mode(50000, 8) == "tiles"


def test_result_shell_geometry():
    """tpg_result_shell_bounds: nested shells on 3840-wide region boundaries — a first square of ~45 % of the area,
    rims of equal area, as many shells (6, 4, 3) as the regions allow; small matrices come home in one piece."""
    import ctypes as C
    from tidypopgen_b200 import _lib as L
    lib = L.lib()
    b = (C.c_int64 * 9)()
    for n in (100, 5000, 20000, 23000):
        assert lib.tpg_result_shell_bounds(n, 0, b) == 0, n
    assert lib.tpg_result_shell_bounds(27136 + 77, 0, b) == 4
    assert lib.tpg_result_shell_bounds(27136 + 77, 6, b) == 0 and lib.tpg_result_shell_bounds(27136 + 77, 1, b) == 0
    for n, asked, want in ((50000, 0, 6), (50000, 4, 4), (50000, 3, 3), (35000, 0, 6), (30000, 0, 4), (400000, 8, 8)):
        K = lib.tpg_result_shell_bounds(n, asked, b)
        assert K == want, (n, asked, K)
        bounds = list(b)[: K + 1]
        assert bounds[0] == 0 and bounds[-1] == n and all(x < y for x, y in zip(bounds, bounds[1:]))
        assert all(x % 3840 == 0 for x in bounds[:-1])
        area = [(y * y - x * x) / (n * n) for x, y in zip(bounds, bounds[1:])]
        assert 0.3 < area[0] < 0.6, area
        assert abs(sum(area) - 1) < 1e-12 and max(area[1:]) <= 0.4 and (K < 6 or max(area[1:]) < 0.25), area
