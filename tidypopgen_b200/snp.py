"""bigSNP-level drivers: snp_king, snp_ibs, snp_allele_sharing.

Host-side mirror of the reference's L4 drivers (R/snp_king.R:32-103, R/snp_ibs.R:42-104,
R/snp_allele_sharing.R:33-82): same argument names (`ind.row` -> `ind_row`, `block.size` ->
`block_size`), same defaults, same return types, same error text.  Where the reference
loops over loci blocks and `.Call`s a native kernel per block, these make ONE call into
libtpg_b200.so per job; there is no CPU implementation behind them.

Indices follow R: `ind_row` / `ind_col` are 1-based, any order, any subset, no negatives
(R/snp_king.R:12-15).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L
from .fbm import FBMCode256, FBMDouble

_STAT_BY_TYPE = {
    "proportion": L.STAT_IBS_PROPORTION,
    "adjusted_counts": L.STAT_IBS_ADJUSTED,
    "raw_counts": L.STAT_IBS_RAW,
}

last_stats: dict | None = None  # per-stage timings of the most recent job (tpg_stats)


def block_size(n: int) -> int:
    """bigstatsr::block_size(n) with the default 1 GB option."""
    return int(L.lib().tpg_block_size(int(n)))


def cut_by_size(m: int, block_size_: int) -> np.ndarray:
    """CutBySize(m, block.size) (R/local_reimplementations.R:13-15): rows (lower, upper, size)."""
    lib = L.lib()
    nb = int(lib.tpg_split_len(int(m), int(block_size_), None, None, None, 0))
    if nb < 0:
        raise ValueError("m and block.size must be positive")
    lo, up, sz = (np.zeros(nb, np.int64) for _ in range(3))
    lib.tpg_split_len(int(m), int(block_size_), lo.ctypes.data, up.ctypes.data, sz.ctypes.data, nb)
    return np.stack([lo, up, sz], axis=1)


def _index(ind, total: int, what: str):
    if ind is None:
        return None, total
    a = np.asarray(ind)
    if a.dtype == bool:
        raise TypeError(f"{what}: logical indices are not supported, pass 1-based integers")
    if a.size and not np.all(np.equal(np.mod(a, 1), 0)):
        raise ValueError(f"{what} must be whole numbers")
    a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1)
    return a, int(a.size)


def make_job(X: FBMCode256, ind_row, ind_col, stat: int, block_size_: int, *, device: int = 0,
             pad_quirk: bool = True, src=None, gpu_chunk_loci: int = 0, flags: int = 0,
             shard_mode: int = L.SHARD_NONE, shard_rank: int = 0, shard_world: int = 1):
    """Builds a tpg_job.  Returns (job, keepalive) — keepalive holds the index arrays."""
    rows, n = _index(ind_row, X.nrow, "ind.row")
    cols, m = _index(ind_col, X.ncol, "ind.col")
    job = L.Job()
    job.abi_version = L.ABI_VERSION
    job.stat = stat
    job.device = device
    keep = [rows, cols]
    if src is None:
        job.src_kind = L.SRC_FILE
        path = X.backingfile.encode()
        keep.append(path)
        job.bk_path = path
    else:
        kind, ptr, holder = src
        job.src_kind = kind
        job.bk_data = ptr
        keep.append(holder)
    job.nrow_total = X.nrow
    job.ncol_total = X.ncol
    job.ind_row = rows.ctypes.data if rows is not None else None
    job.n = n
    job.ind_col = cols.ctypes.data if cols is not None else None
    job.m = m
    job.block_size = int(block_size_)
    job.pad_quirk = 1 if pad_quirk else 0
    job.shard_mode = shard_mode
    job.shard_rank = shard_rank
    job.shard_world = shard_world
    job.gpu_chunk_loci = int(gpu_chunk_loci)
    job.flags = flags
    return job, keep


def _colmajor(n: int) -> np.ndarray:
    return np.zeros((n, n), dtype=np.float64, order="F")


def run_job(X: FBMCode256, ind_row, ind_col, stat: int, block_size_: int, *, want_aux: bool = False,
            out0=None, out1=None, out0_path=None, out1_path=None, devices=None, shard: str = "auto", **kw):
    """One tpg_pairwise call (or tpg_pairwise_multi when `devices` lists several GPUs).
    Returns (out0, out1, aux0, aux1) as [i, j]-indexed arrays.  out0_path / out1_path: the library maps
    these files itself and writes the n x n doubles into them (the .bk of a bigstatsr FBM(n, n))."""
    global last_stats
    if devices is not None and len(devices) >= 1:
        kw.setdefault("device", int(devices[0]))
    job, keep = make_job(X, ind_row, ind_col, stat, block_size_, **kw)
    n = job.n if job.ind_row else X.nrow
    if n <= 0 or (job.m if job.ind_col else X.ncol) <= 0:
        raise ValueError("ind.row and ind.col must not be empty")
    o0 = (_colmajor(n) if out0_path is None else None) if out0 is None else out0
    o1 = out1
    if stat == L.STAT_IBS_RAW and o1 is None and out1_path is None:
        o1 = _colmajor(n)
    a0 = _colmajor(n) if want_aux and stat != L.STAT_IBS_RAW else None
    a1 = _colmajor(n) if want_aux and stat not in (L.STAT_IBS_RAW, L.STAT_GRM) else None
    res = L.Result()
    res.out0 = o0.ctypes.data if o0 is not None else None
    res.out1 = o1.ctypes.data if o1 is not None else None
    p0 = os.fsencode(out0_path) if out0_path is not None else None
    p1 = os.fsencode(out1_path) if out1_path is not None else None
    res.out0_path, res.out1_path = p0, p1
    res.aux0 = a0.ctypes.data if a0 is not None else None
    res.aux1 = a1.ctypes.data if a1 is not None else None
    res.out_on_device = 0
    eb = L.errbuf()
    if devices is not None and len(devices) > 1:
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        mode = {"auto": L.SHARD_NONE, "loci": L.SHARD_LOCI, "tiles": L.SHARD_TILES}[shard]
        rc = L.lib().tpg_pairwise_multi(C.byref(job), C.byref(res), devs.ctypes.data, len(devs), mode, eb, len(eb))
    else:
        rc = L.lib().tpg_pairwise(C.byref(job), C.byref(res), eb, len(eb))
    del keep
    L.check(rc, eb)
    last_stats = res.stats.as_dict()
    return o0, o1, a0, a1


def _check_fbm(X):
    # R/snp_king.R:39-41
    if not (hasattr(X, "inherits") and X.inherits("FBM.code256")):
        raise TypeError("X should be a FBM.code256 matrix")


def snp_king(X, ind_row=None, ind_col=None, block_size=None, **kw) -> np.ndarray:
    """KING-robust kinship matrix (Manichaikul et al. 2010) — R/snp_king.R:32-103.

    Returns a symmetric n x n float64 matrix; pairs with no heterozygous site in either
    individual give NaN / Inf exactly as the reference's IEEE arithmetic does."""
    _check_fbm(X)
    if block_size is None:
        block_size = globals()["block_size"](X.nrow) * 4  # R/snp_king.R:36
    king, _, _, _ = run_job(X, ind_row, ind_col, L.STAT_KING, block_size, **kw)
    return king


def snp_king_counts(X, ind_row=None, ind_col=None, block_size=None, **kw):
    """snp_king plus its integer accumulators: (king, K numerator, N_Aa_i) — the FBMs
    `k_numerator` and `N_Aa_i` of R/snp_king.R:45-48 as in-memory matrices."""
    _check_fbm(X)
    if block_size is None:
        block_size = globals()["block_size"](X.nrow) * 4
    king, _, K, N = run_job(X, ind_row, ind_col, L.STAT_KING, block_size, want_aux=True, **kw)
    return king, K, N


def snp_ibs(X, ind_row=None, ind_col=None, type="proportion", block_size=None, **kw):
    """Identity-by-state — R/snp_ibs.R:42-104.

    type = "proportion" (PLINK "ibs"), "adjusted_counts" (PLINK "distance") return an n x n
    matrix; "raw_counts" returns {"ibs": FBM, "valid_n": FBM}, two file-backed double
    matrices, as the reference does (R/snp_ibs.R:84-85)."""
    if type not in _STAT_BY_TYPE:  # match.arg
        raise ValueError("'arg' should be one of 'proportion', 'adjusted_counts', 'raw_counts'")
    _check_fbm(X)
    if block_size is None:
        block_size = globals()["block_size"](X.nrow)  # R/snp_ibs.R:47
    if type == "raw_counts":
        n = X.nrow if ind_row is None else len(np.asarray(ind_row).reshape(-1))
        ibs, valid = FBMDouble(n, n, init=None), FBMDouble(n, n, init=None)
        # the library maps the two FBMs' backing files itself (tpg_result.out0_path / out1_path) and
        # writes the counts there — what the R glue does with ibs$backingfile / valid_n$backingfile
        run_job(X, ind_row, ind_col, L.STAT_IBS_RAW, block_size, out0_path=ibs.backingfile,
                out1_path=valid.backingfile, **kw)
        return {"ibs": ibs, "valid_n": valid}
    out, _, _, _ = run_job(X, ind_row, ind_col, _STAT_BY_TYPE[type], block_size, **kw)
    return out


def snp_allele_sharing(X, ind_row=None, ind_col=None, block_size=None, pad_quirk=True, **kw) -> np.ndarray:
    """Allele sharing (hierfstat::matching) — R/snp_allele_sharing.R:33-82.

    `pad_quirk=True` (default) reproduces the reference bit for bit, including the +1 its
    native kernel adds to every numerator for each loci block shorter than the longest one
    (src/snp_as.cpp:56-64), which makes the result depend on `block_size`;
    `pad_quirk=False` gives the block-size independent value."""
    _check_fbm(X)
    if block_size is None:
        block_size = globals()["block_size"](X.nrow)  # R/snp_allele_sharing.R:37
    out, _, _, _ = run_job(X, ind_row, ind_col, L.STAT_AS, block_size, pad_quirk=pad_quirk, **kw)
    return out


def snp_allele_sharing_counts(X, ind_row=None, ind_col=None, block_size=None, pad_quirk=True, **kw):
    """(allele sharing, dos_tcross, na_tcross) — the accumulators of R/snp_allele_sharing.R:43-46."""
    _check_fbm(X)
    if block_size is None:
        block_size = globals()["block_size"](X.nrow)
    out, _, num, den = run_job(X, ind_row, ind_col, L.STAT_AS, block_size, want_aux=True,
                               pad_quirk=pad_quirk, **kw)
    return out, num, den


def snp_grm(X, ind_row=None, ind_col=None, block_size=None, pad_quirk=True, **kw) -> np.ndarray:
    """GRM of Weir & Goudet (2017) computed in one job (allele sharing + the scalar `mb`
    correction of R/pairwise_grm.R:42-50, both on the device)."""
    _check_fbm(X)
    if block_size is None:
        block_size = globals()["block_size"](X.nrow)
    out, _, _, _ = run_job(X, ind_row, ind_col, L.STAT_GRM, block_size, pad_quirk=pad_quirk, **kw)
    return out


def grm_from_allele_sharing(as_mat: np.ndarray, device: int = 0):
    """pairwise_grm's post-processing for a precomputed allele-sharing matrix."""
    a = np.asfortranarray(as_mat, dtype=np.float64)
    n = a.shape[0]
    if a.shape != (n, n):
        raise ValueError("allele_sharing_mat must be square")
    out = _colmajor(n)
    mb = C.c_double(0)
    eb = L.errbuf()
    rc = L.lib().tpg_grm_from_as(device, a.ctypes.data, n, out.ctypes.data, C.byref(mb), eb, len(eb))
    L.check(rc, eb)
    return out, mb.value


def snp_pairwise_many(X, stats=("king", "ibs", "allele_sharing"), ind_row=None, ind_col=None, block_size=None,
                      pad_quirk=True, devices=None, shard: str = "auto", **kw) -> dict:
    """KING, IBS and allele sharing (any subset; also "ibs_raw", "ibs_adjusted", "grm") of the same
    (X, ind_row, ind_col) from ONE pass over the genotypes: tpg_pairwise_many computes the four integer
    products once and finishes every statistic from them — what snp_king (R/snp_king.R:32), snp_ibs
    (R/snp_ibs.R:42) and snp_allele_sharing (R/snp_allele_sharing.R:33) would each decode and multiply
    for themselves.  Every matrix is identical to the one the single call returns.  block_size only
    feeds the allele-sharing pad constant (default: bigstatsr block_size(nrow), as snp_allele_sharing)."""
    global last_stats
    _check_fbm(X)
    codes = {"king": L.STAT_KING, "ibs": L.STAT_IBS_PROPORTION, "ibs_raw": L.STAT_IBS_RAW,
             "ibs_adjusted": L.STAT_IBS_ADJUSTED, "allele_sharing": L.STAT_AS, "grm": L.STAT_GRM}
    names = list(stats)
    for s_ in names:
        if s_ not in codes:
            raise ValueError(f"unknown statistic '{s_}'")
    if block_size is None:
        block_size = globals()["block_size"](X.nrow)
    if devices is not None and len(devices) >= 1:
        kw.setdefault("device", int(devices[0]))
    job, keep = make_job(X, ind_row, ind_col, codes[names[0]], block_size, pad_quirk=pad_quirk, **kw)
    n = job.n if job.ind_row else X.nrow
    if n <= 0 or (job.m if job.ind_col else X.ncol) <= 0:
        raise ValueError("ind.row and ind.col must not be empty")
    st = np.ascontiguousarray([codes[s_] for s_ in names], dtype=np.int32)
    results = (L.Result * len(names))()
    outs = {}
    for k, s_ in enumerate(names):
        o0 = _colmajor(n)
        o1 = _colmajor(n) if s_ == "ibs_raw" else None
        results[k].out0 = o0.ctypes.data
        results[k].out1 = o1.ctypes.data if o1 is not None else None
        outs[s_] = o0 if o1 is None else {"ibs": o0, "valid_n": o1}
    devs = np.ascontiguousarray(devices, dtype=np.int32) if devices is not None and len(devices) > 1 else None
    mode = {"auto": L.SHARD_NONE, "loci": L.SHARD_LOCI, "tiles": L.SHARD_TILES}[shard]
    eb = L.errbuf()
    rc = L.lib().tpg_pairwise_many(C.byref(job), st.ctypes.data, len(names), results,
                                   devs.ctypes.data if devs is not None else None,
                                   len(devs) if devs is not None else 0, mode, eb, len(eb))
    del keep
    L.check(rc, eb)
    last_stats = results[0].stats.as_dict()
    return outs
