"""ctypes wrapper of oracle/oracle.c — TEST INFRASTRUCTURE ONLY (see the header of oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "liboracle.so"
_lib = None


def build() -> Path:
    r = subprocess.run(["make", "-C", str(_HERE)], capture_output=True, text=True)
    if r.returncode != 0 or not _SO.exists():
        raise RuntimeError("building the oracle failed:\n" + r.stdout + r.stderr)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        src = _HERE / "oracle.c"
        if not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
            build()
        L = C.CDLL(str(_SO))
        p, i64 = C.c_void_p, C.c_int64
        L.orc_block_size.restype = i64
        L.orc_block_size.argtypes = [i64]
        L.orc_split_len.restype = i64
        L.orc_split_len.argtypes = [i64, i64, p, p, p, i64]
        common = [p, i64, p, i64, p, i64, i64]
        L.orc_snp_king.restype = C.c_int
        L.orc_snp_king.argtypes = common + [p, p, p]
        L.orc_snp_ibs.restype = C.c_int
        L.orc_snp_ibs.argtypes = common + [C.c_int, p, p, p]
        L.orc_snp_as.restype = C.c_int
        L.orc_snp_as.argtypes = common + [p, p, p]
        L.orc_grm.restype = C.c_int
        L.orc_grm.argtypes = [p, i64, p, p]
        L.orc_filter_high_relatedness.restype = C.c_int
        L.orc_filter_high_relatedness.argtypes = [p, i64, C.c_double, p, p, p]
        L.orc_indiv_inbreeding.restype = C.c_int
        L.orc_indiv_inbreeding.argtypes = [p, i64, p, p]
        L.orc_pop_fst.restype = C.c_int
        L.orc_pop_fst.argtypes = [p, i64, p, i64, p]
        L.orc_pop_fis_wg17.restype = C.c_int
        L.orc_pop_fis_wg17.argtypes = [p, i64, p, i64, p]
        L.orc_gt_ind_hetero.restype = C.c_int
        L.orc_gt_ind_hetero.argtypes = [p, i64, p, i64, p, i64, p, p, p]
        L.orc_alt_freq.restype = C.c_int
        L.orc_alt_freq.argtypes = [p, i64, p, i64, p, i64, p, p, C.c_int, p, p]
        _lib = L
    return _lib


NA_BITS_LOW = 1954  # R_IsNA: a NaN whose low 32-bit word is 1954


def is_na(x: np.ndarray) -> np.ndarray:
    """R's is.na() && !is.nan(): NaN carrying R's NA payload."""
    u = np.ascontiguousarray(x, dtype=np.float64).view(np.uint64)
    return np.isnan(x) & ((u & np.uint64(0xFFFFFFFF)) == np.uint64(NA_BITS_LOW))


def block_size(n: int) -> int:
    return int(lib().orc_block_size(int(n)))


def split_len(m: int, block_size_: int) -> np.ndarray:
    """CutBySize(m, block.size): rows of (lower, upper, size), 1-based inclusive."""
    nb = int(lib().orc_split_len(int(m), int(block_size_), None, None, None, 0))
    if nb < 0:
        raise ValueError("bad arguments")
    lo = np.zeros(nb, np.int64)
    up = np.zeros(nb, np.int64)
    sz = np.zeros(nb, np.int64)
    lib().orc_split_len(int(m), int(block_size_), lo.ctypes.data, up.ctypes.data, sz.ctypes.data, nb)
    return np.stack([lo, up, sz], axis=1)


def n_short(m: int, block_size_: int) -> int:
    iv = split_len(m, block_size_)
    return int((iv[:, 2] < iv[:, 2].max()).sum())


def _prep(bk, nrow_total, ncol_total, ind_row, ind_col):
    bk = np.ascontiguousarray(bk, dtype=np.uint8).reshape(-1)
    assert bk.size >= nrow_total * ncol_total
    rows = np.arange(1, nrow_total + 1, dtype=np.int32) if ind_row is None else np.ascontiguousarray(ind_row, np.int32)
    cols = np.arange(1, ncol_total + 1, dtype=np.int32) if ind_col is None else np.ascontiguousarray(ind_col, np.int32)
    assert rows.min() >= 1 and rows.max() <= nrow_total and cols.min() >= 1 and cols.max() <= ncol_total
    return bk, rows, cols


def _mat(n):
    # column-major n x n, exposed to numpy as [i, j]
    return np.zeros((n, n), np.float64, order="F")


def snp_king(bk, nrow_total, ncol_total, ind_row=None, ind_col=None, block_size_=None):
    """Returns dict(king, K, N_Aa_i) — R/snp_king.R:32-103 (default block.size = block_size(nrow)*4)."""
    bk, rows, cols = _prep(bk, nrow_total, ncol_total, ind_row, ind_col)
    n, m = len(rows), len(cols)
    bs = block_size(nrow_total) * 4 if block_size_ is None else int(block_size_)
    king, K, N = _mat(n), _mat(n), _mat(n)
    rc = lib().orc_snp_king(bk.ctypes.data, nrow_total, rows.ctypes.data, n, cols.ctypes.data, m, bs,
                            king.ctypes.data, K.ctypes.data, N.ctypes.data)
    assert rc == 0
    return {"king": king, "K": K, "N_Aa_i": N}


def snp_ibs(bk, nrow_total, ncol_total, ind_row=None, ind_col=None, type="proportion", block_size_=None):
    """R/snp_ibs.R:42-104.  Returns dict(out, ibs, valid_n)."""
    bk, rows, cols = _prep(bk, nrow_total, ncol_total, ind_row, ind_col)
    n, m = len(rows), len(cols)
    bs = block_size(nrow_total) if block_size_ is None else int(block_size_)
    t = {"proportion": 0, "adjusted_counts": 1, "raw_counts": 2}[type]
    out, ibs, valid = _mat(n), _mat(n), _mat(n)
    rc = lib().orc_snp_ibs(bk.ctypes.data, nrow_total, rows.ctypes.data, n, cols.ctypes.data, m, bs, t,
                           out.ctypes.data, ibs.ctypes.data, valid.ctypes.data)
    assert rc == 0
    return {"out": out, "ibs": ibs, "valid_n": valid}


def snp_allele_sharing(bk, nrow_total, ncol_total, ind_row=None, ind_col=None, block_size_=None):
    """R/snp_allele_sharing.R:33-82 (with the src/snp_as.cpp:56-62 padding quirk)."""
    bk, rows, cols = _prep(bk, nrow_total, ncol_total, ind_row, ind_col)
    n, m = len(rows), len(cols)
    bs = block_size(nrow_total) if block_size_ is None else int(block_size_)
    as_, num, den = _mat(n), _mat(n), _mat(n)
    rc = lib().orc_snp_as(bk.ctypes.data, nrow_total, rows.ctypes.data, n, cols.ctypes.data, m, bs,
                          as_.ctypes.data, num.ctypes.data, den.ctypes.data)
    assert rc == 0, "myassert(m2 == m + 1) failed"
    return {"as": as_, "num": num, "denom": den}


def grm_from_as(as_):
    """R/pairwise_grm.R:42-50."""
    a = np.asfortranarray(as_, dtype=np.float64)
    n = a.shape[0]
    out = _mat(n)
    mb = C.c_double(0)
    lib().orc_grm(a.ctypes.data, n, out.ctypes.data, C.byref(mb))
    return out, mb.value


# ------------------------------------------------------------------ consumers

def filter_high_relatedness(mat, kings_threshold):
    """R/filter_high_relatedness.R:26-145.  Returns dict(keep [original order], order [0-based
    visiting order], deleted_rank, error) — error=True where R stops on a NaN comparison."""
    a = np.asfortranarray(mat, dtype=np.float64)
    n = a.shape[0]
    keep, order, dele = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    rc = lib().orc_filter_high_relatedness(a.ctypes.data, n, float(kings_threshold), keep.ctypes.data,
                                           order.ctypes.data, dele.ctypes.data)
    return {"keep": keep.astype(bool), "order": order, "deleted_rank": dele.astype(bool), "error": rc != 0}


def indiv_inbreeding(as_):
    """R/indiv_inbreeding.R:56-81."""
    a = np.asfortranarray(as_, dtype=np.float64)
    n = a.shape[0]
    out = np.zeros(n)
    mb = C.c_double(0)
    lib().orc_indiv_inbreeding(a.ctypes.data, n, out.ctypes.data, C.byref(mb))
    return out, mb.value


def pop_fst(as_, group, npop):
    """R/pop_fst.R:31-76: npop population values + the global one."""
    a = np.asfortranarray(as_, dtype=np.float64)
    g = np.ascontiguousarray(group, np.int32)
    out = np.zeros(npop + 1)
    lib().orc_pop_fst(a.ctypes.data, a.shape[0], g.ctypes.data, npop, out.ctypes.data)
    return out


def pop_fis_wg17(as_, group, npop):
    """R/pop_fis.R:137-191."""
    a = np.asfortranarray(as_, dtype=np.float64)
    g = np.ascontiguousarray(group, np.int32)
    out = np.zeros(npop + 1)
    lib().orc_pop_fis_wg17(a.ctypes.data, a.shape[0], g.ctypes.data, npop, out.ctypes.data)
    return out


CODE_012 = np.array([0.0, 1.0, 2.0] + [np.nan] * 253)                       # bigsnpr::CODE_012
CODE_IMPUTE_PRED = np.array([0.0, 1.0, 2.0, np.nan, 0.0, 1.0, 2.0] + [np.nan] * 249)  # bigsnpr::CODE_IMPUTE_PRED


def gt_ind_hetero(bk, nrow_total, ncol_total, ind_row=None, ind_col=None, code256=CODE_012):
    """src/gt_ind_hetero.cpp:12-42: (het_n, na_n) per individual through code256."""
    bk, rows, cols = _prep(bk, nrow_total, ncol_total, ind_row, ind_col)
    n, m = len(rows), len(cols)
    code = np.ascontiguousarray(code256, np.float64)
    het, na = np.zeros(n, np.int32), np.zeros(n, np.int32)
    lib().orc_gt_ind_hetero(bk.ctypes.data, nrow_total, rows.ctypes.data, n, cols.ctypes.data, m, code.ctypes.data,
                            het.ctypes.data, na.ctypes.data)
    return het, na


def alt_freq(bk, nrow_total, ncol_total, ind_row=None, ind_col=None, code256=CODE_012, ploidy=None, as_counts=False):
    """src/alt_freq_dip_pseudo_cpp.cpp:8-58: (freq or n_alt, n_valid) per locus."""
    bk, rows, cols = _prep(bk, nrow_total, ncol_total, ind_row, ind_col)
    n, m = len(rows), len(cols)
    code = np.ascontiguousarray(code256, np.float64)
    pl = np.full(n, 2.0) if ploidy is None else np.ascontiguousarray(ploidy, np.float64)
    o0, o1 = np.zeros(m), np.zeros(m)
    lib().orc_alt_freq(bk.ctypes.data, nrow_total, rows.ctypes.data, n, cols.ctypes.data, m, code.ctypes.data,
                       pl.ctypes.data, int(as_counts), o0.ctypes.data, o1.ctypes.data)
    return o0, o1
