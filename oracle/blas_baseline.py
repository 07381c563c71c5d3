"""Timed CPU restatement of the reference path for bench.py's cpu_baseline / --impl reference.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle.c).  Same algorithm as oracle.c, but the
n x n products go through a threaded BLAS (MKL via torch.mm, FP64), as the reference's
Armadillo `X * Y.t()` does through the system BLAS (src/snp_king.cpp:70-72), and the decode
fills FP64 indicator matrices per loci block of R's block.size (R/snp_king.R:51-77).  The
decode here is vectorised numpy (the reference's is a serial C++ loop), so this baseline is,
if anything, generous to the CPU.  tests/test_cpu_baseline.py checks it against oracle.c.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import oracle as O


def king_blocks(raw: np.ndarray, n: int, m: int, block_size: int, max_blocks: int | None = None):
    """snp_king's block loop over the first `max_blocks` blocks of CutBySize(m, block_size).

    raw: uint8 [m, n] (column-major n x m FBM bytes).  Returns (K, N_Aa_i, loci_done, seconds)."""
    iv = O.split_len(m, block_size)
    bmax = int(iv[:, 2].max())
    K = torch.zeros((n, n), dtype=torch.float64)
    N = torch.zeros((n, n), dtype=torch.float64)
    g = [torch.zeros((bmax, n), dtype=torch.float64) for _ in range(4)]  # g0, g1, g2, valid (transposed)
    done = 0
    t0 = time.perf_counter()
    for b, (lo, up, sz) in enumerate(iv):
        if max_blocks is not None and b >= max_blocks:
            break
        blk = raw[lo - 1:up]                                  # [sz, n]
        for t in g:                                           # src/snp_king.cpp:37-40
            t.zero_()
        for v in range(3):                                    # :45-58
            g[v][:sz] = torch.from_numpy((blk == v).astype(np.float64))
        g[3][:sz] = torch.from_numpy((blk <= 2).astype(np.float64))
        g0, g1, g2, gv = g
        # :70  K += g1 g1' - 2 (g0 g2' + g2 g0')   (operands stored transposed: X' X)
        K += g1.T @ g1 - 2 * (g0.T @ g2 + g2.T @ g0)
        N += g1.T @ gv                                        # :72
        done += int(sz)
    return K.numpy(), N.numpy(), done, time.perf_counter() - t0


def king_finish(K: np.ndarray, N: np.ndarray) -> np.ndarray:
    """divide_king_sub, R/snp_king.R:82-91."""
    Ni, Nj = N, N.T
    with np.errstate(all="ignore"):
        pm = np.minimum(Ni, Nj)
        return K / (2 * pm) + 0.5 - 0.25 * (Ni + Nj) / pm
