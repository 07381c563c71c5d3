"""Host-side stand-ins for the bigstatsr objects the pairwise path touches.

`FBMCode256` mirrors a bigstatsr::FBM.code256 as tidypopgen creates it
(R/gen_tibble_fbm.R:172-199): a header-less, column-major `nrow x ncol` matrix of one
unsigned byte per genotype in `<backingfile>.bk`, plus a 256-entry `code256` lookup.  The
pairwise path reads the RAW bytes and never consults `code256` (src/snp_king.cpp:35), so
`code256` is carried for API fidelity only.

`FBMDouble` mirrors bigstatsr::FBM(n, n, type = "double"): snp_ibs(type = "raw_counts")
returns two of them (R/snp_ibs.R:84-85).
"""
from __future__ import annotations

import os
import tempfile
from pathlib import Path

import numpy as np

# bigsnpr::CODE_012 as used by tidypopgen for diploids: 0, 1, 2, then NA for everything else.
CODE_012 = np.array([0.0, 1.0, 2.0] + [np.nan] * 253)


class FBMCode256:
    """A file-backed genotype matrix (column-major bytes, individuals x loci)."""

    r_class = ("FBM.code256", "FBM")

    def __init__(self, backingfile: str | os.PathLike, nrow: int, ncol: int, code256=None):
        p = Path(backingfile)
        self.backingfile = str(p if p.suffix == ".bk" else p.with_suffix(p.suffix + ".bk"))
        self.nrow = int(nrow)
        self.ncol = int(ncol)
        self.code256 = CODE_012.copy() if code256 is None else np.asarray(code256, dtype=np.float64)
        size = os.path.getsize(self.backingfile)
        if size < self.nrow * self.ncol:
            raise ValueError(f"{self.backingfile}: {size} bytes < {self.nrow} x {self.ncol}")

    @classmethod
    def from_matrix(cls, genotypes, backingfile=None, na_code: int | None = None) -> "FBMCode256":
        """genotypes: [n_indiv, n_loci] dosages with NaN / None / negative for missing.

        The NA byte is max_ploidy + 1 (= 3 for diploids), as R/gen_tibble_fbm.R:193 does."""
        g = np.asarray(genotypes, dtype=np.float64)
        miss = ~np.isfinite(g) | (g < 0)
        if na_code is None:
            na_code = int(max(2, np.nanmax(np.where(miss, 0, g)))) + 1
        raw = np.where(miss, na_code, g).astype(np.uint8)
        if backingfile is None:
            fd, backingfile = tempfile.mkstemp(suffix=".bk", prefix="tpg_fbm_")
            os.close(fd)
        p = Path(backingfile)
        if p.suffix != ".bk":
            p = p.with_suffix(p.suffix + ".bk")
        p.write_bytes(np.asfortranarray(raw).tobytes(order="F"))
        return cls(p, raw.shape[0], raw.shape[1])

    @classmethod
    def from_bytes(cls, raw_col_major: np.ndarray, nrow: int, ncol: int, backingfile=None) -> "FBMCode256":
        if backingfile is None:
            fd, backingfile = tempfile.mkstemp(suffix=".bk", prefix="tpg_fbm_")
            os.close(fd)
        p = Path(backingfile)
        if p.suffix != ".bk":
            p = p.with_suffix(p.suffix + ".bk")
        np.ascontiguousarray(raw_col_major, dtype=np.uint8).reshape(-1)[: nrow * ncol].tofile(p)
        return cls(p, nrow, ncol)

    def inherits(self, what: str) -> bool:
        return what in self.r_class

    def raw(self) -> np.ndarray:
        """Read-only memmap of the bytes as [ncol, nrow] (row k = locus k: column-major)."""
        return np.memmap(self.backingfile, dtype=np.uint8, mode="r", shape=(self.ncol, self.nrow))

    def __getitem__(self, idx):
        """X[i, j] with numpy (0-based) indexing -> decoded dosages through code256."""
        r = self.raw().T[idx]
        return self.code256[r]

    @property
    def shape(self):
        return (self.nrow, self.ncol)


class FBMDouble:
    """bigstatsr::FBM(nrow, ncol, type = "double"): a column-major file of doubles."""

    r_class = ("FBM",)

    def __init__(self, nrow: int, ncol: int, backingfile=None, init=0.0):
        if backingfile is None:
            fd, backingfile = tempfile.mkstemp(suffix=".bk", prefix="tpg_fbm_dbl_")
            os.close(fd)
        self.backingfile = str(backingfile)
        self.nrow, self.ncol = int(nrow), int(ncol)
        self._mm = np.memmap(self.backingfile, dtype=np.float64, mode="w+", shape=(self.ncol, self.nrow))
        if init is not None:
            self._mm[:] = init
        self.attrs: dict = {}

    @property
    def buffer(self) -> np.ndarray:
        """The column-major storage ([ncol, nrow] C-order view), writable."""
        return self._mm

    def __getitem__(self, idx):
        if idx == slice(None) or idx == ():
            return np.array(self._mm.T)
        return np.array(self._mm.T[idx])

    def __array__(self, dtype=None, copy=None):
        a = np.array(self._mm.T)
        return a.astype(dtype) if dtype is not None else a

    @property
    def shape(self):
        return (self.nrow, self.ncol)
