import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture
def tpg_cfg():
    """tidypopgen_b200.config (the process-wide tpg_config), restored to its defaults after the test."""
    from tidypopgen_b200 import _lib as L
    yield L.config
    L.config(reset=True)


@pytest.fixture(scope="session")
def families():
    """12 x 961 reference fixture (inst/extdata/related/families.bed) as FBM bytes + golden outputs."""
    raw = np.fromfile(GOLDEN / "families.bk", dtype=np.uint8)
    return {
        "raw": raw, "n": 12, "m": 961, "path": str(GOLDEN / "families.bk"),
        "kin0": json.loads((GOLDEN / "families_kin0.json").read_text()),
        "mibs": np.array(json.loads((GOLDEN / "families_mibs.json").read_text())),
    }


@pytest.fixture(scope="session")
def example_gt():
    """7 x 6 load_example_gt() matrix (inst/extdata/bigsnp_obj_gt/old_gen_tibble.bk)."""
    raw = np.fromfile(GOLDEN / "example_gt.bk", dtype=np.uint8)
    return {"raw": raw, "n": 7, "m": 6, "path": str(GOLDEN / "example_gt.bk")}


# the toy matrices of the reference's testthat files
TOY_3x6 = np.array([[1, 1, 0, 1, 1, 0], [2, 1, 0, 0, 0, 0], [2, 2, 0, 0, 1, 1]], dtype=float)  # test_snp_king.R:6-10
TOY_3x6_NA = np.array([[1, 1, 0, 1, 1, 0], [2, 1, 0, np.nan, 0, 0], [2, np.nan, 0, 0, 1, 1]])  # test_pairwise_allele_sharing.R:3-7
EXAMPLE_7x6 = np.array([  # R/load_example_gt.R:24-32, test_pairwise_grm.R:3-11
    [1, 1, 0, 1, 1, 0], [2, 1, 0, np.nan, 0, 0], [2, np.nan, 0, 0, 1, 1], [1, 0, 0, 1, 0, 0],
    [1, 2, 0, 1, 2, 1], [0, 0, 0, 0, np.nan, 1], [0, 1, 1, 0, 1, np.nan]])


def to_bk(genotypes: np.ndarray, na_code: int = 3) -> np.ndarray:
    """[n, m] dosages (NaN = NA) -> column-major FBM bytes."""
    g = np.where(np.isnan(genotypes), na_code, genotypes).astype(np.uint8)
    return np.ascontiguousarray(g.T).reshape(-1)


def same(a, b):
    """Bit-for-bit equality of doubles up to NaN payloads (R's identical())."""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))
