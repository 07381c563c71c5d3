"""Host-side mirror of the reference interface: FBM bytes, gen_tibble subsetting, tidy order,
argument checking.  No GPU."""
import numpy as np
import pytest

import tidypopgen_b200 as tpg
from conftest import EXAMPLE_7x6, TOY_3x6, to_bk
from tidypopgen_b200 import _lib as L
from tidypopgen_b200 import snp


def test_fbm_from_matrix_layout(tmp_path, example_gt):
    X = tpg.FBMCode256.from_matrix(EXAMPLE_7x6, tmp_path / "ex")
    raw = np.fromfile(X.backingfile, dtype=np.uint8)
    # identical to the .bk the reference ships for load_example_gt(): column-major, NA = 3
    assert np.array_equal(raw, example_gt["raw"])
    assert X.shape == (7, 6) and X.inherits("FBM.code256")
    assert np.isnan(X[1, 3]) and X[4, 1] == 2


def test_fbm_rejects_short_file(tmp_path):
    p = tmp_path / "short.bk"
    p.write_bytes(b"\x00" * 10)
    with pytest.raises(ValueError):
        tpg.FBMCode256(p, 3, 4)


def test_snp_king_checks_class():
    # R/snp_king.R:39-41
    with pytest.raises(TypeError, match="X should be a FBM.code256 matrix"):
        tpg.snp_king(np.zeros((3, 3)))


def test_snp_ibs_match_arg():
    X = tpg.FBMCode256.from_matrix(TOY_3x6)
    with pytest.raises(ValueError, match="should be one of"):
        tpg.snp_ibs(X, type="counts")


def test_tidy_dist_matrix_order():
    # combn order (a,b),(a,c),(b,c) — test_snp_ibs.R:52-56
    m = np.arange(9.0).reshape(3, 3)
    t = tpg.tidy_dist_matrix(m, ["a", "b", "c"])
    assert t.item1 == ["a", "a", "b"] and t.item2 == ["b", "c", "c"]
    assert t.value.tolist() == [m[0, 1], m[0, 2], m[1, 2]]
    assert "pairwise_tbl" in t.r_class


def test_gen_tibble_subsetting():
    gt = tpg.gen_tibble(TOY_3x6, indiv_id=["a", "b", "c"])
    sub = gt[[0, 2]].select_loci([1, 2, 4, 5])  # test_snp_ibs.R:58-62 (1-based 2,3,5,6)
    assert sub.id == ["a", "c"]
    assert sub.row_index.tolist() == [1, 3] and sub.big_index.tolist() == [2, 3, 5, 6]
    assert gt.big_index.tolist() == [1, 2, 3, 4, 5, 6]
    assert gt[np.array([True, False, True])].id == ["a", "c"]


def test_job_marshalling():
    X = tpg.FBMCode256.from_matrix(TOY_3x6)
    job, keep = snp.make_job(X, [1, 3], [2, 3, 5, 6], L.STAT_IBS_RAW, 3)
    assert (job.n, job.m, job.nrow_total, job.ncol_total) == (2, 4, 3, 6)
    assert job.src_kind == L.SRC_FILE and job.bk_path.endswith(b".bk") and job.pad_quirk == 1
    job2, _ = snp.make_job(X, None, None, L.STAT_KING, 10)
    assert job2.ind_row is None and job2.n == 3 and job2.m == 6
    with pytest.raises(ValueError):
        snp.make_job(X, [1.5], None, L.STAT_KING, 10)


def test_cut_by_size_and_defaults():
    assert snp.cut_by_size(961, 300)[:, 2].tolist() == [240, 240, 241, 240]
    assert tpg.block_size(929) == 144475
