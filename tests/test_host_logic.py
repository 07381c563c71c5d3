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


def pk2_ref(raw, nrow, ncol, r0, span):
    """numpy statement of the pk2 chunk format (csrc/pack.cpp)."""
    st = int(L.lib().tpg_pk2_stride(span))
    g = np.minimum(raw.reshape(ncol, nrow)[:, r0:r0 + span], 3).astype(np.uint8)
    pad = np.full((ncol, st * 4), 3, np.uint8)
    pad[:, :span] = g
    q = pad.reshape(ncol, st, 4)
    return (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)


@pytest.mark.parametrize("shape", [(929, 300, 0, 929), (131, 77, 5, 101), (7, 6, 0, 7), (1024, 50, 128, 512),
                                   (4097, 33, 1, 4095), (1, 9, 0, 1), (40, 3000, 2, 31)])
def test_pk2_packer_matches_numpy(shape):
    """The streamer's host-side 2-bit packer (AVX2 + scalar tails, worker pool): bytes above 2
    collapse to code 3, padding is code 3, every column starts on a 16-byte boundary."""
    nrow, ncol, r0, span = shape
    rng = np.random.default_rng(nrow * 31 + ncol)
    raw = rng.integers(0, 256, nrow * ncol).astype(np.uint8)
    raw[rng.random(raw.size) < 0.8] %= 4
    lib = L.lib()
    st = int(lib.tpg_pk2_stride(span))
    assert st % 16 == 0 and st * 4 >= span
    out = np.zeros((ncol, st), np.uint8)
    lib.tpg_pk2_pack_host(raw.ctypes.data + r0, nrow, span, ncol, out.ctypes.data)
    assert np.array_equal(out, pk2_ref(raw, nrow, ncol, r0, span))


def test_consumers_argument_checks():
    """Host-side checks of the consumers' mirror (no device needed to fail them)."""
    with pytest.raises(TypeError, match="kings_threshold"):
        tpg.filter_high_relatedness(np.eye(3))
    with pytest.raises(ValueError, match="square"):
        tpg.filter_high_relatedness(np.zeros((2, 3)), kings_threshold=0.1)
    gt = tpg.gen_tibble(TOY_3x6)
    with pytest.raises(ValueError, match="grouped gen_tibble"):
        tpg.pop_fst(gt)
    with pytest.raises(ValueError, match="should be 'WG17'"):
        tpg.indiv_inbreeding(gt, method="x")
    with pytest.raises(ValueError, match="kings_threshold must be a numeric"):
        tpg.qc_report_indiv(gt, kings_threshold="third")
    gt.population = ["a", "a", "b"]
    with pytest.raises(ValueError, match="can not be provided for grouped_df"):
        tpg.indiv_inbreeding(gt, allele_sharing_mat=np.eye(3))


def test_host_pool_serves_concurrent_callers():
    """The in-process multi-GPU driver has one streamer thread per device, all feeding the same
    worker pool (csrc/pack.cpp): concurrent jobs must neither interleave nor deadlock."""
    import threading
    lib = L.lib()
    rng = np.random.default_rng(5)
    jobs = []
    for k in range(6):
        nrow, ncol = 300 + 37 * k, 4000 + 500 * k
        raw = rng.integers(0, 5, nrow * ncol).astype(np.uint8)
        out = np.zeros((ncol, int(lib.tpg_pk2_stride(nrow))), np.uint8)
        jobs.append((raw, nrow, ncol, out))
    def work(j):
        raw, nrow, ncol, out = j
        for _ in range(5):
            lib.tpg_pk2_pack_host(raw.ctypes.data, nrow, nrow, ncol, out.ctypes.data)
    th = [threading.Thread(target=work, args=(j,)) for j in jobs]
    [t.start() for t in th]
    [t.join(60) for t in th]
    assert not any(t.is_alive() for t in th)
    for raw, nrow, ncol, out in jobs:
        assert np.array_equal(out, pk2_ref(raw, nrow, ncol, 0, nrow))


def test_swar_decode_arithmetic(tmp_path):
    """The word-parallel arithmetic of decode_planes_k4_kernel (csrc/swar.cuh) against the reference's
    per-genotype rule, every byte value in every lane position — compiled for the host with g++."""
    import subprocess
    from pathlib import Path
    src = Path(__file__).resolve().parent / "swar_check.cpp"
    exe = tmp_path / "swar_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.startswith("ok ")
