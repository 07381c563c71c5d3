"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle and
the reference's golden vectors.  Integer counts and the FP64 statistics must be bit-exact
(NaN payloads aside); GRM within 1e-10 absolute (north_star) — in practice ~1e-15."""
import ctypes as C

import numpy as np
import pytest

import tidypopgen_b200 as tpg
from conftest import EXAMPLE_7x6, TOY_3x6, TOY_3x6_NA, same
from oracle import oracle as O
from tidypopgen_b200 import _lib as L

pytestmark = pytest.mark.gpu

GRM_ATOL = 1e-10


def fbm_of(raw, n, m):
    return tpg.FBMCode256.from_bytes(raw, n, m)


# ------------------------------------------------------------------ the reference's own tests

def test_snp_king_toy_and_blocks():
    # tests/testthat/test_snp_king.R:50-64
    gt = tpg.gen_tibble(TOY_3x6, indiv_id=["a", "b", "c"])
    king = tpg.snp_king(gt.fbm)
    ref = O.snp_king(np.fromfile(gt.fbm.backingfile, np.uint8), 3, 6)["king"]
    assert same(king, ref)
    assert same(tpg.snp_king(gt.fbm, block_size=3), king)
    pk = tpg.pairwise_king(gt, as_matrix=True)
    assert same(np.asarray(pk), king) and pk.names == ["a", "b", "c"] and "pairwise_matrix" in pk.r_class
    t = tpg.pairwise_king(gt)
    assert t.item1 == ["a", "a", "b"] and same(t.value, king[[0, 0, 1], [1, 2, 2]])


def test_snp_king_families_with_missing(families):
    # test_snp_king.R:66-81
    f = families
    X = tpg.FBMCode256(f["path"], f["n"], f["m"])
    ref = O.snp_king(f["raw"], f["n"], f["m"])
    king, K, N = tpg.snp_king_counts(X)
    assert np.array_equal(K, ref["K"]) and np.array_equal(N, ref["N_Aa_i"])
    assert same(king, ref["king"])
    assert same(tpg.snp_king(X, block_size=300), king)
    assert K[0, 1] == -32 and N[0, 1] == 326 and N[1, 0] == 309


def test_snp_king_vs_king_binary(families):
    # test_snp_king.R:84-131
    f = families
    king = tpg.snp_king(tpg.FBMCode256(f["path"], f["n"], f["m"]))
    for i, j, v in f["kin0"]:
        assert abs(king[i - 1, j - 1] - v) < 6e-5 and king[i - 1, j - 1] == king[j - 1, i - 1]
    assert np.array_equal(np.diag(king), np.full(12, 0.5))


def test_snp_ibs_toy_blocks_and_subsets():
    # test_snp_ibs.R:28-65
    gt = tpg.gen_tibble(TOY_3x6, indiv_id=["a", "b", "c"])
    r = tpg.snp_ibs(gt.fbm, type="raw_counts")
    assert r["ibs"][:][0, 1] == 9
    r2 = tpg.snp_ibs(gt.fbm, block_size=3, type="raw_counts")
    assert np.array_equal(r2["ibs"][:], r["ibs"][:])
    pg = tpg.pairwise_ibs(gt, type="raw_counts")
    assert np.array_equal(pg["ibs"][:], r["ibs"][:]) and pg["ibs"].attrs["indiv_names"] == ["a", "b", "c"]
    prop = tpg.pairwise_ibs(gt)
    by_hand = r["ibs"][:] / r["valid_n"][:]
    assert same(prop.value, by_hand[[0, 0, 1], [1, 2, 2]])
    sub = gt[[0, 2]].select_loci([1, 2, 4, 5])
    assert tpg.pairwise_ibs(sub, type="raw_counts")["ibs"][:][0, 1] == 6


def test_snp_ibs_vs_plink(families):
    # test_snp_ibs.R:69-104
    f = families
    X = tpg.FBMCode256(f["path"], f["n"], f["m"])
    ibs = tpg.snp_ibs(X)
    assert np.array_equal(np.round(ibs, 6), f["mibs"])
    assert same(tpg.snp_ibs(X, type="adjusted_counts"), O.snp_ibs(f["raw"], 12, 961, type="adjusted_counts")["out"])


def test_allele_sharing_toy():
    # test_pairwise_allele_sharing.R:29-46
    gt = tpg.gen_tibble(TOY_3x6_NA, indiv_id=["a", "b", "c"])
    raw = np.fromfile(gt.fbm.backingfile, np.uint8)
    a = tpg.snp_allele_sharing(gt.fbm)
    assert same(a, O.snp_allele_sharing(raw, 3, 6)["as"])
    assert same(tpg.snp_allele_sharing(gt.fbm, block_size=3), a)
    assert same(np.asarray(tpg.pairwise_allele_sharing(gt, as_matrix=True)), a)


def test_pairwise_grm_example(example_gt):
    # test_pairwise_grm.R:33-40
    e = example_gt
    gt = tpg.gen_tibble(tpg.FBMCode256(e["path"], 7, 6), indiv_id=list("abcdefg"))
    ref, mb = O.grm_from_as(O.snp_allele_sharing(e["raw"], 7, 6)["as"])
    g = tpg.pairwise_grm(gt)
    assert np.allclose(np.asarray(g), ref, rtol=0, atol=GRM_ATOL)
    assert abs(g[0, 0] - 0.396947) < 1e-6 and abs(g[1, 1] - 1.519084) < 1e-6
    as_m = tpg.pairwise_allele_sharing(gt, as_matrix=True)
    g2 = tpg.pairwise_grm(gt, allele_sharing_mat=as_m)
    assert np.allclose(np.asarray(g2), ref, rtol=0, atol=GRM_ATOL)


def test_example_gt_nan_pattern(example_gt):
    # test_autoplot_pairwise.R:17-24 pins 0/0 -> NaN in KING on the example
    e = example_gt
    X = tpg.FBMCode256(e["path"], 7, 6)
    king, K, N = tpg.snp_king_counts(X)
    ref = O.snp_king(e["raw"], 7, 6)
    assert same(king, ref["king"]) and np.isnan(king[1, 2]) and np.isnan(king[5, 6]) and king[1, 4] == -1.25
    assert K[0].tolist() == [4, 1, 1, 2, 2, 0, 2] and N[0].tolist() == [4, 3, 3, 4, 4, 3, 4]
    r = tpg.snp_ibs(X, type="raw_counts")
    assert r["ibs"][:][0].tolist() == [12, 8, 7, 10, 9, 6, 7]
    assert r["valid_n"][:][0].tolist() == [12, 10, 10, 12, 12, 10, 10]


# ------------------------------------------------------------------ randomized parity vs the oracle

def random_bk(rng, n, m, p_missing=0.05, codes=(0, 1, 2, 3)):
    p = np.array([0.45, 0.3, 0.25]) * (1 - p_missing)
    extra = len(codes) - 3
    probs = np.concatenate([p, np.full(extra, p_missing / extra)])
    return rng.choice(np.array(codes, np.uint8), size=(m, n), p=probs).reshape(-1)


CASES = [
    # n, m, block_size, gpu_chunk_loci, rows, cols
    (1, 5, None, 0, "all", "all"),
    (2, 1, None, 0, "all", "all"),
    (33, 257, 50, 128, "all", "all"),
    (129, 1000, 333, 256, "all", "all"),
    (257, 700, None, 0, "all", "all"),          # second row tile, ragged
    (300, 2000, 777, 512, "shuffle", "shuffle"),
    (130, 900, 100, 0, "subset", "subset"),
    (64, 1500, 1, 384, "dup", "every3"),        # duplicated rows, block.size = 1
    (513, 300, None, 0, "window", "window"),
]


def pick(kind, total, rng):
    if kind == "all":
        return None
    idx = np.arange(1, total + 1)
    if kind == "shuffle":
        return rng.permutation(idx).astype(np.int32)
    if kind == "subset":
        return np.sort(rng.choice(idx, size=max(1, total * 7 // 10), replace=False)).astype(np.int32)[::-1].copy()
    if kind == "dup":
        return rng.choice(idx, size=total, replace=True).astype(np.int32)
    if kind == "every3":
        tail = rng.permutation(idx)[: max(1, total // 20)]
        return np.concatenate([idx[::3], tail]).astype(np.int32)
    if kind == "window":
        lo = total // 5
        return idx[lo: lo + max(1, total // 2)].astype(np.int32)
    raise ValueError(kind)


@pytest.mark.parametrize("case", CASES, ids=[f"n{c[0]}_m{c[1]}_{c[4]}" for c in CASES])
@pytest.mark.parametrize("variant", [0, L.FLAG_GRAM_I8, L.FLAG_PLANES_K_MAJOR, L.FLAG_GRAM_1CTA, L.FLAG_PLANES_K_MAJOR | L.FLAG_GRAM_1CTA,
                                     L.FLAG_NO_PACK, L.FLAG_NO_PACK | L.FLAG_GRAM_I8],
                         ids=["pk2_mxf4", "pk2_i8", "kmajor", "mn_1cta", "kmajor_1cta", "raw_mxf4", "raw_i8"])
def test_random_parity_all_stats(case, variant):
    n_tot, m_tot, bs, chunk, rk, ck = case
    rng = np.random.default_rng(n_tot * 1000 + m_tot)
    raw = random_bk(rng, n_tot, m_tot, 0.08, codes=(0, 1, 2, 3, 4, 5, 6, 255))
    X = fbm_of(raw, n_tot, m_tot)
    rows, cols = pick(rk, n_tot, rng), pick(ck, m_tot, rng)
    kw = dict(gpu_chunk_loci=chunk, flags=variant)
    ref_k = O.snp_king(raw, n_tot, m_tot, rows, cols, block_size_=bs)
    king, K, N = tpg.snp_king_counts(X, rows, cols, block_size=bs, **kw)
    assert np.array_equal(K, ref_k["K"]), "KING numerator"
    assert np.array_equal(N, ref_k["N_Aa_i"]), "N_Aa_i"
    assert same(king, ref_k["king"]), "KING"
    for typ in ("raw_counts", "proportion", "adjusted_counts"):
        ref_i = O.snp_ibs(raw, n_tot, m_tot, rows, cols, type=typ, block_size_=bs)
        got = tpg.snp_ibs(X, rows, cols, type=typ, block_size=bs, **kw)
        if typ == "raw_counts":
            assert np.array_equal(got["ibs"][:], ref_i["ibs"]) and np.array_equal(got["valid_n"][:], ref_i["valid_n"])
        else:
            assert same(got, ref_i["out"]), typ
    ref_a = O.snp_allele_sharing(raw, n_tot, m_tot, rows, cols, block_size_=bs)
    a, num, den = tpg.snp_allele_sharing_counts(X, rows, cols, block_size=bs, **kw)
    assert np.array_equal(num, ref_a["num"]) and np.array_equal(den, ref_a["denom"])
    assert same(a, ref_a["as"]) and np.array_equal(O.is_na(a), O.is_na(ref_a["as"]))
    g = tpg.snp_grm(X, rows, cols, block_size=bs, **kw)
    ref_g, _ = O.grm_from_as(ref_a["as"])
    ok = ~np.isnan(ref_g)
    assert np.array_equal(np.isnan(g), ~ok) and np.allclose(g[ok], ref_g[ok], rtol=0, atol=GRM_ATOL)


# builds of the 4-bit decode kernel: the default and the round's first one (decode.cu)
K4_BUILDS = {"k4": {}, "k4_v1": {"decode_k4_v1": 1}}


@pytest.mark.parametrize("build", list(K4_BUILDS))
def test_k4_builds_whole_job(build, tpg_cfg):
    """Every build of the 4-bit decode kernel through whole KING / allele-sharing jobs (3 and 2 planes)."""
    tpg_cfg(**K4_BUILDS[build])
    rng = np.random.default_rng(11)
    n, m = 389, 3000
    raw = random_bk(rng, n, m, 0.08, codes=(0, 1, 2, 3, 4, 5, 6, 255))
    X = fbm_of(raw, n, m)
    for rows, cols in ((None, None), (pick("shuffle", n, rng), pick("every3", m, rng))):
        ref_k = O.snp_king(raw, n, m, rows, cols)
        king, K, N = tpg.snp_king_counts(X, rows, cols, flags=L.FLAG_NO_PACK, gpu_chunk_loci=1024)
        assert tpg.snp.last_stats["n_mxf4_chunks"] >= 1
        assert np.array_equal(K, ref_k["K"]) and np.array_equal(N, ref_k["N_Aa_i"]) and same(king, ref_k["king"])
        ref_a = O.snp_allele_sharing(raw, n, m, rows, cols)
        a, num, den = tpg.snp_allele_sharing_counts(X, rows, cols, flags=L.FLAG_NO_PACK)
        assert np.array_equal(num, ref_a["num"]) and np.array_equal(den, ref_a["denom"]) and same(a, ref_a["as"])


@pytest.mark.parametrize("compact", [True, False], ids=["gather_pack", "gather_in_decode"])
@pytest.mark.parametrize("rk,ck", [("shuffle", "every3"), ("dup", "all"), ("subset", "shuffle")])
def test_device_source_gathered_rows(rk, ck, compact, tpg_cfg):
    """Device-resident bytes with gathered rows: the chunk is gathered through shared memory into a dense
    2-bit packed chunk (gather_pack_kernel) and then decoded on the contiguous packed-input path;
    tpg_config.no_compact gathers inside the decode kernel instead.  Both exact."""
    import torch
    if not compact:
        tpg_cfg(no_compact=1)
    n_tot, m_tot = 1500, 2100
    rng = np.random.default_rng(77)
    raw = random_bk(rng, n_tot, m_tot, 0.08, codes=(0, 1, 2, 3, 4, 5, 6, 255))
    dev = torch.from_numpy(raw).cuda()
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<device>", n_tot, m_tot, None
    src = (L.SRC_DEVICE, dev.data_ptr(), dev)
    rows, cols = pick(rk, n_tot, rng), pick(ck, m_tot, rng)
    for chunk in (0, 512):   # one chunk / several (the last one short)
        ref_k = O.snp_king(raw, n_tot, m_tot, rows, cols)
        king, K, N = tpg.snp_king_counts(X, rows, cols, src=src, gpu_chunk_loci=chunk)
        assert tpg.snp.last_stats["n_mxf4_chunks"] >= 1
        assert np.array_equal(K, ref_k["K"]) and np.array_equal(N, ref_k["N_Aa_i"]) and same(king, ref_k["king"])
    ref_a = O.snp_allele_sharing(raw, n_tot, m_tot, rows, cols)
    a, num, den = tpg.snp_allele_sharing_counts(X, rows, cols, src=src)
    assert np.array_equal(num, ref_a["num"]) and np.array_equal(den, ref_a["denom"]) and same(a, ref_a["as"])


def test_pad_quirk_switch():
    rng = np.random.default_rng(5)
    raw = random_bk(rng, 50, 1003)
    X = fbm_of(raw, 50, 1003)
    assert O.n_short(1003, 100) > 0
    with_q, num_q, _ = tpg.snp_allele_sharing_counts(X, block_size=100)
    clean, num_c, _ = tpg.snp_allele_sharing_counts(X, block_size=100, pad_quirk=False)
    assert np.array_equal(num_q, num_c + O.n_short(1003, 100))
    assert same(with_q, O.snp_allele_sharing(raw, 50, 1003, block_size_=100)["as"])
    assert same(clean, O.snp_allele_sharing(raw, 50, 1003, block_size_=1003)["as"])  # one block: no quirk


def test_all_missing_and_haploid_na_code():
    # every genotype missing -> KING NaN, IBS NaN, AS NA_real_
    raw = np.full(20 * 10, 3, np.uint8)
    X = fbm_of(raw, 10, 20)
    assert np.all(np.isnan(tpg.snp_king(X))) and np.all(np.isnan(tpg.snp_ibs(X)))
    assert np.all(O.is_na(tpg.snp_allele_sharing(X)))
    # all-haploid FBM: NA byte is 2 (max_ploidy + 1) and this path counts it as hom-alt
    rng = np.random.default_rng(9)
    raw = rng.choice(np.array([0, 1, 2], np.uint8), size=40 * 9).astype(np.uint8)
    assert same(tpg.snp_king(fbm_of(raw, 9, 40)), O.snp_king(raw, 9, 40)["king"])


def test_sources_file_host_device_agree(families):
    import torch
    f = families
    X = tpg.FBMCode256(f["path"], f["n"], f["m"])
    ref = O.snp_king(f["raw"], 12, 961)["king"]
    a = tpg.snp_king(X)
    host = np.array(f["raw"])
    b = tpg.snp_king(X, src=(L.SRC_HOST, host.ctypes.data, host))
    dev = torch.from_numpy(host).cuda()
    c = tpg.snp_king(X, src=(L.SRC_DEVICE, dev.data_ptr(), dev), ind_col=np.arange(961, 0, -1))
    pinned = torch.from_numpy(host).pin_memory()
    d = tpg.snp_king(X, src=(L.SRC_HOST, pinned.data_ptr(), pinned))
    assert same(a, ref) and same(b, ref) and same(c, ref) and same(d, ref)


def test_errors():
    X = tpg.FBMCode256.from_matrix(TOY_3x6)
    with pytest.raises(L.TpgError, match="out of bounds") as e:
        tpg.snp_king(X, ind_row=[1, 4])
    assert e.value.code == L.ERR_ARG
    with pytest.raises(L.TpgError, match="out of bounds"):
        tpg.snp_ibs(X, ind_col=[0, 1])
    with pytest.raises(ValueError):
        tpg.snp_king(X, ind_row=[])
    X.backingfile = "/nonexistent/file.bk"
    with pytest.raises(L.TpgError) as e:
        tpg.snp_king(X)
    assert e.value.code == L.ERR_IO


# ------------------------------------------------------------------ kernel-level parity

def planes_ref(g):
    return ((g == 2).astype(np.int8) - (g == 0).astype(np.int8), (g <= 2).astype(np.int8), (g == 1).astype(np.int8))


@pytest.mark.parametrize("k_major", [False, True])
@pytest.mark.parametrize("gather", [False, True])
def test_decode_planes_kernel(k_major, gather):
    import torch
    ctx = L.Context(0)
    rng = np.random.default_rng(1)
    nrow, ncol = 929, 700
    raw = rng.integers(0, 256, size=(ncol, nrow)).astype(np.uint8)
    raw[rng.random(raw.shape) < 0.8] %= 4
    if gather:
        rows = rng.permutation(nrow)[:613].astype(np.int32)
        cols = rng.permutation(ncol)[:391].astype(np.int32)
    else:
        rows = np.arange(5, 5 + 801, dtype=np.int32)
        cols = np.arange(3, 3 + 515, dtype=np.int32)
    n, lc = len(rows), len(cols)
    n_pad, lc_pad = (n + 255) // 256 * 256, (lc + 127) // 128 * 128
    d_raw = torch.from_numpy(raw.reshape(-1)).cuda()
    d_rows, d_cols = torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda()
    ps = n_pad * lc_pad
    d_planes = torch.full((3 * ps,), 77, dtype=torch.int8, device="cuda")
    torch.cuda.synchronize()  # the library runs on its own non-blocking stream
    eb = L.errbuf()
    L.check(L.lib().tpg_decode_planes(ctx.handle, d_raw.data_ptr(), raw.size, nrow,
                                      d_rows.data_ptr() if gather else None, int(rows[0]), n, n_pad,
                                      d_cols.data_ptr() if gather else None, int(cols[0]), lc, lc_pad,
                                      d_planes.data_ptr(), ps, int(k_major), eb, len(eb)), eb)
    got = d_planes.cpu().numpy().reshape(3, -1)
    ref = np.zeros((3, n_pad, lc_pad), np.int8)
    for p, x in enumerate(planes_ref(raw[np.ix_(cols, rows)].T)):
        ref[p, :n, :lc] = x
    got = got.reshape(3, n_pad, lc_pad) if k_major else got.reshape(3, lc_pad, n_pad).transpose(0, 2, 1)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("ctas", [2, 1, "wide"])
@pytest.mark.parametrize("k_major", [False, True])
@pytest.mark.parametrize("shape", [(256, 128), (512, 1152), (1024, 4096), (2560, 2048)])
def test_gram_kernel(k_major, shape, ctas, tpg_cfg):
    """The int8 -> s32 Gram kernels: CTA pairs (256 x 256 ping-pong), single CTA, and the wide 256 x 512 kernel with the
    TMA-reduction epilogue; MN-major and K-major planes."""
    import torch
    if ctas == "wide":
        tpg_cfg(gram_wide=2)
        ctas = 2
    n_pad, lc_pad = shape
    ctx = L.Context(0)
    rng = np.random.default_rng(2)
    g = rng.choice(np.array([0, 1, 2, 3], np.uint8), size=(n_pad, lc_pad), p=[.4, .3, .25, .05])
    planes = planes_ref(g)
    host = np.stack([p if k_major else p.T for p in planes]).copy()
    d_planes = torch.from_numpy(host.reshape(-1)).cuda()
    d_counts = torch.zeros(4 * n_pad * n_pad, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()  # the library runs on its own non-blocking stream
    eb, ms = L.errbuf(), C.c_float(0)
    for rep in range(2):  # two launches accumulate
        L.check(L.lib().tpg_gram(ctx.handle, d_planes.data_ptr(), n_pad * lc_pad, n_pad, lc_pad,
                                 int(k_major) | (2 if ctas == 1 else 0), 0b1111, 0, 1, d_counts.data_ptr(), C.byref(ms), eb, len(eb)), eb)
    got = d_counts.cpu().numpy().reshape(4, n_pad, n_pad).transpose(0, 2, 1)  # [mat][i][j]
    d, nm, h1 = [p.astype(np.float64) for p in planes]
    iu = np.triu_indices(n_pad)
    for mat, ref in ((L.MAT_DD, d @ d.T), (L.MAT_NN, nm @ nm.T), (L.MAT_HH, h1 @ h1.T)):
        assert np.array_equal(got[mat][iu], 2 * ref[iu].astype(np.int64)), mat
    assert np.array_equal(got[L.MAT_HN], 2 * (h1 @ nm.T).astype(np.int64))


@pytest.mark.parametrize("shape", [(929, 700), (100, 37), (513, 130), (2049, 64)])
def test_decode_planes_pk2_kernel(shape):
    """decode of a 2-bit packed chunk (what host-sourced chunks look like after staging)."""
    import torch
    n, lc = shape
    ctx = L.Context(0)
    rng = np.random.default_rng(11)
    raw = rng.integers(0, 256, size=(lc, n)).astype(np.uint8)
    raw[rng.random(raw.shape) < 0.8] %= 4
    lib = L.lib()
    st = int(lib.tpg_pk2_stride(n))
    packed = np.zeros((lc, st), np.uint8)
    lib.tpg_pk2_pack_host(raw.ctypes.data, n, n, lc, packed.ctypes.data)
    n_pad, lc_pad = (n + 255) // 256 * 256, (lc + 127) // 128 * 128
    d_raw = torch.from_numpy(packed.reshape(-1)).cuda()
    ps = n_pad * lc_pad
    d_planes = torch.full((3 * ps,), 77, dtype=torch.int8, device="cuda")
    torch.cuda.synchronize()  # the library runs on its own non-blocking stream
    eb = L.errbuf()
    L.check(lib.tpg_decode_planes(ctx.handle, d_raw.data_ptr(), packed.size, st, None, 0, n, n_pad,
                                  None, 0, lc, lc_pad, d_planes.data_ptr(), ps, 4, eb, len(eb)), eb)
    got = d_planes.cpu().numpy().reshape(3, lc_pad, n_pad).transpose(0, 2, 1)
    ref = np.zeros((3, n_pad, lc_pad), np.int8)
    for p, x in enumerate(planes_ref(raw.T)):
        ref[p, :n, :lc] = x
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("shape", [(929, 700), (100, 37), (513, 130), (2049, 64), (128, 256)])
def test_decode_planes_pk2_k4_kernel(shape):
    """a 2-bit packed chunk decoded straight to the K-major 4-bit planes of the kind::mxf4 Gram kernel."""
    import torch
    n, lc = shape
    ctx = L.Context(0)
    rng = np.random.default_rng(12)
    raw = rng.integers(0, 256, size=(lc, n)).astype(np.uint8)
    raw[rng.random(raw.shape) < 0.8] %= 4
    lib = L.lib()
    st = int(lib.tpg_pk2_stride(n))
    packed = np.zeros((lc, st), np.uint8)
    lib.tpg_pk2_pack_host(raw.ctypes.data, n, n, lc, packed.ctypes.data)
    n_pad, lc_pad = (n + 255) // 256 * 256, (lc + 255) // 256 * 256
    d_raw = torch.from_numpy(packed.reshape(-1)).cuda()
    ps = n_pad * lc_pad // 2
    d_planes = torch.full((3 * ps,), 0x77, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    eb = L.errbuf()
    L.check(lib.tpg_decode_planes(ctx.handle, d_raw.data_ptr(), packed.size, st, None, 0, n, n_pad,
                                  None, 0, lc, lc_pad, d_planes.data_ptr(), ps, 4 | 8, eb, len(eb)), eb)
    got = d_planes.cpu().numpy().reshape(3, n_pad, lc_pad // 2)
    for p, x in enumerate(planes_ref(raw.T)):
        full = np.zeros((n_pad, lc_pad), np.int8)
        full[:n, :lc] = x
        assert np.array_equal(got[p], k4_pack(full)), p


def test_hybrid_link_and_pack_chunks(tpg_cfg):
    """Pinned host source, many small chunks: some chunks are DMA'd raw, the rest cross PCIe
    2-bit packed (the hybrid schedule); the result does not depend on which chunk went how."""
    import torch
    rng = np.random.default_rng(21)
    n, m = 700, 40000
    raw = random_bk(rng, n, m, 0.05, codes=(0, 1, 2, 3, 9))
    X = fbm_of(raw, n, m)
    ref, Kr, Nr = tpg.snp_king_counts(X, flags=L.FLAG_NO_PACK)
    pinned = torch.from_numpy(raw).pin_memory()
    src = (L.SRC_HOST, pinned.data_ptr(), pinned)
    win = np.arange(11, 611, dtype=np.int32)          # contiguous row window: still packable
    ref_w = tpg.snp_king(X, ind_row=win, flags=L.FLAG_NO_PACK)
    for link in ("45", "0.5"):                        # a slow modelled link makes later chunks pack
        tpg_cfg(link_gbps=float(link))
        king, K, N = tpg.snp_king_counts(X, src=src, gpu_chunk_loci=2048)
        st = dict(tpg.snp.last_stats)
        assert st["n_chunks"] == 20 and 0 <= st["n_packed_chunks"] <= 20
        if link == "0.5":
            assert 1 <= st["n_packed_chunks"] < 20    # chunk 0 always goes raw (idle link)
        assert np.array_equal(K, Kr) and np.array_equal(N, Nr) and same(king, ref)
        assert same(tpg.snp_king(X, ind_row=win, src=src, gpu_chunk_loci=2048), ref_w)
    # file source: every chunk is packed
    tpg.snp_king(X, gpu_chunk_loci=4096)
    assert tpg.snp.last_stats["n_packed_chunks"] == tpg.snp.last_stats["n_chunks"] == 10


def k4_pack(plane):
    """numpy statement of the K-major 4-bit (E2M1) plane layout of decode_planes_k4: plane [rows, loci]
    in {-1, 0, +1} -> [rows, loci / 2] bytes; byte k of a group of 8 loci holds locus k (low nibble)
    and locus k + 4 (high nibble); +1 = 0x2, -1 = 0xA."""
    code = np.where(plane > 0, 0x2, np.where(plane < 0, 0xA, 0)).astype(np.uint8)
    r, lc = code.shape
    g = code.reshape(r, lc // 8, 2, 4)              # [row, group, half, k]
    return (g[:, :, 0, :] | (g[:, :, 1, :] << 4)).reshape(r, lc // 2)


@pytest.mark.parametrize("build", list(K4_BUILDS))
@pytest.mark.parametrize("nrow", [929, 928, 1023], ids=["odd_stride", "aligned_stride", "stride_3mod4"])
@pytest.mark.parametrize("gather", [False, True])
def test_decode_planes_k4_kernel(gather, nrow, build, tpg_cfg):
    import torch
    tpg_cfg(**K4_BUILDS[build])  # the default kernel and the round's first one
    ctx = L.Context(0)
    rng = np.random.default_rng(3)
    ncol = 900
    raw = rng.integers(0, 256, size=(ncol, nrow)).astype(np.uint8)
    raw[rng.random(raw.shape) < 0.8] %= 4
    if gather:
        rows = rng.permutation(nrow)[:613].astype(np.int32)
        cols = rng.permutation(ncol)[:391].astype(np.int32)
    else:
        rows = np.arange(5, 5 + 801, dtype=np.int32)
        cols = np.arange(3, 3 + 700, dtype=np.int32)
    n, lc = len(rows), len(cols)
    n_pad, lc_pad = (n + 255) // 256 * 256, (lc + 255) // 256 * 256
    d_raw = torch.from_numpy(raw.reshape(-1)).cuda()
    d_rows, d_cols = torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda()
    ps = n_pad * lc_pad // 2
    d_planes = torch.full((3 * ps,), 0x77, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    eb = L.errbuf()
    L.check(L.lib().tpg_decode_planes(ctx.handle, d_raw.data_ptr(), raw.size, nrow,
                                      d_rows.data_ptr() if gather else None, int(rows[0]), n, n_pad,
                                      d_cols.data_ptr() if gather else None, int(cols[0]), lc, lc_pad,
                                      d_planes.data_ptr(), ps, 8, eb, len(eb)), eb)
    got = d_planes.cpu().numpy().reshape(3, n_pad, lc_pad // 2)
    for p, x in enumerate(planes_ref(raw[np.ix_(cols, rows)].T)):
        full = np.zeros((n_pad, lc_pad), np.int8)
        full[:n, :lc] = x
        assert np.array_equal(got[p], k4_pack(full)), p


@pytest.mark.parametrize("shape", [(256, 256), (512, 1280), (1024, 4096), (2560, 2048)])
@pytest.mark.parametrize("wide", [1, 2], ids=["pingpong_256x240", "wide_256x480"])
def test_gram_mxf4_kernel(shape, wide, tpg_cfg):
    """kind::mxf4 Gram kernels on K-major 4-bit planes (the ping-pong 256 x 240 kernel and the wide 256 x 480
    one): integer counts, exact."""
    import torch
    tpg_cfg(gram_wide=wide)
    n_pad, lc_pad = shape
    ctx = L.Context(0)
    rng = np.random.default_rng(2)
    g = rng.choice(np.array([0, 1, 2, 3], np.uint8), size=(n_pad, lc_pad), p=[.4, .3, .25, .05])
    planes = planes_ref(g)
    host = np.stack([k4_pack(p) for p in planes]).copy()
    d_planes = torch.from_numpy(host.reshape(-1)).cuda()
    d_counts = torch.zeros(4 * n_pad * n_pad, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    eb, ms = L.errbuf(), C.c_float(0)
    for rep in range(2):  # two launches accumulate
        L.check(L.lib().tpg_gram(ctx.handle, d_planes.data_ptr(), n_pad * lc_pad // 2, n_pad, lc_pad, 8, 0b1111, 0, 1,
                                 d_counts.data_ptr(), C.byref(ms), eb, len(eb)), eb)
    got = d_counts.cpu().numpy().reshape(4, n_pad, n_pad).transpose(0, 2, 1)  # [mat][i][j]
    d, nm, h1 = [p.astype(np.float64) for p in planes]
    iu = np.triu_indices(n_pad)
    for mat, ref in ((L.MAT_DD, d @ d.T), (L.MAT_NN, nm @ nm.T), (L.MAT_HH, h1 @ h1.T)):
        assert np.array_equal(got[mat][iu], 2 * ref[iu].astype(np.int64)), mat
    assert np.array_equal(got[L.MAT_HN], 2 * (h1 @ nm.T).astype(np.int64))


def test_synth_device_equals_host():
    import torch
    ctx = L.Context(0)
    nrow, ncol = 301, 999
    d = torch.zeros(nrow * ncol, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()  # the library runs on its own non-blocking stream
    eb = L.errbuf()
    for mode in (0, 1, 2):
        L.check(L.lib().tpg_synth_fill(ctx.handle, d.data_ptr(), nrow, ncol, 17, 2026, 0.1, mode, eb, len(eb)), eb)
        h = np.zeros(nrow * ncol, np.uint8)
        L.lib().tpg_synth_fill_host(h.ctypes.data, nrow, ncol, 17, 2026, 0.1, mode)
        assert np.array_equal(d.cpu().numpy(), h), mode


# ------------------------------------------------------------------ full-size properties (BASELINE configs[1])

def test_hgdp_shape_properties():
    """929 x 650,000, 2% missing — too big for the oracle; size-independent properties instead:
    linearity over loci (counts of A u B = counts(A) + counts(B)), chunk-size invariance,
    operand-layout invariance, symmetry, diagonal identities, and a subsample checked
    against the oracle."""
    import torch
    n, m = 929, 650_000
    ctx = L.Context(0)
    dev = torch.empty(n * m, dtype=torch.uint8, device="cuda")
    eb = L.errbuf()
    L.check(L.lib().tpg_synth_fill(ctx.handle, dev.data_ptr(), n, m, 0, 20260002, 0.02, 0, eb, len(eb)), eb)
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<device>", n, m, None
    src = (L.SRC_DEVICE, dev.data_ptr(), dev)
    king, K, N = tpg.snp_king_counts(X, src=src)
    cut = 271_001
    a = tpg.snp_king_counts(X, ind_col=np.arange(1, cut + 1, dtype=np.int32), src=src, gpu_chunk_loci=65536)
    b = tpg.snp_king_counts(X, ind_col=np.arange(cut + 1, m + 1, dtype=np.int32), src=src,
                            flags=L.FLAG_PLANES_K_MAJOR)
    assert np.array_equal(K, a[1] + b[1]) and np.array_equal(N, a[2] + b[2])          # linearity
    assert np.array_equal(king, king.T) and np.array_equal(K, K.T)                     # symmetry
    assert np.array_equal(np.diag(king), np.full(n, 0.5))                              # KING(i,i) = 0.5
    het = np.diag(N)
    assert np.array_equal(np.diag(K), het)                                             # K(i,i) = #het_i
    r = tpg.snp_ibs(X, type="raw_counts", src=src)
    ibs, valid = r["ibs"][:], r["valid_n"][:]
    assert np.array_equal(np.diag(ibs), np.diag(valid)) and np.all(ibs <= valid) and valid.max() <= 2 * m
    assert abs(np.diag(valid).mean() / (2 * m) - 0.98) < 0.002                         # 2% missing
    # a 1,200-locus subsample against the oracle
    cols = np.sort(np.random.default_rng(0).choice(m, 1200, replace=False)).astype(np.int32) + 1
    host = np.zeros(n * m, np.uint8)
    L.lib().tpg_synth_fill_host(host.ctypes.data, n, m, 0, 20260002, 0.02, 0)
    assert np.array_equal(dev.cpu().numpy(), host)
    sub = tpg.snp_king_counts(X, ind_col=cols, src=src)
    ref = O.snp_king(host, n, m, None, cols)
    assert np.array_equal(sub[1], ref["K"]) and np.array_equal(sub[2], ref["N_Aa_i"]) and same(sub[0], ref["king"])


def test_hgdp_shape_full_size_against_the_blas_port():
    """C2 at FULL size (929 x 650,000, 2 % missing), every locus: the KING numerator, N_Aa_i and the finished
    matrix equal, bit for bit, the reference's block loop restated over a threaded FP64 BLAS
    (oracle/blas_baseline.py — pinned to oracle.c by tests/test_cpu_baseline.py) — on the default 4-bit route
    and on the int8 route.  ~20 s of CPU dgemm."""
    from oracle import blas_baseline as B
    n, m = 929, 650_000
    host = np.zeros(n * m, np.uint8)
    L.lib().tpg_synth_fill_host(host.ctypes.data, n, m, 0, 20260002, 0.02, 0)
    Kr, Nr, done, _ = B.king_blocks(host.reshape(m, n), n, m, O.block_size(n))
    assert done == m
    ref = B.king_finish(Kr, Nr)
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<host>", n, m, None
    for flags in (0, L.FLAG_GRAM_I8):
        king, K, N = tpg.snp_king_counts(X, src=(L.SRC_HOST, host.ctypes.data, host), flags=flags)
        assert np.array_equal(K, Kr) and np.array_equal(N, Nr), flags
        assert same(king, ref), flags


# ------------------------------------------------------------------ sharding (staged C-ABI), ranks emulated on one GPU

def _staged(engine, X, stat, bs, mode, rank, world, counts=None, finish=True, **kw):
    from tidypopgen_b200 import snp
    shard = {"loci": L.SHARD_LOCI, "tiles": L.SHARD_TILES, "none": L.SHARD_NONE}[mode]
    job, keep = snp.make_job(X, kw.get("rows"), kw.get("cols"), stat, bs, shard_mode=shard, shard_rank=rank,
                             shard_world=world, gpu_chunk_loci=kw.get("chunk", 0))
    n = job.n if job.ind_row else X.nrow
    counts = engine.new_counts(n) if counts is None else counts
    engine.accumulate(job, counts)
    if not finish:
        return counts, job, keep
    return engine.finish(job, counts, n, stat == L.STAT_IBS_RAW)[:2]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_loci_and_tile_sharding_emulated(world):
    from tidypopgen_b200 import dist as tdist
    rng = np.random.default_rng(world)
    n, m = 700, 6000
    raw = random_bk(rng, n, m, 0.06)
    X = fbm_of(raw, n, m)
    eng = tdist.CudaEngine(0)
    cols = pick("every3", m, rng)
    ref_king = tpg.snp_king(X, ind_col=cols, block_size=500)
    ref_as = tpg.snp_allele_sharing(X, ind_col=cols, block_size=500)
    # loci: every rank adds its slice into the same buffer (what the all-reduce sums), finish once
    for stat, ref in ((L.STAT_KING, ref_king), (L.STAT_AS, ref_as)):
        counts = None
        for r in range(world):
            counts, job, keep = _staged(eng, X, stat, 500, "loci", r, world, counts, finish=False, cols=cols, chunk=512)
        job.shard_mode, job.shard_world, job.shard_rank = L.SHARD_NONE, 1, 0
        out, _, _ = eng.finish(job, counts, n, False)
        assert same(out, ref)
    # tiles: each rank's blocks, merged by ownership
    merged = np.full((n, n), -7.0)
    for r in range(world):
        out, _ = _staged(eng, X, L.STAT_KING, 500, "tiles", r, world, cols=cols)
        mask = tdist.owned_mask(n, r, world)
        assert np.all(out[~mask] == 0)
        merged[mask] = out[mask]
    assert same(merged, ref_king)


def _king_pair(gi, gj):
    """KING-robust of one pair from raw bytes, the reference's arithmetic (src/snp_king.cpp:45-72,
    R/snp_king.R:82-91)."""
    h0i, h1i, h2i, vi = gi == 0, gi == 1, gi == 2, gi <= 2
    h0j, h1j, h2j, vj = gj == 0, gj == 1, gj == 2, gj <= 2
    K = float((h1i & h1j).sum()) - 2.0 * float((h0i & h2j).sum() + (h2i & h0j).sum())
    Ni, Nj = float((h1i & vj).sum()), float((vi & h1j).sum())
    pm = min(Ni, Nj)
    with np.errstate(all="ignore"):
        return np.float64(K) / np.float64(2 * pm) + 0.5 - np.float64(0.25 * (Ni + Nj)) / np.float64(pm)


def test_biobank_regime_sampled_pairs_and_region_sharding():
    """The many-tiles regime of BASELINE configs[3] at a size the suite can afford (20,000 x 40,000:
    79 x 79 blocks, 20 x 20 ownership regions): whole matrix on one GPU, and ranks 0 / 3 / 7 of an
    8-way tile-sharded run emulated on the same GPU.  The oracle is out of reach, so sampled entries
    are checked EXACTLY against the reference formula on the two genotype columns; owned regions must
    equal the unsharded result bit for bit and everything else must stay untouched."""
    import torch
    from tidypopgen_b200 import dist as tdist
    n, m, world = 20000, 40000, 8
    eng = tdist.CudaEngine(0)
    eb = L.errbuf()
    raw = torch.empty(n * m + 64, dtype=torch.uint8, device="cuda")
    L.check(L.lib().tpg_synth_fill(eng.ctx.handle, raw.data_ptr(), n, m, 0, 20260004, 0.01, 0, eb, len(eb)), eb)
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<device>", n, m, None
    src = (L.SRC_DEVICE, raw.data_ptr(), raw)
    from tidypopgen_b200 import snp
    def run(shard, rank, flags=0):
        job, keep = snp.make_job(X, None, None, L.STAT_KING, snp.block_size(n), src=src, flags=flags,
                                 shard_mode=L.SHARD_TILES if shard else L.SHARD_NONE, shard_rank=rank,
                                 shard_world=world if shard else 1)
        counts = eng.new_counts(n)
        eng.accumulate(job, counts)
        out, _, _ = eng.finish(job, counts, n, False, on_device=True)
        del counts
        return out                                   # torch [i, j] view of the column-major result
    full = run(False, 0)
    assert eng.last_stats["n_mxf4_chunks"] >= 1          # the default (4-bit) route ...
    i8 = run(False, 0, L.FLAG_GRAM_I8)                   # ... and the int8 route give the same matrix, bit for bit
    assert eng.last_stats["n_mxf4_chunks"] == 0
    assert torch.equal(i8.nan_to_num(-9.0), full.nan_to_num(-9.0))
    del i8
    cols = raw[: n * m].view(m, n)
    rng = np.random.default_rng(4)
    for k in range(48):
        i, j = (int(x) for x in rng.integers(0, n, 2))
        if k % 6 == 0:
            j = i
        ref = _king_pair(cols[:, i].cpu().numpy(), cols[:, j].cpu().numpy())
        got = full[i, j].item()
        assert got == ref or (np.isnan(got) and np.isnan(ref)), (i, j)
        assert full[j, i].item() == got or np.isnan(got)
    nb = (n + 255) // 256
    assert tdist.shard_region_blocks(nb, world) == 4
    blk = torch.arange(n, device="cuda") // 256 // 4
    bi, bj = torch.minimum(blk[:, None], blk[None, :]), torch.maximum(blk[:, None], blk[None, :])
    nr = (nb + 3) // 4
    owner = (bi * nr - bi * (bi - 1) // 2 + (bj - bi)) % world
    del bi, bj
    for rank in (0, 3, 7):
        part = run(True, rank)
        mine = owner == rank
        assert torch.equal(part[mine].nan_to_num(-9.0), full[mine].nan_to_num(-9.0))
        assert bool((part[~mine] == 0).all())
        del part, mine


def test_two_gpus_nccl_if_available():
    """Real NCCL run (torchrun, 2 ranks) when the box has >= 2 GPUs; skipped on the 1-GPU box."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = __import__("pathlib").Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(root / "scripts" / "dist_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


# ------------------------------------------------------------------ BASELINE configs[2] and [4] at full size

def _device_fbm(n, m):
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<device>", n, m, None
    return X


def test_1000g_shape_grm_imputed_bytes():
    """2504 x 1,000,000, imputed codes 4/5/6 (configs[2]): the raw-byte rule makes imputed files
    identical to NA files; default block.size gives the pad-quirk constant 8; GRM is centred;
    a subsample is checked against the oracle."""
    import torch
    n, m = 2504, 1_000_000
    ctx = L.Context(0)
    eb = L.errbuf()
    X = _device_fbm(n, m)
    res = {}
    for mode in (1, 0):
        dev = torch.empty(n * m, dtype=torch.uint8, device="cuda")
        L.check(L.lib().tpg_synth_fill(ctx.handle, dev.data_ptr(), n, m, 0, 20260003, 0.01, mode, eb, len(eb)), eb)
        src = (L.SRC_DEVICE, dev.data_ptr(), dev)
        res[mode] = tpg.snp_allele_sharing_counts(X, src=src)           # default block.size = block_size(2504)
        if mode == 1:
            assert tpg.snp.last_stats["n_short"] == 8 == O.n_short(m, O.block_size(n))
            assert set(torch.unique(dev[:10_000_000]).tolist()) <= {0, 1, 2, 4, 5, 6}
            grm = tpg.snp_grm(X, src=src)
            clean = tpg.snp_allele_sharing_counts(X, src=src, pad_quirk=False)
            cols = np.sort(np.random.default_rng(1).choice(m, 700, replace=False)).astype(np.int32) + 1
            sub = tpg.snp_allele_sharing_counts(X, ind_col=cols, src=src, block_size=101)
            host_cols = dev.view(m, n)[torch.from_numpy(cols.astype(np.int64) - 1).cuda()].cpu().numpy()
        del dev
    a1, num1, den1 = res[1]
    a0, num0, den0 = res[0]
    assert np.array_equal(num1, num0) and np.array_equal(den1, den0) and same(a1, a0)     # Q2: imputed == missing
    assert np.array_equal(num1, clean[1] + 8) and np.array_equal(den1, clean[2])          # Q1: +n_short
    assert np.array_equal(num1, num1.T) and np.array_equal(np.diag(clean[1]) <= np.diag(den1), np.ones(n, bool))
    off = ~np.eye(n, dtype=bool)
    assert abs(grm[off].mean()) < 1e-12                                                    # centred by mb
    mb = a1[off].mean()
    assert np.allclose(grm, 2 * (a1 - mb) / (1 - mb), rtol=0, atol=GRM_ATOL)
    ref = O.snp_allele_sharing(host_cols.reshape(-1), n, 700, block_size_=101)
    assert np.array_equal(sub[1], ref["num"]) and np.array_equal(sub[2], ref["denom"]) and same(sub[0], ref["as"])


def test_mixed_ploidy_shape_ibs_subsets():
    """10,000 x 1,000,000, 10% missing, pseudohaploid + tetraploid-coded rows (configs[4]):
    snp_ibs with a shuffled 70% row subset and every-3rd + shuffled-tail loci, all three types."""
    import torch
    n_tot, m_tot = 10_000, 1_000_000
    ctx = L.Context(0)
    eb = L.errbuf()
    dev = torch.empty(n_tot * m_tot, dtype=torch.uint8, device="cuda")
    L.check(L.lib().tpg_synth_fill(ctx.handle, dev.data_ptr(), n_tot, m_tot, 0, 20260005, 0.10, 2, eb, len(eb)), eb)
    assert set(torch.unique(dev[:20_000_000]).tolist()) == {0, 1, 2, 3, 4, 5}
    X = _device_fbm(n_tot, m_tot)
    src = (L.SRC_DEVICE, dev.data_ptr(), dev)
    rng = np.random.default_rng(5)
    rows = (rng.permutation(n_tot)[: n_tot * 7 // 10] + 1).astype(np.int32)
    tail = rng.permutation(m_tot)[: m_tot // 20]
    cols = (np.concatenate([np.arange(0, m_tot, 3), tail]) + 1).astype(np.int32)
    n, m = len(rows), len(cols)
    raw_counts = tpg.snp_ibs(X, rows, cols, type="raw_counts", src=src)
    ibs, valid = raw_counts["ibs"][:], raw_counts["valid_n"][:]
    prop = tpg.snp_ibs(X, rows, cols, type="proportion", src=src)
    adj = tpg.snp_ibs(X, rows, cols, type="adjusted_counts", src=src)
    with np.errstate(all="ignore"):
        assert same(prop, ibs / valid) and same(adj, (ibs / valid) * m)
    assert np.array_equal(ibs, ibs.T) and np.array_equal(np.diag(ibs), np.diag(valid)) and valid.max() <= 2 * m
    assert 0.78 < np.diag(valid).mean() / (2 * m) < 0.92   # 10% missing, tetraploid rows lose dosages 3 and 4 too
    # exact check of a 48-row x 900-locus sub-job against the oracle (bytes pulled from the device)
    r_sub, c_sub = rows[:48], cols[rng.permutation(m)[:900]]
    host = dev.view(m_tot, n_tot)[torch.from_numpy(c_sub.astype(np.int64) - 1).cuda()].cpu().numpy()  # [900, n_tot]
    got = tpg.snp_ibs(X, r_sub, c_sub, type="raw_counts", src=src)
    ref = O.snp_ibs(host.reshape(-1), n_tot, 900, r_sub, None, type="raw_counts")
    assert np.array_equal(got["ibs"][:], ref["ibs"]) and np.array_equal(got["valid_n"][:], ref["valid_n"])
    # and the same 48 individuals inside the big job
    assert np.array_equal(ibs[:48, :48].diagonal(), np.diag(valid)[:48])


@pytest.mark.parametrize("shard", ["loci", "tiles"])
def test_in_process_multi_gpu_driver(shard, families):
    """tpg_pairwise_multi (one process, one host thread per device, counts summed on the first
    device).  With one GPU the ranks share it (devices = [0, 0, 0]); with more they spread out."""
    import torch
    ndev = torch.cuda.device_count()
    devices = [r % ndev for r in range(3)]
    rng = np.random.default_rng(21)
    n, m = 600, 9000
    raw = random_bk(rng, n, m, 0.07, codes=(0, 1, 2, 3, 5))
    X = fbm_of(raw, n, m)
    rows, cols = pick("subset", n, rng), pick("every3", m, rng)
    for fn, kw in ((tpg.snp_king, {}), (tpg.snp_allele_sharing, {"block_size": 333}), (tpg.snp_grm, {"block_size": 333})):
        ref = fn(X, rows, cols, **kw)
        got = fn(X, rows, cols, devices=devices, shard=shard, gpu_chunk_loci=1024, **kw)
        if fn is tpg.snp_grm:
            assert np.allclose(got, ref, rtol=0, atol=GRM_ATOL, equal_nan=True)
        else:
            assert same(got, ref)
    r1 = tpg.snp_ibs(X, type="raw_counts")
    r2 = tpg.snp_ibs(X, type="raw_counts", devices=devices, shard=shard)
    assert np.array_equal(r1["ibs"][:], r2["ibs"][:]) and np.array_equal(r1["valid_n"][:], r2["valid_n"][:])
    f = families
    Xf = tpg.FBMCode256(f["path"], 12, 961)   # fewer loci blocks than ranks in places: still exact
    assert same(tpg.snp_king(Xf, devices=[0] * 8, shard=shard), O.snp_king(f["raw"], 12, 961)["king"])
