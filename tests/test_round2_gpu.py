"""Round-2 parity tests (all through the C ABI, against the CPU oracle):
the exactness guard of the 4-bit route, one pass for several statistics and the count-matrix
cache, file-backed results (snp_ibs raw_counts), the shared-memory row gather for device and host
sources (packed over PCIe), O_DIRECT reads, and the int32 / FP64 finish arithmetic."""
import os

import numpy as np
import pytest

import tidypopgen_b200 as tpg
from conftest import same
from oracle import oracle as O
from tidypopgen_b200 import _lib as L
from test_parity_gpu import GRM_ATOL, fbm_of, pick, random_bk

pytestmark = pytest.mark.gpu


def device_fbm(raw, n, m):
    import torch
    dev = torch.from_numpy(np.ascontiguousarray(raw)).cuda()
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<device>", n, m, None
    return X, (L.SRC_DEVICE, dev.data_ptr(), dev)


# ------------------------------------------------------------------ 4-bit route: exact below 2^24, guarded above

def test_mxf4_chain_guard_matches_int8(tpg_cfg):
    """One accumulation chain of the kind::mxf4 kernel must never exceed 2^23 loci: a single chunk of
    2^24 + 2^21 loci with the k-split forced to 1 is split anyway, and the counts equal the int8 -> s32
    route bit for bit (they would not: NN ~ 0.98 m > 2^24 with odd partial sums in an FP32 chain).
    Exact-integer requirement of src/snp_king.cpp:70-72."""
    import torch
    n, m = 256, (1 << 24) + (1 << 21)
    ctx = L.Context(0)
    eb = L.errbuf()
    dev = torch.empty(n * m + 64, dtype=torch.uint8, device="cuda:0")
    L.check(L.lib().tpg_synth_fill(ctx.handle, dev.data_ptr(), n, m, 0, 77, 0.02, 0, eb, len(eb)), eb)
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<device>", n, m, None
    src = (L.SRC_DEVICE, dev.data_ptr(), dev)
    tpg_cfg(gram_ksplit=1)                              # ask for ONE k-split per tile: the guard must override it
    king4, K4, N4 = tpg.snp_king_counts(X, src=src, gpu_chunk_loci=m)
    st = dict(tpg.snp.last_stats)
    assert st["n_chunks"] == 1 and st["n_mxf4_chunks"] == 1 and st["n_i8_chunks"] == 0
    assert 0 < st["max_chain_loci"] <= L.MXF4_MAX_CHAIN_LOCI
    ibs4 = tpg.snp_ibs(X, type="raw_counts", src=src, gpu_chunk_loci=m)
    tpg_cfg(reset=True)
    king8, K8, N8 = tpg.snp_king_counts(X, src=src, gpu_chunk_loci=m, flags=L.FLAG_GRAM_I8)
    st8 = dict(tpg.snp.last_stats)
    assert st8["n_mxf4_chunks"] == 0 and st8["n_i8_chunks"] == 1 and st8["max_chain_loci"] == 0
    ibs8 = tpg.snp_ibs(X, type="raw_counts", src=src, gpu_chunk_loci=m, flags=L.FLAG_GRAM_I8)
    assert np.array_equal(K4, K8) and np.array_equal(N4, N8) and same(king4, king8)
    assert np.array_equal(ibs4["ibs"][:], ibs8["ibs"][:]) and np.array_equal(ibs4["valid_n"][:], ibs8["valid_n"][:])
    v = ibs8["valid_n"][:]
    assert v.max() > 2 * (1 << 24) and np.all(np.mod(v, 2) == 0)  # the chain really crossed 2^24
    # a sampled pair against the reference's per-pair arithmetic (src/snp_king.cpp:45-72)
    cols = dev[: n * m].view(m, n)
    gi, gj = cols[:, 3].cpu().numpy(), cols[:, 200].cpu().numpy()
    n_aa = int(((gi == 1) & (gj <= 2)).sum())
    assert N8[3, 200] == n_aa and N4[3, 200] == n_aa


# ------------------------------------------------------------------ one pass, several statistics

@pytest.mark.parametrize("rk,ck", [("all", "all"), ("shuffle", "every3")])
def test_pairwise_many_equals_single_calls(rk, ck):
    """tpg_pairwise_many: KING + IBS (three types) + allele sharing + GRM from ONE decode and one set of
    products equal the statistics computed one job at a time (R/snp_king.R:32, R/snp_ibs.R:42,
    R/snp_allele_sharing.R:33) and the oracle."""
    rng = np.random.default_rng(31)
    n_tot, m_tot, bs = 421, 5000, 777
    raw = random_bk(rng, n_tot, m_tot, 0.08, codes=(0, 1, 2, 3, 4, 5, 6, 255))
    X = fbm_of(raw, n_tot, m_tot)
    rows, cols = pick(rk, n_tot, rng), pick(ck, m_tot, rng)
    out = tpg.snp_pairwise_many(X, ("king", "ibs", "ibs_raw", "ibs_adjusted", "allele_sharing", "grm"), rows, cols,
                                block_size=bs, gpu_chunk_loci=1024)
    st = dict(tpg.snp.last_stats)
    assert st["n_chunks"] == -(-(m_tot if cols is None else len(cols)) // 1024)   # ONE pass over the loci
    ref_k = O.snp_king(raw, n_tot, m_tot, rows, cols)
    assert same(out["king"], ref_k["king"])
    assert same(out["king"], tpg.snp_king(X, rows, cols))
    for key, typ in (("ibs", "proportion"), ("ibs_adjusted", "adjusted_counts")):
        assert same(out[key], O.snp_ibs(raw, n_tot, m_tot, rows, cols, type=typ)["out"]), typ
    ref_r = O.snp_ibs(raw, n_tot, m_tot, rows, cols, type="raw_counts")
    assert np.array_equal(out["ibs_raw"]["ibs"], ref_r["ibs"]) and np.array_equal(out["ibs_raw"]["valid_n"], ref_r["valid_n"])
    ref_a = O.snp_allele_sharing(raw, n_tot, m_tot, rows, cols, block_size_=bs)
    assert same(out["allele_sharing"], ref_a["as"]) and np.array_equal(O.is_na(out["allele_sharing"]), O.is_na(ref_a["as"]))
    ref_g, _ = O.grm_from_as(ref_a["as"])
    ok = ~np.isnan(ref_g)
    assert np.array_equal(np.isnan(out["grm"]), ~ok) and np.allclose(out["grm"][ok], ref_g[ok], rtol=0, atol=GRM_ATOL)
    assert same(out["grm"], tpg.snp_grm(X, rows, cols, block_size=bs))


def test_counts_cache_serves_the_next_statistic(tpg_cfg, tmp_path):
    """tpg_config.cache_counts: pairwise_king, then pairwise_ibs, then pairwise_allele_sharing on the same
    .bk / ind.row / ind.col = one pass over the genotypes; a rewritten file or other indices miss."""
    rng = np.random.default_rng(32)
    n, m = 300, 4000
    raw = random_bk(rng, n, m, 0.06)
    X = tpg.FBMCode256.from_bytes(raw, n, m, backingfile=tmp_path / "cache.bk")
    rows = pick("subset", n, rng)
    plain_k, plain_i, plain_a = tpg.snp_king(X, rows), tpg.snp_ibs(X, rows), tpg.snp_allele_sharing(X, rows, block_size=333)
    tpg_cfg(cache_counts=1)
    k = tpg.snp_king(X, rows)
    assert tpg.snp.last_stats["n_cached_jobs"] == 0 and tpg.snp.last_stats["n_chunks"] >= 1
    i = tpg.snp_ibs(X, rows)
    assert tpg.snp.last_stats["n_cached_jobs"] == 1 and tpg.snp.last_stats["n_chunks"] == 0
    a = tpg.snp_allele_sharing(X, rows, block_size=333)
    assert tpg.snp.last_stats["n_cached_jobs"] == 1
    assert same(k, plain_k) and same(i, plain_i) and same(a, plain_a)
    k_all = tpg.snp_king(X)                                               # other indices: a miss, still right
    assert tpg.snp.last_stats["n_cached_jobs"] == 0
    assert same(k_all, O.snp_king(raw, n, m)["king"])
    # the file changes under the same name: size / mtime differ -> recomputed
    raw2 = random_bk(np.random.default_rng(33), n, m, 0.06)
    tpg.snp_king(X, rows)
    os.utime(X.backingfile, ns=(1, 1))
    np.ascontiguousarray(raw2).tofile(X.backingfile)
    k2 = tpg.snp_king(X, rows)
    assert tpg.snp.last_stats["n_cached_jobs"] == 0
    assert same(k2, O.snp_king(raw2, n, m, rows)["king"])


@pytest.mark.parametrize("n,m", [(300, 4000), (1031, 2500)])
def test_genotype_cache_serves_other_subsets_and_statistics(n, m, tpg_cfg, tmp_path):
    """tpg_config.cache_genotypes: a whole-file job leaves the 2-bit packed .bk in HBM; later jobs on that file —
    any ind.row / ind.col (all, windows, shuffles, duplicates, gathered loci), any statistic — are gathered from
    it on the device (no chunk crosses PCIe) and equal the oracle; subset jobs do not fill the cache; a rewritten
    file misses and refills."""
    rng = np.random.default_rng(n + m)
    raw = random_bk(rng, n, m, 0.07)
    X = tpg.FBMCode256.from_bytes(raw, n, m, backingfile=tmp_path / "geno.bk")
    tpg_cfg(cache_genotypes=1)
    rows_sub = pick("shuffle", n, rng)
    i0 = tpg.snp_ibs(X, rows_sub)                                 # a subset job first: streams, does not fill
    assert tpg.snp.last_stats["n_geno_cached"] == 0 and same(i0, O.snp_ibs(raw, n, m, rows_sub)["out"])
    k_all = tpg.snp_king(X, gpu_chunk_loci=1024)                  # the whole file, in several chunks: streams AND fills
    st = dict(tpg.snp.last_stats)
    assert st["n_geno_cached"] == 0 and st["bytes_h2d"] > 0 and st["n_chunks"] > 1
    assert same(k_all, O.snp_king(raw, n, m)["king"])
    for rk, ck in (("all", "all"), ("shuffle", "all"), ("window", "every3"), ("dup", "subset"), ("subset", "shuffle")):
        rows, cols = pick(rk, n, rng), pick(ck, m, rng)
        k = tpg.snp_king(X, rows, cols)
        st = dict(tpg.snp.last_stats)
        assert st["n_geno_cached"] == 1 and st["bytes_h2d"] == 0 and st["bytes_file"] == 0, (rk, ck, st)
        assert same(k, O.snp_king(raw, n, m, rows, cols)["king"]), (rk, ck)
        a = tpg.snp_allele_sharing(X, rows, cols, block_size=257)
        assert tpg.snp.last_stats["n_geno_cached"] == 1
        assert same(a, O.snp_allele_sharing(raw, n, m, rows, cols, block_size_=257)["as"]), (rk, ck)
    rows = pick("subset", n, rng)
    i8 = tpg.snp_ibs(X, rows, flags=L.FLAG_GRAM_I8)              # int8 route from the cache
    assert tpg.snp.last_stats["n_geno_cached"] == 1 and same(i8, O.snp_ibs(raw, n, m, rows)["out"])
    # the file changes under the same name: the stale image is not used, the next whole-file job refills
    raw2 = random_bk(np.random.default_rng(5), n, m, 0.07)
    os.utime(X.backingfile, ns=(1, 1))
    np.ascontiguousarray(raw2).tofile(X.backingfile)
    rows = pick("window", n, rng)
    k2 = tpg.snp_king(X, rows)
    assert tpg.snp.last_stats["n_geno_cached"] == 0 and same(k2, O.snp_king(raw2, n, m, rows)["king"])
    tpg.snp_king(X)
    k3 = tpg.snp_king(X, rows)
    assert tpg.snp.last_stats["n_geno_cached"] == 1 and same(k3, k2)


def test_result_shells_equal_the_one_piece_path(tpg_cfg, tmp_path):
    """tpg_config.result_shells: a whole-file job with >= 27,136 individuals computes and hands back its matrix in
    nested shells (products of shell k+1 beside the D2H of shell k, shells 1.. from the packed image shell 0's
    streaming pass left in HBM).  Every statistic equals the one-piece path bit for bit — all outputs, every shell
    count — and the oracle on individuals sampled from every shell."""
    n, m = 27_136 + 77, 2_000
    rng = np.random.default_rng(77)
    raw = np.zeros(n * m, np.uint8)
    L.lib().tpg_synth_fill_host(raw.ctypes.data, n, m, 0, 977, 0.03, 0)
    X = tpg.FBMCode256.from_bytes(raw, n, m, backingfile=tmp_path / "shells.bk")
    tpg_cfg(result_shells=1)
    king1, K1, N1 = tpg.snp_king_counts(X)
    assert tpg.snp.last_stats["n_result_shells"] == 0
    for shells in (0, 3):
        tpg_cfg(result_shells=shells)
        king, K, N = tpg.snp_king_counts(X, gpu_chunk_loci=512)
        st = dict(tpg.snp.last_stats)
        assert st["n_result_shells"] == (shells or 4) and st["n_chunks"] >= 4 + (shells or 4) - 1, st   # (8 regions: auto = 4)
        assert same(king, king1) and np.array_equal(K, K1) and np.array_equal(N, N1), shells
        del king, K, N
    pick_rows = np.sort(rng.choice(n, 240, replace=False)).astype(np.int32) + 1     # spans every shell
    ref = O.snp_king(raw, n, m, pick_rows)
    ix = np.ix_(pick_rows - 1, pick_rows - 1)
    assert same(king1[ix], ref["king"]) and np.array_equal(K1[ix], ref["K"]) and np.array_equal(N1[ix], ref["N_Aa_i"])
    del king1, K1, N1
    tpg_cfg(result_shells=1)
    a1 = tpg.snp_allele_sharing(X, block_size=333)
    r1 = tpg.snp_ibs(X, type="raw_counts")
    i1, v1 = r1["ibs"][:].copy(), r1["valid_n"][:].copy()
    tpg_cfg(result_shells=0)
    a = tpg.snp_allele_sharing(X, block_size=333)                                 # pad quirk constant in every shell
    assert tpg.snp.last_stats["n_result_shells"] == 4 and same(a, a1)
    del a, a1
    r = tpg.snp_ibs(X, type="raw_counts")                                          # two file-backed outputs
    assert tpg.snp.last_stats["n_result_shells"] == 4
    assert np.array_equal(r["ibs"][:], i1) and np.array_equal(r["valid_n"][:], v1)
    g = tpg.snp_ibs(X, pick_rows)                                                  # a subset job: the one-piece path
    assert tpg.snp.last_stats["n_result_shells"] == 0 and same(g, O.snp_ibs(raw, n, m, pick_rows)["out"])


# ------------------------------------------------------------------ file-backed results (raw_counts)

def test_raw_counts_written_into_fbm_backing_files(tmp_path):
    """snp_ibs(type = "raw_counts") returns two file-backed double FBMs (R/snp_ibs.R:54-56,84-85): the
    library maps the files named in tpg_result.out0_path / out1_path and writes the counts there."""
    rng = np.random.default_rng(41)
    n, m = 129, 900
    raw = random_bk(rng, n, m, 0.1)
    X = fbm_of(raw, n, m)
    rows = pick("shuffle", n, rng)
    ref = O.snp_ibs(raw, n, m, rows, None, type="raw_counts")
    got = tpg.snp_ibs(X, rows, type="raw_counts")
    for name in ("ibs", "valid_n"):
        on_disk = np.fromfile(got[name].backingfile, dtype=np.float64).reshape(n, n).T   # column-major doubles
        assert np.array_equal(on_disk, ref[name]) and np.array_equal(got[name][:], ref[name])
    # plain ctypes, as the R glue calls it: out0 / out1 NULL, paths set
    p0, p1 = tmp_path / "ibs.bk", tmp_path / "valid.bk"
    for p in (p0, p1):
        np.zeros(n * n, np.float64).tofile(p)
    tpg.snp.run_job(X, rows, None, L.STAT_IBS_RAW, 100, out0_path=str(p0), out1_path=str(p1))
    assert np.array_equal(np.fromfile(p0, np.float64).reshape(n, n).T, ref["ibs"])
    assert np.array_equal(np.fromfile(p1, np.float64).reshape(n, n).T, ref["valid_n"])
    np.zeros(10, np.float64).tofile(p1)                                                 # too small: refused, no write
    with pytest.raises(L.TpgError) as ei:
        tpg.snp.run_job(X, rows, None, L.STAT_IBS_RAW, 100, out0_path=str(p0), out1_path=str(p1))
    assert ei.value.code == L.ERR_ARG and "smaller" in ei.value.msg
    with pytest.raises(L.TpgError) as ei:
        tpg.snp.run_job(X, rows, None, L.STAT_IBS_RAW, 100, out0_path=str(tmp_path / "nope.bk"), out1_path=str(p0))
    assert ei.value.code == L.ERR_IO


# ------------------------------------------------------------------ gathered rows through shared memory

@pytest.mark.parametrize("layout", [0, 1], ids=["column_major", "transposed_t4"])
@pytest.mark.parametrize("source", ["file", "host", "pinned"])
@pytest.mark.parametrize("rk", ["shuffle", "dup", "subset"])
def test_host_sources_gathered_rows_cross_packed(source, rk, layout, tpg_cfg):
    """ind.row in arbitrary order on file / host / pinned sources: the row span is 2-bit packed while staging,
    crosses PCIe at 0.25 B per genotype and is gathered on the device (gather_pack_kernel<packed>) — every
    chunk packed, every statistic exact; the int8 route and the raw route agree."""
    import torch
    tpg_cfg(gather_layout=layout)
    rng = np.random.default_rng(51)
    n_tot, m_tot = 1237, 3300
    raw = random_bk(rng, n_tot, m_tot, 0.08, codes=(0, 1, 2, 3, 4, 5, 6, 255))
    X = fbm_of(raw, n_tot, m_tot)
    src = None
    if source == "host":
        holder = np.ascontiguousarray(raw)
        src = (L.SRC_HOST, holder.ctypes.data, holder)
    elif source == "pinned":
        holder = torch.from_numpy(np.ascontiguousarray(raw)).pin_memory()
        src = (L.SRC_HOST, holder.data_ptr(), holder)
    rows, cols = pick(rk, n_tot, rng), pick("every3", m_tot, rng)
    ref = O.snp_king(raw, n_tot, m_tot, rows, cols)
    for flags in (0, L.FLAG_GRAM_I8):
        king, K, N = tpg.snp_king_counts(X, rows, cols, src=src, gpu_chunk_loci=512, flags=flags)
        st = dict(tpg.snp.last_stats)
        assert st["n_packed_chunks"] == st["n_chunks"] == 3, st
        assert np.array_equal(K, ref["K"]) and np.array_equal(N, ref["N_Aa_i"]) and same(king, ref["king"])
    king, K, N = tpg.snp_king_counts(X, rows, cols, src=src, gpu_chunk_loci=512, flags=L.FLAG_NO_PACK)
    assert tpg.snp.last_stats["n_packed_chunks"] == 0
    assert np.array_equal(K, ref["K"]) and same(king, ref["king"])
    ref_i = O.snp_ibs(raw, n_tot, m_tot, rows, cols, type="raw_counts")
    got = tpg.snp_ibs(X, rows, cols, type="raw_counts", src=src)
    assert np.array_equal(got["ibs"][:], ref_i["ibs"]) and np.array_equal(got["valid_n"][:], ref_i["valid_n"])


@pytest.mark.parametrize("where", ["device", "file"])
def test_gathered_rows_of_long_columns(where):
    """Columns too long for a CTA's shared memory take the older paths: 230,000-row columns as raw bytes on
    the device (compaction through L2), but still packed (57.5 KB of codes per column) from a file."""
    rng = np.random.default_rng(52)
    n_tot, m_tot = 230_000, 130
    raw = rng.integers(0, 4, size=n_tot * m_tot, dtype=np.uint8)
    rows = rng.choice(np.arange(1, n_tot + 1), size=300, replace=False).astype(np.int32)
    rows[0], rows[1] = 1, n_tot                           # the span is the whole column
    ref = O.snp_king(raw, n_tot, m_tot, rows, None)
    if where == "device":
        X, src = device_fbm(raw, n_tot, m_tot)
        king, K, N = tpg.snp_king_counts(X, rows, None, src=src)
    else:
        X = fbm_of(raw, n_tot, m_tot)
        king, K, N = tpg.snp_king_counts(X, rows, None)
        assert tpg.snp.last_stats["n_packed_chunks"] == tpg.snp.last_stats["n_chunks"]
    assert np.array_equal(K, ref["K"]) and np.array_equal(N, ref["N_Aa_i"]) and same(king, ref["king"])


@pytest.mark.parametrize("layout", [0, 1], ids=["column_major", "transposed_t4"])
def test_gather_pack_kernel_edges(layout, tpg_cfg):
    """Row spans that start / end at the first / last byte of the allocation, n not a multiple of 16,
    duplicated rows, a single locus, loci counts that are not multiples of 4: dense result equals the oracle,
    for both staging layouts of gather_pack."""
    tpg_cfg(gather_layout=layout)
    rng = np.random.default_rng(53)
    for n_tot, m_tot, n_sel in ((17, 1, 5), (64, 3, 64), (1000, 40, 999), (4099, 7, 33), (10000, 70, 2000)):
        raw = random_bk(rng, n_tot, m_tot, 0.1, codes=(0, 1, 2, 3, 7, 200))
        rows = rng.choice(np.arange(1, n_tot + 1), size=n_sel, replace=True).astype(np.int32)
        rows[0], rows[-1] = 1, n_tot
        X, src = device_fbm(raw, n_tot, m_tot)
        ref = O.snp_ibs(raw, n_tot, m_tot, rows, None, type="raw_counts")
        got = tpg.snp_ibs(X, rows, None, type="raw_counts", src=src)
        assert np.array_equal(got["ibs"][:], ref["ibs"]) and np.array_equal(got["valid_n"][:], ref["valid_n"]), (n_tot, m_tot)


# ------------------------------------------------------------------ file reader

@pytest.mark.parametrize("cols_kind", ["all", "window", "every3"])
def test_o_direct_reader(cols_kind, tpg_cfg, tmp_path):
    """io_mode = 1 reads the .bk with O_DIRECT (4 KiB-aligned offsets, lengths and buffers; unaligned
    columns, a short tail): same bytes, same result as the page-cache reader and the mmap."""
    rng = np.random.default_rng(61)
    n, m = 1013, 4100                                    # 1013-byte columns: nothing is aligned
    raw = random_bk(rng, n, m, 0.05)
    X = tpg.FBMCode256.from_bytes(raw, n, m, backingfile=tmp_path / "dio.bk")
    cols = pick(cols_kind, m, rng)
    ref = O.snp_king(raw, n, m, None, cols)["king"]
    for mode in (0, 1, 2):
        tpg_cfg(io_mode=mode)
        king = tpg.snp_king(X, None, cols, gpu_chunk_loci=1024)
        st = dict(tpg.snp.last_stats)
        assert same(king, ref), mode
        if mode == 1:
            assert st["n_direct_io"] in (0, st["n_chunks"])   # 0 where the filesystem refuses O_DIRECT
        if mode != 2:
            assert st["bytes_file"] >= (m if cols is None else len(cols)) * n
    win = np.arange(100, 300, dtype=np.int32)            # a short row window of the columns
    tpg_cfg(io_mode=0)
    assert same(tpg.snp_king(X, win, cols), O.snp_king(raw, n, m, win, cols)["king"])


# ------------------------------------------------------------------ finish arithmetic

def test_grm_mb_is_the_double_double_mean():
    """pairwise_grm's mb (R/pairwise_grm.R:42-50, R's long-double mean) comes from the same double-double
    reduction as indiv_inbreeding / pop_fst: GRM and the WG17 consumers report ONE mb."""
    rng = np.random.default_rng(71)
    n, m = 333, 2500
    raw = random_bk(rng, n, m, 0.1)
    X = fbm_of(raw, n, m)
    a = tpg.snp_allele_sharing(X)
    g, mb = tpg.grm_from_allele_sharing(a)
    ref_g, ref_mb = O.grm_from_as(a)
    assert mb == ref_mb                                   # to the last bit of the double
    ok = ~np.isnan(ref_g)
    assert np.allclose(g[ok], ref_g[ok], rtol=0, atol=1e-14)
    from tidypopgen_b200 import consumers
    _, _, _, mbs = consumers._wg17(a, None, 1, want_inb=True)
    assert mbs[0] == mb


@pytest.mark.parametrize("route", [0, L.FLAG_GRAM_I8, L.FLAG_GRAM_I8 | L.FLAG_PLANES_K_MAJOR], ids=["mxf4", "i8", "i8_kmajor"])
@pytest.mark.parametrize("n_tot,rk", [(389, "all"), (929, "shuffle"), (1300, "subset")])
def test_wide_gram_kernel_whole_jobs(n_tot, rk, route, tpg_cfg):
    """gram_mxf4_wide_kernel (256 x 480 tiles) and gram_i8_wide_kernel (256 x 512) forced on small matrices (ragged
    last tiles: 240 + 149, 480 + 449, 256 + 133, ...): every statistic equals the oracle, for one chunk and several,
    with and without a row gather."""
    rng = np.random.default_rng(n_tot)
    m_tot = 2100
    raw = random_bk(rng, n_tot, m_tot, 0.08, codes=(0, 1, 2, 3, 4, 5, 6, 255))
    X = fbm_of(raw, n_tot, m_tot)
    rows = pick(rk, n_tot, rng)
    tpg_cfg(gram_wide=2)
    ref_k = O.snp_king(raw, n_tot, m_tot, rows, None)
    for chunk in (0, 512):
        king, K, N = tpg.snp_king_counts(X, rows, None, gpu_chunk_loci=chunk, flags=route)
        assert (tpg.snp.last_stats["n_mxf4_chunks"] >= 1) == (route == 0)
        assert np.array_equal(K, ref_k["K"]) and np.array_equal(N, ref_k["N_Aa_i"]) and same(king, ref_k["king"])
    ref_i = O.snp_ibs(raw, n_tot, m_tot, rows, None, type="raw_counts")
    got = tpg.snp_ibs(X, rows, None, type="raw_counts", flags=route)
    assert np.array_equal(got["ibs"][:], ref_i["ibs"]) and np.array_equal(got["valid_n"][:], ref_i["valid_n"])
    ref_a = O.snp_allele_sharing(raw, n_tot, m_tot, rows, None, block_size_=500)
    a, num, den = tpg.snp_allele_sharing_counts(X, rows, None, block_size=500, flags=route)
    assert np.array_equal(num, ref_a["num"]) and np.array_equal(den, ref_a["denom"]) and same(a, ref_a["as"])


@pytest.mark.parametrize("shard", ["loci", "tiles"])
def test_multi_gpu_large_result_comes_home_over_every_link(shard):
    """tpg_pairwise_multi with a result of >= 256 MB: finished on the first device into device buffers, then
    every device pulls its slab of columns over NVLink and sends it home over its own PCIe link; the count
    matrices are summed block-wise (upper blocks of the symmetric products).  Identical to one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(91)
    n, m = 6100, 2048
    raw = random_bk(rng, n, m, 0.05)
    X = fbm_of(raw, n, m)
    devices = list(range(min(4, torch.cuda.device_count())))
    ref = tpg.snp_king(X)
    got = tpg.snp_king(X, devices=devices, shard=shard)
    assert same(got, ref)
    many1 = tpg.snp_pairwise_many(X, ("king", "ibs_raw", "allele_sharing"), block_size=500)
    manyN = tpg.snp_pairwise_many(X, ("king", "ibs_raw", "allele_sharing"), block_size=500, devices=devices, shard=shard)
    assert same(manyN["king"], many1["king"]) and same(manyN["allele_sharing"], many1["allele_sharing"])
    assert np.array_equal(manyN["ibs_raw"]["ibs"], many1["ibs_raw"]["ibs"])
    assert np.array_equal(manyN["ibs_raw"]["valid_n"], many1["ibs_raw"]["valid_n"])


def test_mma_peak_probe_reports_a_plausible_issue_rate():
    """tpg_mma_peak (csrc/peak.cu): the roofline denominators bench.py reports — the 4-bit kind issues about
    twice the MACs of kind::i8, both within reach of the nominal 9 / 4.5 POP/s."""
    import ctypes as C
    eb = L.errbuf()
    got = {}
    for kind in (0, 1):
        tops, ms = C.c_double(0), C.c_double(0)
        L.check(L.lib().tpg_mma_peak(0, kind, 0.02, C.byref(tops), C.byref(ms), eb, len(eb)), eb)
        got[kind] = tops.value
        assert ms.value > 1.0
    assert 2500 < got[0] < 5200 and 5000 < got[1] < 10400 and 1.6 < got[1] / got[0] < 2.4, got
    assert L.lib().tpg_mma_peak(0, 7, 0.02, None, None, eb, len(eb)) == L.ERR_ARG


def test_bench_line_at_a_reduced_shape():
    """bench.py end to end at a development size (labelled as such in the line): the contract's keys, the parity
    flags, the e2e leg from a .bk file and the extras all run; the numbers themselves are not the point."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--indiv", "3000", "--loci", "16384", "--steps", "2",
                        "--warmup", "1", "--pairs", "8", "--no-extras"], capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["unit"] == "pair-SNP/s" and d["scaling"] == "strong" and d["n_gpus"] == 1 and d["value"] > 0
    assert "REDUCED" in d["config"]["workload"]
    for key in ("roofline", "roofline_int8", "roofline_decode", "cpu_baseline", "e2e", "clocks", "parity", "mma_peaks_measured"):
        assert d[key] is not None, key
    assert d["roofline"]["bound"] == "tensor" and 0 < d["roofline"]["frac"] < 1.2 and d["roofline"]["peak"] > 1000
    assert d["parity"]["sampled_pairs_exact"] and d["parity"]["int8_and_4bit_routes_agree"] and d["parity"]["parity_first_block"]
    assert d["e2e"]["sampled_pairs_exact"] and d["e2e"]["matches_device_run"] and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["gpu_launches"] >= 2 * 4
