"""Timings of the consumers of the pairwise matrices on one B200:
code256 scans at the HGDP shape (HBM-bound: 1 B / genotype), filter_high_relatedness and the
WG17 statistics on matrices resident in HBM.  One JSON line per measurement."""
import ctypes as C, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidypopgen_b200 as tpg
from tidypopgen_b200 import _lib as L
from tidypopgen_b200 import dist as tdist

HBM = 6459.6
try:
    HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
eng = tdist.CudaEngine(0)
lib, eb = L.lib(), L.errbuf()


def synth(n, m, seed, miss):
    raw = torch.empty(n * m + 64, dtype=torch.uint8, device="cuda:0")
    L.check(lib.tpg_synth_fill(eng.ctx.handle, raw.data_ptr(), n, m, 0, seed, miss, 0, eb, len(eb)), eb)
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<synthetic>", n, m, None
    return X, raw


# ---- scans at C2, device-resident bytes
n, m = 929, 650000
X, raw = synth(n, m, 20260002, 0.02)
src = (L.SRC_DEVICE, raw.data_ptr(), raw)
for _ in range(3):
    r = tpg.scan_counts(X, src=src)
best = min(tpg.scan_counts(X, src=src)["stats"]["ms_decode"] for _ in range(5))
print(json.dumps({"bench": "scan_counts (gt_ind_hetero + alt_freq + loci NA counts, one pass)", "n": n, "m": m,
                  "kernel_ms": best, "GBps": n * m / best / 1e6, "hbm_peak_GBps": HBM, "frac": n * m / best / 1e6 / HBM,
                  "algorithmic_bytes_per_genotype": 1}), flush=True)
host = raw[: n * m].cpu().numpy()
het = (host.reshape(m, n) == 1).sum(0)
assert np.array_equal(r["het_n"], het), "scan mismatch"

# ---- KING + filter with the matrix kept in HBM
for (n, m, thr) in [(5000, 20000, 0.01), (20000, 20000, 0.015), (20000, 20000, 0.088)]:
    X, raw = synth(n, m, 7, 0.01)
    gt = tpg.gen_tibble(X)
    kw = dict(src=(L.SRC_DEVICE, raw.data_ptr(), raw))
    tpg.king_filter(gt, thr, **kw)
    t0 = time.perf_counter()
    rel, _ = tpg.king_filter(gt, thr, **kw)
    wall = time.perf_counter() - t0
    st = tpg.snp.last_stats
    print(json.dumps({"bench": "king_filter (KING + filter_high_relatedness, matrix stays in HBM)", "n": n, "m": m,
                      "threshold": thr, "removed": int((~rel[2]).sum()), "wall_ms": wall * 1e3, "gram_ms": st["ms_gram"],
                      "decode_ms": st["ms_decode"], "filter_ms": st["ms_finish"],
                      "matrix_GB": 8 * n * n / 1e9}), flush=True)
    del raw

# ---- WG17 statistics: fused (allele sharing computed and reduced in HBM) and from a host matrix
n, m, npop = 8000, 20000, 50
X, raw = synth(n, m, 11, 0.02)
kw = dict(src=(L.SRC_DEVICE, raw.data_ptr(), raw))
rng = np.random.default_rng(0)
grp = np.sort(rng.integers(0, npop, n)).astype(np.int32)
grp[:npop] = np.arange(npop)
gt = tpg.gen_tibble(X, population=[f"p{g:03d}" for g in grp])
tpg.pop_fst(gt, **kw)
t0 = time.perf_counter()
fst = tpg.pop_fst(gt, include_global=True, **kw)
wall = time.perf_counter() - t0
st = tpg.snp.last_stats
print(json.dumps({"bench": "pop_fst (WG17) fused: allele-sharing job + block means, matrix stays in HBM", "n": n, "m": m,
                  "npop": npop, "wall_ms": wall * 1e3, "gram_ms": st["ms_gram"], "decode_ms": st["ms_decode"],
                  "reduce_ms": st["ms_finish"], "fst_global": fst["global"]}), flush=True)
a = tpg.snp_allele_sharing(X, **kw)
t0 = time.perf_counter()
fst2 = tpg.pop_fst(gt, include_global=True, allele_sharing_mat=a)
wall = time.perf_counter() - t0
print(json.dumps({"bench": "pop_fst (WG17) from a host allele-sharing matrix (H2D of 8 n^2 bytes included)", "n": n,
                  "npop": npop, "wall_ms": wall * 1e3, "same_as_fused": bool(fst2["global"] == fst["global"])}), flush=True)
