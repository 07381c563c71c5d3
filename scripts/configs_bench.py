"""Stage timings of the other BASELINE.json configs on one B200, device-resident synthetic bytes:
C3 (2504 x 1,000,000 GRM, imputed bytes) and C5 (10,000 x 1,000,000, IBS on a shuffled 70 % row
subset and every-3rd + shuffled-tail loci).  One JSON line per job."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidypopgen_b200 as tpg
from tidypopgen_b200 import _lib as L

ctx = L.Context(0)
eb = L.errbuf()


def fbm(n, m):
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<device>", n, m, None
    return X


def timed(name, fn, n, m, ops):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    wall = time.perf_counter() - t0
    st = tpg.snp.last_stats
    print(json.dumps({"job": name, "n": n, "m": m, "wall_ms": wall * 1e3, "decode_ms": st["ms_decode"], "gram_ms": st["ms_gram"],
                      "finish_ms": st["ms_finish"], "d2h_ms": st["ms_d2h"], "chunks": st["n_chunks"],
                      "gram_alg_TOPS": ops * n * n * m / (st["ms_gram"] * 1e-3) / 1e12,
                      "route": "4-bit (kind::mxf4)" if st["n_mxf4_chunks"] else "int8",
                      # 1 B read + one byte (int8 route) or half a byte (4-bit route) written per plane and genotype
                      "decode_GBps": (1 + (2 if ops == 2 else 3) * (0.5 if st["n_mxf4_chunks"] else 1.0)) * n * m
                                     / (st["ms_decode"] * 1e-3) / 1e9}), flush=True)


# C3
n, m = 2504, 1_000_000
dev = torch.empty(n * m, dtype=torch.uint8, device="cuda")
L.check(L.lib().tpg_synth_fill(ctx.handle, dev.data_ptr(), n, m, 0, 20260003, 0.01, 1, eb, len(eb)), eb)
src = (L.SRC_DEVICE, dev.data_ptr(), dev)
timed("C3 pairwise_grm 2504 x 1e6 (imputed bytes)", lambda: tpg.snp_grm(fbm(n, m), src=src), n, m, 2)
timed("C3 pairwise_king 2504 x 1e6", lambda: tpg.snp_king(fbm(n, m), src=src), n, m, 4)
del dev
# C5
n_tot, m_tot = 10_000, 1_000_000
dev = torch.empty(n_tot * m_tot, dtype=torch.uint8, device="cuda")
L.check(L.lib().tpg_synth_fill(ctx.handle, dev.data_ptr(), n_tot, m_tot, 0, 20260005, 0.10, 2, eb, len(eb)), eb)
src = (L.SRC_DEVICE, dev.data_ptr(), dev)
rng = np.random.default_rng(5)
rows = (rng.permutation(n_tot)[: n_tot * 7 // 10] + 1).astype(np.int32)
tail = rng.permutation(m_tot)[: m_tot // 20]
cols = (np.concatenate([np.arange(0, m_tot, 3), tail]) + 1).astype(np.int32)
timed("C5 snp_ibs proportion, 7000 shuffled rows x 383k gathered loci of 10,000 x 1e6",
      lambda: tpg.snp_ibs(fbm(n_tot, m_tot), rows, cols, type="proportion", src=src), len(rows), len(cols), 3)
timed("C5 snp_ibs proportion, all rows x all loci 10,000 x 1e6",
      lambda: tpg.snp_ibs(fbm(n_tot, m_tot), type="proportion", src=src), n_tot, m_tot, 3)
