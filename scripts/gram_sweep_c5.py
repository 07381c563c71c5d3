"""Gram scheduling variants at the C5 shape (10,000 x 1,000,000 IBS) and at C4: tpg_config gram_order / gram_l2_* /
gram_max_kblocks.  One JSON line per variant (best of 3 after a warm-up)."""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidypopgen_b200 as tpg
from tidypopgen_b200 import _lib as L
from tidypopgen_b200 import dist as tdist

eng = tdist.CudaEngine(0)
lib, eb = L.lib(), L.errbuf()


def sweep(n, m, stat, seed, miss, mode, variants, reps):
    raw = torch.empty(n * m + 64, dtype=torch.uint8, device="cuda")
    L.check(lib.tpg_synth_fill(eng.ctx.handle, raw.data_ptr(), n, m, 0, seed, miss, mode, eb, len(eb)), eb)
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<synthetic>", n, m, None
    src = (L.SRC_DEVICE, raw.data_ptr(), raw)
    out = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    res = L.Result(out0=out.data_ptr(), out1=None, aux0=None, aux1=None, out_on_device=1)
    job, keep = tpg.snp.make_job(X, None, None, stat, 1000, src=src)
    for order, r, c, kb in variants:
        tpg.config(gram_order=order, gram_l2_rows=r, gram_l2_cols=c, gram_max_kblocks=kb)
        best = 1e30
        for k in range(reps + 1):
            L.check(lib.tpg_pairwise_ctx(eng.ctx.handle, C.byref(job), C.byref(res), eb, len(eb)), eb)
            if k:
                best = min(best, res.stats.ms_gram)
        print(json.dumps({"n": n, "m": m, "order": order, "l2_rows": r, "l2_cols": c, "max_kblocks": kb, "gram_ms_best": best}), flush=True)
    del raw, out
    lib.tpg_ctx_trim(eng.ctx.handle)
    torch.cuda.empty_cache()


V = [(2, 4, 4, 128), (1, 4, 4, 128), (1, 8, 8, 128), (2, 4, 4, 256), (1, 4, 4, 256), (1, 8, 8, 256), (2, 8, 8, 256), (1, 4, 8, 192), (2, 4, 4, 192)]
sweep(10000, 1000000, L.STAT_IBS_PROPORTION, 20260005, 0.10, 2, V, 3)
sweep(50000, 500000, L.STAT_KING, 20260004, 0.01, 0, [(2, 4, 4, 128), (2, 4, 4, 256), (1, 4, 4, 256), (1, 8, 8, 256), (2, 4, 4, 192)], 2)
