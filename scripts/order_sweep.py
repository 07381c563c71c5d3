"""Gram work-list order / L2 neighbourhood sweep at C4 on one GPU: one process, bytes synthesized once, every
variant warmed once and timed over a few jobs (tpg_config is read when a job starts).  One JSON line per variant."""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidypopgen_b200 as tpg
from tidypopgen_b200 import _lib as L
from tidypopgen_b200 import dist as tdist

n, m = 50000, 500000
eng = tdist.CudaEngine(0)
lib, eb = L.lib(), L.errbuf()
raw = torch.empty(n * m + 64, dtype=torch.uint8, device="cuda")
L.check(lib.tpg_synth_fill(eng.ctx.handle, raw.data_ptr(), n, m, 0, 20260004, 0.01, 0, eb, len(eb)), eb)
X = tpg.FBMCode256.__new__(tpg.FBMCode256)
X.backingfile, X.nrow, X.ncol, X.code256 = "<synthetic>", n, m, None
src = (L.SRC_DEVICE, raw.data_ptr(), raw)
out = torch.zeros((n, n), dtype=torch.float64, device="cuda")
res = L.Result(out0=out.data_ptr(), out1=None, aux0=None, aux1=None, out_on_device=1)


def run(flags, steps):
    job, keep = tpg.snp.make_job(X, None, None, L.STAT_KING, 2684, src=src, flags=flags)
    best = 1e30
    for k in range(steps + 1):
        L.check(lib.tpg_pairwise_ctx(eng.ctx.handle, C.byref(job), C.byref(res), eb, len(eb)), eb)
        if k:
            best = min(best, res.stats.ms_gram)
    return best


variants = [(2, 4, 4), (1, 4, 4), (1, 4, 8), (1, 8, 8), (1, 8, 16), (1, 16, 8), (1, 16, 16), (1, 8, 32), (1, 32, 8)]
for route, flags, steps in (("mxf4", 0, 3), ("int8", L.FLAG_GRAM_I8, 2)):
    for order, r, c in variants:
        tpg.config(gram_order=order, gram_l2_rows=r, gram_l2_cols=c)
        print(json.dumps({"route": route, "order": "product-major" if order == 1 else "tile-major (round 1)", "l2_rows": r * 256,
                          "l2_cols": c * (240 if route == "mxf4" else 256), "gram_ms_best": run(flags, steps)}), flush=True)
