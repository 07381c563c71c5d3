"""L2 neighbourhood of the 4-bit wide kernel at C4 (tpg_config.gram_l2_rows x gram_l2_cols, in 256-row / 240-column
tiles) and k-blocks per item: Gram ms, best of 3.  One JSON line per variant."""
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
counts = eng.new_counts(n)
job, keep = tpg.snp.make_job(X, None, None, L.STAT_KING, 2684, src=src)
variants = [(0, 0, 0), (8, 16, 0), (16, 8, 0), (16, 16, 0), (4, 16, 0), (12, 12, 0), (8, 12, 0), (6, 8, 0), (8, 8, 384), (8, 16, 384), (16, 16, 128), (0, 0, 0)]
for r, c, kb in variants:
    tpg.config(reset=True, gram_l2_rows=r, gram_l2_cols=c, gram_max_kblocks=kb)
    best = 1e30
    for k in range(4):
        st = L.Stats()
        L.check(lib.tpg_accumulate(eng.ctx.handle, C.byref(job), counts.data_ptr(), C.byref(st), eb, len(eb)), eb)
        torch.cuda.synchronize()
        if k:
            best = min(best, st.ms_gram)
    print(json.dumps({"gram_l2_rows": r, "gram_l2_cols": c, "gram_max_kblocks": kb, "gram_ms_best": round(best, 2)}), flush=True)
