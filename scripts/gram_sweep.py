"""k-blocks per work item (tpg_config.gram_max_kblocks) at C4 on one GPU, 4-bit wide kernel: longer items amortise the
exposed epilogue, shorter ones keep the operand panels L2-resident.  One JSON line per variant."""
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
job, keep = tpg.snp.make_job(X, None, None, L.STAT_KING, 2684, src=src)
for kb in (128, 96, 192, 256, 384, 128):
    tpg.config(gram_max_kblocks=kb)
    best = 1e30
    for k in range(4):
        L.check(lib.tpg_pairwise_ctx(eng.ctx.handle, C.byref(job), C.byref(res), eb, len(eb)), eb)
        if k:
            best = min(best, res.stats.ms_gram)
    print(json.dumps({"gram_max_kblocks": kb, "gram_ms_best": best}), flush=True)
