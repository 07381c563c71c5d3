"""C2 (929 x 650,000 KING, device-resident bytes): decode / Gram kernel times under work-list variants."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidypopgen_b200 as tpg
from tidypopgen_b200 import _lib as L
from tidypopgen_b200 import dist as tdist

n, m = 929, 650000
eng = tdist.CudaEngine(0)
lib, eb = L.lib(), L.errbuf()
raw = torch.empty(n * m + 64, dtype=torch.uint8, device="cuda")
L.check(lib.tpg_synth_fill(eng.ctx.handle, raw.data_ptr(), n, m, 0, 20260002, 0.02, 0, eb, len(eb)), eb)
X = tpg.FBMCode256.__new__(tpg.FBMCode256)
X.backingfile, X.nrow, X.ncol, X.code256 = "<synthetic>", n, m, None
src = (L.SRC_DEVICE, raw.data_ptr(), raw)
variants = [dict(), dict(gram_wide=2), dict(gram_max_kblocks=64), dict(gram_max_kblocks=256), dict(gram_max_kblocks=512),
            dict(gram_ksplit=2), dict(gram_ksplit=4), dict(gram_ksplit=37), dict(gram_ksplit=74), dict(gram_route=1),
            dict(gram_route=1, gram_wide=2), dict(gram_order=1), dict(gram_order=2)]
for v in variants:
    tpg.config(reset=True, **v)
    best = None
    for k in range(6):
        r = tpg.snp.run_job(X, None, None, L.STAT_KING, 2684, src=src, out_on_device=True) if False else None
    # the staged C ABI keeps the result on the device: only kernel times matter here
    job, keep = tpg.snp.make_job(X, None, None, L.STAT_KING, 2684, src=src)
    counts = eng.new_counts(n)
    g, d = 1e9, 1e9
    for k in range(8):
        st = L.Stats()
        L.check(lib.tpg_accumulate(eng.ctx.handle, __import__("ctypes").byref(job), counts.data_ptr(), __import__("ctypes").byref(st), eb, len(eb)), eb)
        torch.cuda.synchronize()
        if k >= 2:
            g, d = min(g, st.ms_gram), min(d, st.ms_decode)
    print(json.dumps({**v, "decode_ms": round(d, 4), "gram_ms": round(g, 4), "macs": st.mma_macs, "wide": st.n_wide_launches}), flush=True)
