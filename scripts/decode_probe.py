"""decode_planes_k4 alone at the C2 shape (929 x 650,000) for build variants (tpg_config.decode_k4_v1)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidypopgen_b200 as tpg
from tidypopgen_b200 import _lib as L
from tidypopgen_b200 import dist as tdist
eng = tdist.CudaEngine(0)
lib, eb = L.lib(), L.errbuf()
for n, m in ((929, 650000), (10000, 200000)):
    raw = torch.empty(n * m + 64, dtype=torch.uint8, device="cuda")
    L.check(lib.tpg_synth_fill(eng.ctx.handle, raw.data_ptr(), n, m, 0, 20260002, 0.02, 0, eb, len(eb)), eb)
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<synthetic>", n, m, None
    src = (L.SRC_DEVICE, raw.data_ptr(), raw)
    for variant in (0, 2, 1):
        tpg.config(decode_k4_v1=variant)
        best = 1e9
        for _ in range(6):
            tpg.snp_king(X, src=src)
            best = min(best, tpg.snp.last_stats["ms_decode"])
        print(json.dumps({"n": n, "m": m, "variant": variant, "decode_ms": best, "GBps": 2.5 * n * m / best / 1e6}), flush=True)
    del raw
