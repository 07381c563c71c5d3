import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidypopgen_b200 as tpg
from tidypopgen_b200 import _lib as L
from tidypopgen_b200 import dist as tdist
eng = tdist.CudaEngine(0); lib, eb = L.lib(), L.errbuf()
n, m = 929, 650000
raw = torch.empty(n * m + 64, dtype=torch.uint8, device="cuda:0")
L.check(lib.tpg_synth_fill(eng.ctx.handle, raw.data_ptr(), n, m, 0, 20260002, 0.02, 0, eb, len(eb)), eb)
X = tpg.FBMCode256.__new__(tpg.FBMCode256); X.backingfile, X.nrow, X.ncol, X.code256 = "<synthetic>", n, m, None
src = (L.SRC_DEVICE, raw.data_ptr(), raw)
for _ in range(3): tpg.scan_counts(X, src=src)
print(os.environ.get("TPG_SCAN_LOCI"), os.environ.get("TPG_SCAN_NO_INDIV"), min(tpg.scan_counts(X, src=src)["stats"]["ms_decode"] for _ in range(5)))
