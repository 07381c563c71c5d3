"""Where the end-to-end time of a host-sourced job goes: tpg_stats of snp_king() on pinned host bytes
(the bench's e2e leg), plus the Python-side wall clock around the call.
    python scripts/e2e_profile.py [n m]      default: C2, 929 x 650000"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidypopgen_b200 as tpg
from tidypopgen_b200 import _lib as L
from tidypopgen_b200 import dist as tdist

n, m = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (929, 650000)
eng = tdist.CudaEngine(0)
lib, eb = L.lib(), L.errbuf()
dev_raw = torch.empty(n * m + 64, dtype=torch.uint8, device="cuda:0")
L.check(lib.tpg_synth_fill(eng.ctx.handle, dev_raw.data_ptr(), n, m, 0, 20260002, 0.02, 0, eb, len(eb)), eb)
host_raw = torch.empty(n * m, dtype=torch.uint8).pin_memory()
host_raw.copy_(dev_raw[: n * m])
X = tpg.FBMCode256.__new__(tpg.FBMCode256)
X.backingfile, X.nrow, X.ncol, X.code256 = "<synthetic>", n, m, None
bs = tpg.block_size(n)
src = (L.SRC_HOST, host_raw.data_ptr(), host_raw)
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    k = tpg.snp_king(X, block_size=bs, src=src, device=0)
    wall = (time.perf_counter() - t0) * 1e3
    st = tpg.snp.last_stats
    print(f"rep {rep}: wall {wall:.2f} ms | lib total {st['ms_total']:.2f} read/pack {st['ms_read']:.2f} h2d {st['ms_h2d']:.2f} "
          f"decode {st['ms_decode']:.2f} gram {st['ms_gram']:.2f} finish {st['ms_finish']:.3f} d2h {st['ms_d2h']:.2f} "
          f"chunks {st['n_chunks']} packed {st['n_packed_chunks']} mxf4 {st['n_mxf4_chunks']} h2d_MB {st['bytes_h2d']/1e6:.0f}", flush=True)
