"""One process for the profiler: KING on a C4-like shape (20,000 x 131,072: the wide kernels) and on C2 (929 x 650,000:
the ping-pong kernels), both tensor routes, device-resident bytes.  Run plain first, then under
  ncu --set full --clock-control none --import-source on -k regex:'gram_|decode_planes_k4' -o gpurun_out/r02_final ..."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidypopgen_b200 as tpg
from tidypopgen_b200 import _lib as L
from tidypopgen_b200 import dist as tdist

eng = tdist.CudaEngine(0)
lib, eb = L.lib(), L.errbuf()
for n, m, seed in ((20000, 131072, 20260004), (929, 650000, 20260002)):
    raw = torch.empty(n * m + 64, dtype=torch.uint8, device="cuda")
    L.check(lib.tpg_synth_fill(eng.ctx.handle, raw.data_ptr(), n, m, 0, seed, 0.01, 0, eb, len(eb)), eb)
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<synthetic>", n, m, None
    out = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    res = L.Result(out0=out.data_ptr(), out1=None, aux0=None, aux1=None, out_on_device=1)
    for route in (0, 1):
        tpg.config(reset=True, gram_route=route)
        job, keep = tpg.snp.make_job(X, None, None, L.STAT_KING, tpg.block_size(n), src=(L.SRC_DEVICE, raw.data_ptr(), raw))
        L.check(lib.tpg_pairwise_ctx(eng.ctx.handle, C.byref(job), C.byref(res), eb, len(eb)), eb)
        torch.cuda.synchronize()
        print(n, m, "route", route, "gram_ms", round(res.stats.ms_gram, 3), "decode_ms", round(res.stats.ms_decode, 3), "wide", res.stats.n_wide_launches, flush=True)
    del raw, out
    lib.tpg_ctx_trim(eng.ctx.handle)
    torch.cuda.empty_cache()
