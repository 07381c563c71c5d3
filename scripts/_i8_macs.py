"""Work-list schedule of the int8 wide Gram kernel at C4, whole job (world 1) and on one emulated tile-sharded rank
(world 2 / 8): k-blocks per item, L2 neighbourhood, order.  One JSON line per variant (Gram ms, best of 2)."""
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
variants = [dict(gram_wide=w, gram_max_kblocks=kb, gram_l2_rows=r, gram_l2_cols=c, gram_order=o)
            for (w, kb, r, c, o) in [(2,0,0,0,0),(1,0,0,0,0)]]
worlds = [(1, 0), (2, 0), (8, 3)] if len(sys.argv) < 2 else [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]]
for world, rank in worlds:
    job, keep = tpg.snp.make_job(X, None, None, L.STAT_KING, 2684, src=src,
                                 shard_mode=L.SHARD_TILES if world > 1 else L.SHARD_NONE, shard_rank=rank, shard_world=world)
    for v in variants:
        tpg.config(reset=True, gram_route=1, **v)
        best = 1e30
        for k in range(3):
            st = L.Stats()
            L.check(lib.tpg_accumulate(eng.ctx.handle, C.byref(job), counts.data_ptr(), C.byref(st), eb, len(eb)), eb)
            torch.cuda.synchronize()
            if k:
                best = min(best, st.ms_gram)
        print(json.dumps({"world": world, "rank": rank, **v, "gram_ms_best": round(best, 2), "wide_launches": st.n_wide_launches, "mma_macs": st.mma_macs}), flush=True)
