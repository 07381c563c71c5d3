"""scan_counts on the C2 shape (929 x 650,000) and a biobank-width one (50,000 x 20,000): kernel time of each build
(tpg_config.scan_table: 1 lookup table, 2 table-free 16-locus rounds, 3 table-free warp per locus)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidypopgen_b200 as tpg
from tidypopgen_b200 import _lib as L
from tidypopgen_b200 import dist as tdist

eng = tdist.CudaEngine(0)
lib, eb = L.lib(), L.errbuf()
for n, m in ((929, 650000), (50000, 20000)):
    raw = torch.empty(n * m + 64, dtype=torch.uint8, device="cuda:0")
    L.check(lib.tpg_synth_fill(eng.ctx.handle, raw.data_ptr(), n, m, 0, 20260002, 0.02, 0, eb, len(eb)), eb)
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<synthetic>", n, m, None
    src = (L.SRC_DEVICE, raw.data_ptr(), raw)
    ref = None
    for kern, lpc in ((1, 0), (2, 0), (3, 0), (3, 256), (3, 512), (3, 1024), (3, 2032)):
        tpg.config(reset=True, scan_table=kern, scan_loci_per_cta=lpc)
        for _ in range(3):
            r = tpg.scan_counts(X, src=src)
        ms = min(tpg.scan_counts(X, src=src)["stats"]["ms_decode"] for _ in range(7))
        key = (r["het_n"].tobytes(), r["na_n"].tobytes(), r["n_valid"].tobytes(), r["loc0"].tobytes())
        ref = ref or key
        print(json.dumps({"n": n, "m": m, "scan_table": kern, "scan_loci_per_cta": lpc, "kernel_ms": round(ms, 4),
                          "GBps": round(n * m / ms / 1e6, 1), "same_as_lookup_kernel": key == ref}), flush=True)
    del raw
