"""In-process multi-GPU driver (tpg_pairwise_multi) on every GPU of the box, from a .bk file:
loci- and tile-sharded KING / IBS / allele sharing against the single-GPU job.  One JSON line."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidypopgen_b200 as tpg  # noqa: E402
from tidypopgen_b200 import _lib as L  # noqa: E402


def main():
    ndev = torch.cuda.device_count()
    n, m = int(sys.argv[1]) if len(sys.argv) > 1 else 6000, int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    ctx = L.Context(0)
    eb = L.errbuf()
    dev = torch.empty(n * m, dtype=torch.uint8, device="cuda:0")
    L.check(L.lib().tpg_synth_fill(ctx.handle, dev.data_ptr(), n, m, 0, 777, 0.03, 0, eb, len(eb)), eb)
    same = lambda a, b: bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))  # noqa: E731
    out = {"check": "multi_driver", "n": n, "m": m, "devices": ndev, "ok": True}
    with tempfile.NamedTemporaryFile(suffix=".bk", prefix="tpg_multi_") as tf:
        dev.cpu().numpy().tofile(tf.name)
        del dev
        X = tpg.FBMCode256(tf.name, n, m)
        devices = list(range(ndev))
        for name, fn, kw in (("king", tpg.snp_king, {}), ("as", tpg.snp_allele_sharing, {"block_size": 7001})):
            t0 = time.perf_counter()
            ref = fn(X, **kw)
            t1 = time.perf_counter()
            out[f"{name}_1gpu_s"] = t1 - t0
            for shard in ("loci", "tiles"):
                fn(X, devices=devices, shard=shard, **kw)  # warm the per-rank contexts
                t0 = time.perf_counter()
                got = fn(X, devices=devices, shard=shard, **kw)
                out[f"{name}_{shard}_{ndev}gpu_s"] = time.perf_counter() - t0
                ok = same(got, ref)
                out[f"{name}_{shard}_ok"] = ok
                out["ok"] &= ok
    print(json.dumps(out), flush=True)
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
