"""Biobank shape (BASELINE.json configs[3]) end to end from ONE process: 50,000 x 500,000 genotypes in
pinned host memory -> tpg_pairwise_multi over every GPU of the box -> 50,000 x 50,000 KING matrix in
host memory.  Sampled entries are checked exactly against the reference formula.  One JSON line."""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidypopgen_b200 as tpg
from tidypopgen_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--indiv", dest="n", type=int, default=50000)
ap.add_argument("--loci", dest="m", type=int, default=500000)
ap.add_argument("--shard", default="auto")
ap.add_argument("--pairs", type=int, default=48)
args = ap.parse_args()
n, m = args.n, args.m
ndev = torch.cuda.device_count()
ctx = L.Context(0)
eb = L.errbuf()
dev = torch.empty(n * m, dtype=torch.uint8, device="cuda:0")
L.check(L.lib().tpg_synth_fill(ctx.handle, dev.data_ptr(), n, m, 0, 20260004, 0.01, 0, eb, len(eb)), eb)
t0 = time.perf_counter()
host = torch.empty(n * m, dtype=torch.uint8, pin_memory=True)
host.copy_(dev)
torch.cuda.synchronize()
t_pin = time.perf_counter() - t0
del dev
ctx = None
torch.cuda.empty_cache()
X = tpg.FBMCode256.__new__(tpg.FBMCode256)
X.backingfile, X.nrow, X.ncol, X.code256 = "<pinned host>", n, m, None
src = (L.SRC_HOST, host.data_ptr(), host)
out = np.zeros((n, n), np.float64, order="F")
times = []
for rep in range(2):
    t0 = time.perf_counter()
    tpg.snp.run_job(X, None, None, L.STAT_KING, tpg.block_size(n), out0=out, devices=list(range(ndev)), shard=args.shard, src=src)
    times.append(time.perf_counter() - t0)
st = dict(tpg.snp.last_stats)
cols = host.view(m, n)
rng = np.random.default_rng(3)
bad = 0
for k in range(args.pairs):
    i, j = (int(x) for x in rng.integers(0, n, 2))
    if k % 6 == 0:
        j = i
    gi, gj = cols[:, i].numpy(), cols[:, j].numpy()
    h0i, h1i, h2i, vi = gi == 0, gi == 1, gi == 2, gi <= 2
    h0j, h1j, h2j, vj = gj == 0, gj == 1, gj == 2, gj <= 2
    K = float((h1i & h1j).sum()) - 2.0 * float((h0i & h2j).sum() + (h2i & h0j).sum())
    Ni, Nj = float((h1i & vj).sum()), float((vi & h1j).sum())
    pm = min(Ni, Nj)
    with np.errstate(all="ignore"):
        ref = np.float64(K) / np.float64(2 * pm) + 0.5 - np.float64(0.25 * (Ni + Nj)) / np.float64(pm)
    got = out[i, j]
    ok = (got == ref or (np.isnan(got) and np.isnan(ref))) and (out[j, i] == got or np.isnan(got))
    bad += 0 if ok else 1
print(json.dumps({"check": "c4_inprocess_e2e", "n": n, "m": m, "devices": ndev, "shard": args.shard,
                  "first_call_s": times[0], "second_call_s": times[1], "pair_snp_per_s": n * n * m / times[1],
                  "host_bytes_GB": n * m / 1e9, "result_GB": 8 * n * n / 1e9, "pin_and_copy_s": t_pin,
                  "stats_max_over_devices_ms": {k: st[k] for k in ("ms_read", "ms_h2d", "ms_decode", "ms_gram", "ms_finish", "ms_d2h")},
                  "h2d_GB": st["bytes_h2d"] / 1e9, "packed_chunks": st["n_packed_chunks"], "chunks": st["n_chunks"],
                  "mismatches": bad, "ok": bad == 0}), flush=True)
sys.exit(0 if bad == 0 else 1)
