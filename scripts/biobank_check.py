"""Biobank-shape run (BASELINE.json configs[3]): pairwise_king on n individuals x m SNPs,
upper-triangle tile-sharded across the ranks of one node (no data-path collective), with an
exact parity check on sampled pairs.

    python scripts/biobank_check.py [--indiv 50000] [--loci 500000]                 # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29540 scripts/biobank_check.py --indiv 50000 --loci 500000       # 8 GPUs

Every rank generates the same synthetic .bk image on its GPU (counter-based RNG), streams all
loci, and computes + finishes only the 256x256 output blocks it owns.  The full-size oracle is
out of reach (8.75e15 FP64 flop), so parity is checked EXACTLY on sampled entries: for random
owned pairs (i, j) the two genotype columns are pulled to the host and the reference formula
(src/snp_king.cpp:45-72, R/snp_king.R:82-91) is evaluated in numpy; the GPU value must be
bit-identical.  Prints one JSON line per run (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidypopgen_b200 as tpg  # noqa: E402
from tidypopgen_b200 import _lib as L  # noqa: E402
from tidypopgen_b200 import dist as tdist  # noqa: E402


def king_pair(gi, gj):
    """KING-robust for one pair from raw bytes, the reference's arithmetic."""
    h0i, h1i, h2i, vi = gi == 0, gi == 1, gi == 2, gi <= 2
    h0j, h1j, h2j, vj = gj == 0, gj == 1, gj == 2, gj <= 2
    K = float((h1i & h1j).sum()) - 2.0 * float((h0i & h2j).sum() + (h2i & h0j).sum())
    Ni, Nj = float((h1i & vj).sum()), float((vi & h1j).sum())
    pm = min(Ni, Nj)
    with np.errstate(all="ignore"):
        return np.float64(K) / np.float64(2 * pm) + 0.5 - np.float64(0.25 * (Ni + Nj)) / np.float64(pm)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--indiv", dest="n", type=int, default=50000)
    ap.add_argument("--loci", dest="m", type=int, default=500000)
    ap.add_argument("--missing", type=float, default=0.01)
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--emulate-world", type=int, default=0, help="one process plays rank --emulate-rank of this many tile-sharded ranks")
    ap.add_argument("--emulate-rank", type=int, default=0)
    args = ap.parse_args()
    rank, world, local = tdist.init_process_group("nccl")
    emulated = args.emulate_world > 1 and world == 1
    if emulated:
        rank, world = args.emulate_rank, args.emulate_world
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    eng = tdist.CudaEngine(local)
    lib, eb = L.lib(), L.errbuf()
    n, m = args.n, args.m
    raw = torch.empty(n * m + 64, dtype=torch.uint8, device=dev)
    L.check(lib.tpg_synth_fill(eng.ctx.handle, raw.data_ptr(), n, m, 0, 20260004, args.missing, 0, eb, len(eb)), eb)
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<synthetic>", n, m, None
    job, keep = tpg.snp.make_job(X, None, None, L.STAT_KING, tpg.block_size(n), device=local,
                                 src=(L.SRC_DEVICE, raw.data_ptr(), raw),
                                 shard_mode=L.SHARD_TILES if world > 1 else L.SHARD_NONE, shard_rank=rank,
                                 shard_world=world)
    counts = eng.new_counts(n)
    out = torch.zeros((n, n), dtype=torch.float64, device=dev)   # column-major storage of KING
    res = L.Result(out0=out.data_ptr(), out1=None, aux0=None, aux1=None, out_on_device=1)
    torch.cuda.synchronize()
    if world > 1 and not emulated:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(eng.stream)
    st = L.Stats()
    L.check(lib.tpg_accumulate(eng.ctx.handle, C.byref(job), counts.data_ptr(), C.byref(st), eb, len(eb)), eb)
    L.check(lib.tpg_finish(eng.ctx.handle, C.byref(job), counts.data_ptr(), C.byref(res), eb, len(eb)), eb)
    e1.record(eng.stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1 and not emulated:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)

    # ---- exact parity on sampled owned pairs (plus a few diagonal ones)
    rng = np.random.default_rng(1000 + rank)
    nb = (n + 255) // 256
    bad, checked = 0, 0
    cols = raw[: n * m].view(m, n)
    while checked < args.pairs:
        i, j = (int(x) for x in rng.integers(0, n, 2))
        if checked % 8 == 0:
            j = i
        bi, bj = sorted((i // 256, j // 256))
        if world > 1 and tdist.block_owner(bi, bj, nb, world) != rank:
            continue
        gi, gj = cols[:, i].cpu().numpy(), cols[:, j].cpu().numpy()
        ref = king_pair(gi, gj)
        got = out[j, i].item()   # column-major: element (i, j) lives at [j, i]
        got_t = out[i, j].item()
        ok = (got == ref or (np.isnan(got) and np.isnan(ref))) and (got_t == got or (np.isnan(got_t) and np.isnan(got)))
        bad += 0 if ok else 1
        checked += 1
    flag = torch.tensor([bad], device=dev)
    if world > 1 and not emulated:
        torch.distributed.all_reduce(flag)
    if rank == 0 or emulated:
        msv = float(ms.item())
        print(json.dumps({
            "check": "biobank_king", "n": n, "m": m, "n_gpus": world, "emulated_rank": rank if emulated else None, "sharding": "tiles" if world > 1 else "none",
            "ms": msv, "wall_s": wall, "pair_snp_per_s": n * n * m / (msv * 1e-3),
            "int8_alg_TOPS_per_gpu": 4.0 * n * n * m / (msv * 1e-3) / 1e12 / world,
            "stage_ms_rank0": {"decode": st.ms_decode, "gram": st.ms_gram, "chunks": st.n_chunks},
            "pairs_checked_per_rank": args.pairs, "mismatches": int(flag.item()), "ok": int(flag.item()) == 0}), flush=True)
    if world > 1 and not emulated:
        torch.distributed.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
