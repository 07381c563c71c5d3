"""torchrun target: loci-sharded (NCCL all-reduce) and tile-sharded jobs across the ranks of
one node, checked against the same job run unsharded on rank 0's GPU.  Prints DIST_CHECK_OK."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidypopgen_b200 as tpg  # noqa: E402
from tidypopgen_b200 import _lib as L  # noqa: E402
from tidypopgen_b200 import dist as tdist  # noqa: E402


def main():
    rank, world, local = tdist.init_process_group("nccl")
    eng = tdist.CudaEngine(local)
    n, m = 1500, 40000
    raw = np.zeros(n * m, np.uint8)
    L.lib().tpg_synth_fill_host(raw.ctypes.data, n, m, 0, 99, 0.03, 0)
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<mem>", n, m, None
    src = (L.SRC_HOST, raw.ctypes.data, raw)
    same = lambda a, b: bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))  # noqa: E731
    ok = True
    for stat, bs in ((L.STAT_KING, 7000), (L.STAT_IBS_RAW, 7000), (L.STAT_AS, 7001), (L.STAT_GRM, 7001)):
        ref0, ref1 = tdist.pairwise_sharded(eng, X, None, None, stat, bs, mode="none", rank=0, world=1, src=src)
        out0, out1 = tdist.pairwise_sharded(eng, X, None, None, stat, bs, mode="loci", rank=rank, world=world, src=src)
        if stat == L.STAT_GRM:
            ok &= bool(np.allclose(out0, ref0, rtol=0, atol=1e-10, equal_nan=True))
        else:
            ok &= same(out0, ref0) and (out1 is None or same(out1, ref1))
    ref0, _ = tdist.pairwise_sharded(eng, X, None, None, L.STAT_KING, 7000, mode="none", rank=0, world=1, src=src)
    mine, _ = tdist.pairwise_sharded(eng, X, None, None, L.STAT_KING, 7000, mode="tiles", rank=rank, world=world, src=src)
    mask = tdist.owned_mask(n, rank, world)
    ok &= same(mine[mask], ref0[mask]) and bool(np.all(mine[~mask] == 0))
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK_OK" if int(flag.item()) == 1 else "DIST_CHECK_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
