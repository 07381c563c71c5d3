"""GPU bring-up harness: runs each kernel-level check in its own subprocess (a trapped
kernel kills the CUDA context, not the whole harness) and logs one JSON line per step.

    python scripts/gpu_bringup.py            # all steps
    python scripts/gpu_bringup.py STEP ...   # selected steps (inside the subprocess: --run STEP)
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _planes_ref(g):
    """g: uint8 [n, l] genotype bytes -> (d, nm, h1) int8 [n, l]."""
    nm = (g <= 2).astype(np.int8)
    h1 = (g == 1).astype(np.int8)
    d = (g == 2).astype(np.int8) - (g == 0).astype(np.int8)
    return d, nm, h1


def step_decode(k_major, gather):
    import torch
    from tidypopgen_b200 import _lib as L

    ctx = L.Context(0)
    rng = np.random.default_rng(1)
    nrow, ncol = 929, 700
    raw = rng.choice(np.array([0, 1, 2, 3, 4, 5, 6, 255, 128, 7], dtype=np.uint8), size=(ncol, nrow),
                     p=[.3, .3, .25, .05, .02, .02, .02, .01, .02, .01])
    if gather:
        rows = rng.permutation(nrow)[:613].astype(np.int32)
        cols = rng.permutation(ncol)[:391].astype(np.int32)
    else:
        rows = np.arange(5, 5 + 801, dtype=np.int32)
        cols = np.arange(3, 3 + 515, dtype=np.int32)
    n, lc = len(rows), len(cols)
    n_pad = (n + 255) // 256 * 256
    lc_pad = (lc + 127) // 128 * 128
    d_raw = torch.from_numpy(raw.reshape(-1)).cuda()
    d_rows = torch.from_numpy(rows).cuda()
    d_cols = torch.from_numpy(cols).cuda()
    plane_stride = n_pad * lc_pad
    d_planes = torch.full((3 * plane_stride,), 77, dtype=torch.int8, device="cuda")
    eb = L.errbuf()
    rc = L.lib().tpg_decode_planes(
        ctx.handle, d_raw.data_ptr(), raw.size, nrow,
        d_rows.data_ptr() if gather else None, int(rows[0]), n, n_pad,
        d_cols.data_ptr() if gather else None, int(cols[0]), lc, lc_pad,
        d_planes.data_ptr(), plane_stride, 1 if k_major else 0, eb, len(eb))
    L.check(rc, eb)
    torch.cuda.synchronize()
    got = d_planes.cpu().numpy().reshape(3, -1)
    g = raw[np.ix_(cols, rows)].T  # [n, lc]
    ref = np.zeros((3, n_pad, lc_pad), np.int8)
    for p, x in enumerate(_planes_ref(g)):
        ref[p, :n, :lc] = x
    if k_major:
        got = got.reshape(3, n_pad, lc_pad)
    else:
        got = got.reshape(3, lc_pad, n_pad).transpose(0, 2, 1)
    bad = int((got != ref).sum())
    return {"mismatches": bad, "n": n, "lc": lc, "ok": bad == 0}


def _gram_ref(planes, mask):
    from tidypopgen_b200 import _lib as L
    d, nm, h1 = [p.astype(np.float64) for p in planes]
    out = {}
    if mask & (1 << L.MAT_DD):
        out[L.MAT_DD] = d @ d.T
    if mask & (1 << L.MAT_NN):
        out[L.MAT_NN] = nm @ nm.T
    if mask & (1 << L.MAT_HH):
        out[L.MAT_HH] = h1 @ h1.T
    if mask & (1 << L.MAT_HN):
        out[L.MAT_HN] = h1 @ nm.T
    return {k: v.astype(np.int64) for k, v in out.items()}


def step_gram(k_major, n_pad, lc_pad, reps=1):
    import torch
    from tidypopgen_b200 import _lib as L

    ctx = L.Context(0)
    rng = np.random.default_rng(2)
    g = rng.choice(np.array([0, 1, 2, 3], dtype=np.uint8), size=(n_pad, lc_pad), p=[.4, .3, .25, .05])
    planes = _planes_ref(g)  # each [n_pad, lc_pad]
    host = np.stack([p if k_major else p.T for p in planes]).copy()
    d_planes = torch.from_numpy(host.reshape(-1)).cuda()
    mask = 0b1111
    d_counts = torch.zeros(4 * n_pad * n_pad, dtype=torch.int32, device="cuda")
    eb = L.errbuf()
    ms = C.c_float(0)
    times = []
    for r in range(reps):
        d_counts.zero_()
        torch.cuda.synchronize()
        rc = L.lib().tpg_gram(ctx.handle, d_planes.data_ptr(), n_pad * lc_pad, n_pad, lc_pad,
                              1 if k_major else 0, mask, 0, 1, d_counts.data_ptr(), C.byref(ms), eb, len(eb))
        L.check(rc, eb)
        times.append(ms.value)
    got = d_counts.cpu().numpy().reshape(4, n_pad, n_pad)  # [mat][j][i] (column-major)
    res = {"ms": times, "ok": True}
    if n_pad * n_pad * lc_pad <= 5e10:
        ref = _gram_ref(planes, mask)
        iu = np.triu_indices(n_pad)
        for mat, r in ref.items():
            G = got[mat].T  # G[i, j]
            if mat == L.MAT_HN:
                bad = int((G != r).sum())
            else:
                bad = int((G[iu] != r[iu]).sum())
            res[f"bad_mat{mat}"] = bad
            if bad:
                res["ok"] = False
                idx = np.argwhere(G != r)
                res[f"first_bad_mat{mat}"] = [[int(a), int(b), int(G[a, b]), int(r[a, b])] for a, b in idx[:6]]
    macs = 0
    tm, tn = n_pad // 128, n_pad // 256
    upper = sum(1 for ti in range(tm) for tj in range(tn) if 256 * tj + 255 >= 128 * ti)
    macs = (3 * upper + tm * tn) * 128 * 256 * lc_pad
    res["tensor_TOPS_issued"] = 2 * macs / (min(times) * 1e-3) / 1e12
    return res


def step_pairwise(stat, k_major):
    import torch  # noqa: F401  (initialises CUDA the same way the tests do)
    from tidypopgen_b200 import _lib as L

    rng = np.random.default_rng(3)
    nrow, ncol = 301, 2000
    raw = rng.choice(np.array([0, 1, 2, 3], dtype=np.uint8), size=(ncol, nrow), p=[.4, .3, .25, .05])
    n = nrow
    out0 = np.zeros((n, n), np.float64)
    out1 = np.zeros((n, n), np.float64)
    aux0 = np.zeros((n, n), np.float64)
    aux1 = np.zeros((n, n), np.float64)
    job = L.Job(abi_version=1, stat=stat, src_kind=L.SRC_HOST, device=0, bk_path=None,
                bk_data=raw.ctypes.data, nrow_total=nrow, ncol_total=ncol, ind_row=None, n=0,
                ind_col=None, m=0, block_size=0, pad_quirk=1, shard_mode=0, shard_rank=0, shard_world=1,
                gpu_chunk_loci=512, flags=L.FLAG_PLANES_K_MAJOR if k_major else 0)
    res = L.Result(out0=out0.ctypes.data, out1=out1.ctypes.data, aux0=aux0.ctypes.data,
                   aux1=aux1.ctypes.data, out_on_device=0)
    eb = L.errbuf()
    rc = L.lib().tpg_pairwise(C.byref(job), C.byref(res), eb, len(eb))
    L.check(rc, eb)
    g = raw.T
    d, nm, h1 = [p.astype(np.float64) for p in _planes_ref(g)]
    info = {"stats": res.stats.as_dict()}
    with np.errstate(all="ignore"):
        if stat == L.STAT_KING:
            h0 = (g == 0).astype(np.float64)
            h2 = (g == 2).astype(np.float64)
            K = h1 @ h1.T - 2 * (h0 @ h2.T + h2 @ h0.T)
            Ni = h1 @ nm.T
            Nj = Ni.T
            pm = np.minimum(Ni, Nj)
            king = K / (2 * pm) + 0.5 - 0.25 * (Ni + Nj) / pm
            info["K_bad"] = int((aux0.T != K).sum())      # column-major buffers: transpose
            info["N_bad"] = int((aux1.T != Ni).sum())
            info["king_bad"] = int((~((out0.T == king) | (np.isnan(out0.T) & np.isnan(king)))).sum())
            info["ok"] = info["K_bad"] == 0 and info["N_bad"] == 0 and info["king_bad"] == 0
        elif stat == L.STAT_IBS_RAW:
            h0 = (g == 0).astype(np.float64)
            h2 = (g == 2).astype(np.float64)
            ibs = 2 * (h2 @ h2.T + h1 @ h1.T + h0 @ h0.T) + h1 @ (h0 + h2).T + (h0 + h2) @ h1.T
            valid = 2 * nm @ nm.T
            info["ibs_bad"] = int((out0.T != ibs).sum())
            info["valid_bad"] = int((out1.T != valid).sum())
            info["ok"] = info["ibs_bad"] == 0 and info["valid_bad"] == 0
        elif stat == L.STAT_AS:
            num = d @ d.T
            den = nm @ nm.T
            as_ = 0.5 * (1 + num / den)
            info["as_bad"] = int((out0.T != as_).sum())
            info["ok"] = info["as_bad"] == 0
    return info


def step_bench(k_major):
    """HGDP-shape kernel timings: decode 929 x 65536 chunk + gram for KING products."""
    import torch
    from tidypopgen_b200 import _lib as L

    ctx = L.Context(0)
    nrow, lc = 929, 262144
    n_pad = 1024
    d_raw = torch.empty(nrow * lc + 16, dtype=torch.uint8, device="cuda")
    eb = L.errbuf()
    L.check(L.lib().tpg_synth_fill(ctx.handle, d_raw.data_ptr(), nrow, lc, 0, 20260002, 0.02, 0, eb, len(eb)), eb)
    plane_stride = n_pad * lc
    d_planes = torch.empty(3 * plane_stride, dtype=torch.int8, device="cuda")
    out = {}
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    ts = []
    for r in range(5):
        t0 = time.perf_counter()
        L.check(L.lib().tpg_decode_planes(ctx.handle, d_raw.data_ptr(), nrow * lc, nrow, None, 0, nrow, n_pad,
                                          None, 0, lc, lc, d_planes.data_ptr(), plane_stride,
                                          1 if k_major else 0, eb, len(eb)), eb)
        ts.append((time.perf_counter() - t0) * 1e3)
    out["decode_ms_wall"] = ts
    out["decode_GBps_alg"] = 4 * nrow * lc / (min(ts) * 1e-3) / 1e9
    d_counts = torch.zeros(4 * n_pad * n_pad, dtype=torch.int32, device="cuda")
    ms = C.c_float(0)
    for mask, name in ((0b1011, "king"), (0b0111, "ibs"), (0b0011, "as")):
        tt = []
        for r in range(4):
            d_counts.zero_()
            torch.cuda.synchronize()
            L.check(L.lib().tpg_gram(ctx.handle, d_planes.data_ptr(), plane_stride, n_pad, lc,
                                     1 if k_major else 0, mask, 0, 1, d_counts.data_ptr(), C.byref(ms), eb, len(eb)), eb)
            tt.append(ms.value)
        out[f"gram_{name}_ms"] = tt
        out[f"gram_{name}_pairSNP_per_s"] = nrow * nrow * lc / (min(tt) * 1e-3)
    out["ok"] = True
    return out


def step_bench_large(n, lc, mask=0b1011):
    """Large-n Gram timing (tile-order / L2 behaviour): n x lc chunk, KING products."""
    import torch
    from tidypopgen_b200 import _lib as L

    ctx = L.Context(0)
    n_pad = (n + 255) // 256 * 256
    d_raw = torch.empty(n * lc + 16, dtype=torch.uint8, device="cuda")
    eb = L.errbuf()
    L.check(L.lib().tpg_synth_fill(ctx.handle, d_raw.data_ptr(), n, lc, 0, 20260004, 0.01, 0, eb, len(eb)), eb)
    ps = n_pad * lc
    d_planes = torch.empty(3 * ps, dtype=torch.int8, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.ExternalStream(L.lib().tpg_ctx_stream(ctx.handle))
    out = {"n": n, "lc": lc}
    ts = []
    for r in range(3):
        e0.record(stream)
        L.check(L.lib().tpg_decode_planes(ctx.handle, d_raw.data_ptr(), n * lc, n, None, 0, n, n_pad, None, 0, lc, lc,
                                          d_planes.data_ptr(), ps, 0, eb, len(eb)), eb)
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out["decode_ms"] = ts
    out["decode_GBps_alg"] = 4 * n * lc / (min(ts) * 1e-3) / 1e9
    d_counts = torch.zeros(4 * n_pad * n_pad, dtype=torch.int32, device="cuda")
    ms = C.c_float(0)
    tt = []
    for r in range(3):
        L.check(L.lib().tpg_gram(ctx.handle, d_planes.data_ptr(), ps, n_pad, lc, 0, mask, 0, 1,
                                 d_counts.data_ptr(), C.byref(ms), eb, len(eb)), eb)
        tt.append(ms.value)
    out["gram_ms"] = tt
    out["alg_TOPS"] = 4.0 * n * n * lc / (min(tt) * 1e-3) / 1e12
    out["pairSNP_per_s"] = n * n * lc / (min(tt) * 1e-3)
    # spot check a few entries against numpy on the host
    host = d_raw[: n * lc].cpu().numpy().reshape(lc, n)
    rng = np.random.default_rng(0)
    got = d_counts.view(4, n_pad, n_pad)
    bad = 0
    for _ in range(40):
        i, j = sorted(int(x) for x in rng.integers(0, n, 2))
        gi, gj = host[:, i].astype(np.int64), host[:, j].astype(np.int64)
        di = (gi == 2).astype(np.int64) - (gi == 0)
        dj = (gj == 2).astype(np.int64) - (gj == 0)
        dd = int((di * dj).sum()) * 3
        hn = int(((gi == 1) & (gj <= 2)).sum()) * 3
        bad += int(got[0, j, i].item()) != dd
        bad += int(got[3, j, i].item()) != hn
    out["spot_bad"] = bad
    out["ok"] = bad == 0
    return out


STEPS = {
    "decode_mn": lambda: step_decode(False, False),
    "decode_mn_gather": lambda: step_decode(False, True),
    "decode_k": lambda: step_decode(True, False),
    "decode_k_gather": lambda: step_decode(True, True),
    "gram_k_small": lambda: step_gram(True, 256, 256),
    "gram_mn_small": lambda: step_gram(False, 256, 256),
    "gram_k_mid": lambda: step_gram(True, 1024, 4096, reps=3),
    "gram_mn_mid": lambda: step_gram(False, 1024, 4096, reps=3),
    "king_mn": lambda: step_pairwise(0, False),
    "king_k": lambda: step_pairwise(0, True),
    "ibs_mn": lambda: step_pairwise(1, False),
    "as_mn": lambda: step_pairwise(4, False),
    "bench_mn": lambda: step_bench(False),
    "bench_k": lambda: step_bench(True),
    "ab_hgdp": lambda: step_bench(False),
    "large_8k": lambda: step_bench_large(8192, 65536),
    "large_20k": lambda: step_bench_large(20000, 32768),
    "large_50k": lambda: step_bench_large(50000, 8192),
}


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        name = sys.argv[2]
        try:
            r = STEPS[name]()
        except Exception as e:  # noqa: BLE001
            r = {"ok": False, "error": f"{type(e).__name__}: {e}"}
        print("RESULT " + json.dumps({"step": name, **r}), flush=True)
        return
    names = sys.argv[1:] or list(STEPS)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bringup.log", "a") as log:
        for name in names:
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, __file__, "--run", name], capture_output=True, text=True,
                                   timeout=240)
                lines = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
                msg = lines[-1][7:] if lines else json.dumps(
                    {"step": name, "ok": False, "rc": p.returncode, "stderr": p.stderr[-1500:]})
            except subprocess.TimeoutExpired:
                msg = json.dumps({"step": name, "ok": False, "error": "timeout"})
            line = f"[{time.time() - t0:6.1f}s] {msg}"
            print(line, flush=True)
            log.write(line + "\n")
            log.flush()


if __name__ == "__main__":
    main()
