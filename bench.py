"""bench.py — pairwise_king pair-SNP ops/s (n^2 * m / s) on BASELINE.json's north-star workload:
configs[3], the biobank shape — 50,000 individuals x 500,000 SNPs, 1 % missing — on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step = one whole pairwise_king job on HBM-resident .bk bytes: zero the count matrices, decode_planes_k4,
gram_mxf4_pair (tcgen05 kind::mxf4 on exact 4-bit operands, chain-length guarded), finish_king -> the n x n
FP64 kinship matrix in HBM.  The same 25 GB fit one GPU, so every N runs the SAME job (STRONG scaling):
N > 1 shards the upper-triangle output tiles over the ranks (1024 x 1024 ownership regions, every rank
streams all loci, no data-path collective) — the partition that replaces the reference's serial block loop
(R/snp_king.R:63-77).  `value` is device-timed (CUDA events on the library's stream, max over ranks).

Beside it, in the same JSON line:
  e2e          the same job through the public API from a .bk FILE to a host matrix (tpg_pairwise /
               tpg_pairwise_multi over all N GPUs from one process, as an R session would call it)
  roofline     the Gram kernel against the MMA issue rate measured in this run (tpg_mma_peak);
               roofline_int8 = the int8 -> s32 route (TPG_FLAG_GRAM_I8) against its own measured peak
  parity       sampled pairs against the reference's per-pair arithmetic, both routes, first loci block
               against the CPU baseline's K and N
  cpu_baseline the reference's CPU path (restated, oracle/blas_baseline.py) on this host, N = 1 only
  extras       the other named shapes: C2 (929 x 650k KING; KING + IBS + AS in one pass), C3 (2504 x 1M GRM,
               loci-sharded + NCCL all-reduce), C5 (10k x 1M IBS; shuffled row subset + gathered loci)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_IND, M_LOCI, MISSING, SEED = 50_000, 500_000, 0.01, 20260004
METRIC = "pairwise_king pair-SNP ops/sec (n^2*m)"
UNIT = "pair-SNP/s"
KING_OPS_PER_PAIRSNP = 4.0   # 2.0 MAC: DD (upper) + NN (upper) + HN (full) — DESIGN.md "Algebra"


def workload_name(n, m):
    base = f"biobank-shape synthetic {n} x {m}, 1% missing, pairwise_king"
    if (n, m) == (N_IND, M_LOCI):
        return base + " (BASELINE.json configs[3], the north-star target)"
    return base + " (REDUCED size for development: not the bench workload)"


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
        p.update(hbm_gbs=float(j["hbm_gbs"]), bf16_tflops=float(j["bf16_tflops"]),
                 bf16_tflops_sustained=float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                 source="MEASURED_PEAKS.json")
    except Exception:  # noqa: BLE001
        pass
    return p


class ClockSampler(threading.Thread):
    """SM clocks + throttle reasons DURING the timed region (B200_PROFILING.md), read through NVML
    in-process (the library nvidia-smi itself uses): a `nvidia-smi` subprocess every few ms from
    each of 8 ranks serialises on the driver and stalls everybody's launches.  Falls back to the
    nvidia-smi query when pynvml is missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:  # noqa: BLE001
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:  # noqa: BLE001
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        try:
            pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        except Exception:  # noqa: BLE001
            pw = 0.0
        flag = lambda name: "Active" if r & getattr(n, name, 0) else "Not Active"  # noqa: E731
        self.rows.append([str(self.index), str(sm), str(mx), str(pw),
                          flag("nvmlClocksThrottleReasonHwSlowdown"), flag("nvmlClocksThrottleReasonHwThermalSlowdown"),
                          flag("nvmlClocksThrottleReasonSwThermalSlowdown"), flag("nvmlClocksThrottleReasonSwPowerCap")])

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                    for ln in out.strip().splitlines():
                        self.rows.append([x.strip() for x in ln.split(",")])
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05 if self.nvml is not None else 0.2)

    def summary(self):
        ok = [r for r in self.rows if len(r) >= 8]
        sm = [float(r[1]) for r in ok if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in ok if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in ok if r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({nm for r in ok for nm, v in zip(names, r[4:8]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm),
                "source": "NVML" if self.nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------- reference arm

def synth_numpy(n, m, seed, missing):
    """Same law as include/tpg_synth.h (p ~ U(0.05, 0.5), Binomial(2, p), iid missing -> 3),
    drawn with numpy so the reference arm touches none of this repo's native code."""
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.05, 0.5, size=(m, 1))
    g = (rng.random((m, n)) < p).astype(np.uint8) + (rng.random((m, n)) < p).astype(np.uint8)
    g[rng.random((m, n)) < missing] = 3
    return g


CPU_SAMPLE_INDIV = 12_500   # a quarter of the individuals = 1/16 of the pairs


def cpu_sample_desc(n_s, sample, n, m, bs, cores):
    return (f"loci block 1 of R's block loop ({sample} of {m} loci, block.size {bs} = bigstatsr::block_size({n})) on the first "
            f"{n_s} of {n} individuals; numpy decode into FP64 indicator planes + FP64 dgemm (MKL) on {cores} threads, "
            "K / N accumulators and finish as the reference (g1 g1' goes through dgemm where Armadillo emits dsyrk: 8 flop per "
            "pair-SNP against the reference's 7, i.e. up to ~12 % pessimistic for the CPU); the rate is per pair-SNP and FP64 GEMM at this size is "
            "compute bound, so it extrapolates to the full 187-block, 50,000-individual job (which needs 60 GB of FP64 "
            "temporaries and ~2 h of these cores)")


def run_reference(args, rank):
    """The reference's CPU path (restated: oracle/blas_baseline.py) on the host cores."""
    if rank != 0:
        return
    import torch
    from oracle import blas_baseline as B
    from oracle import oracle as O

    # every host thread the process may use (torchrun exports OMP_NUM_THREADS=1 for its ranks)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:  # noqa: BLE001
        pass
    n, m = args.indiv, args.loci
    bs = O.block_size(n)                       # pairwise_king passes block_size(nrow(x)): 2,684 at n = 50,000
    sample = int(O.split_len(m, bs)[0, 2])     # loci of R's first block
    n_s = min(n, CPU_SAMPLE_INDIV)
    raw = synth_numpy(n_s, sample, SEED, MISSING)
    t0 = time.perf_counter()
    B.king_blocks(raw, n_s, sample, bs, max_blocks=1)
    first = time.perf_counter() - t0
    if first > 6.0:                            # keep the whole run within minutes on a small host
        n_s = max(1024, int(n_s * (4.0 / first) ** 0.5) // 64 * 64)
        raw = np.ascontiguousarray(raw[:, :n_s])
    for _ in range(max(0, args.warmup - 1)):
        B.king_blocks(raw, n_s, sample, bs, max_blocks=1)
    times = []
    for _ in range(args.steps):
        K, N, done, sec = B.king_blocks(raw, n_s, sample, bs, max_blocks=1)
        t1 = time.perf_counter()
        B.king_finish(K, N)
        times.append(sec + (time.perf_counter() - t1) * (sample / m))
    sec = float(np.mean(times))
    value = n_s * n_s * sample / sec
    cores = torch.get_num_threads()
    desc = cpu_sample_desc(n_s, sample, n, m, bs, cores)
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, m), "sample": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ---------------------------------------------------------------------------- B200 arm

def king_pair(gi, gj):
    """KING-robust for one pair from raw bytes, the reference's arithmetic
    (src/snp_king.cpp:45-72, R/snp_king.R:82-91)."""
    h0i, h1i, h2i, vi = gi == 0, gi == 1, gi == 2, gi <= 2
    h0j, h1j, h2j, vj = gj == 0, gj == 1, gj == 2, gj <= 2
    K = float((h1i & h1j).sum()) - 2.0 * float((h0i & h2j).sum() + (h2i & h0j).sum())
    Ni, Nj = float((h1i & vj).sum()), float((vi & h1j).sum())
    pm = min(Ni, Nj)
    with np.errstate(all="ignore"):
        return np.float64(K) / np.float64(2 * pm) + 0.5 - np.float64(0.25 * (Ni + Nj)) / np.float64(pm)


def same_f64(a, b):
    return bool(a == b or (np.isnan(a) and np.isnan(b)))


def fake_fbm(tpg, n, m, name="<synthetic>"):
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = name, n, m, None
    return X


def measure_mma_peaks(L, local):
    """tcgen05 issue rate of this GPU for both Gram routes (operands resident in shared memory, every
    CTA pair of the chip): 0.05 s = burst, 1 s = sustained under the power cap."""
    out = {}
    eb = L.errbuf()
    for kind, name in ((1, "mxf4"), (0, "i8")):
        for secs, tag, k in ((0.05, "burst", kind), (1.0, "sustained", kind), (0.05, "burst_zero_operands", kind + 2)):
            tops, ms = C.c_double(0), C.c_double(0)
            rc = L.lib().tpg_mma_peak(local, k, secs, C.byref(tops), C.byref(ms), eb, len(eb))
            out[f"{name}_{tag}_tops"] = tops.value if rc == 0 else None
            if rc != 0:
                out[f"{name}_{tag}_error"] = eb.value.decode(errors="replace")
    out["how"] = ("tpg_mma_peak: tcgen05.mma.cta_group::2 M256 N256 (K32 int8 / K64 E2M1) issued back to back on all 74 CTA "
                  "pairs, pseudo-random -1/0/+1 operands resident in shared memory; 2 ops per MAC; burst = 0.05 s, sustained = 1 s "
                  "(power-capped)")
    return out


def write_bk_file(path, dev_raw, nbytes):
    """Streams the synthetic .bk image from HBM to a file (pinned bounce buffer, 256 MB pieces)."""
    import torch
    piece = 256 << 20
    bounce = torch.empty(piece, dtype=torch.uint8).pin_memory()
    t0 = time.perf_counter()
    with open(path, "wb", buffering=0) as f:
        for off in range(0, nbytes, piece):
            k = min(piece, nbytes - off)
            bounce[:k].copy_(dev_raw[off:off + k])
            f.write(memoryview(bounce.numpy())[:k])
        os.fsync(f.fileno())
    return time.perf_counter() - t0


def drop_page_cache(path):
    fd = os.open(path, os.O_RDONLY)
    try:
        os.fsync(fd)
        os.posix_fadvise(fd, 0, 0, os.POSIX_FADV_DONTNEED)
    finally:
        os.close(fd)


def run_b200(args, rank, world, local):
    import torch
    import torch.distributed as dist

    import tidypopgen_b200 as tpg
    from tidypopgen_b200 import _lib as L
    from tidypopgen_b200 import dist as tdist

    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    lib = L.lib()
    eb = L.errbuf()
    n, m = args.indiv, args.loci
    gloo = dist.new_group(backend="gloo") if world > 1 else None   # host-side waits that keep the GPUs free

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=gloo)

    # ---- measured tensor peaks of this GPU (all ranks at once: the box's power budget is shared)
    barrier()
    mma = measure_mma_peaks(L, local)
    lib.tpg_release(local)
    eng = tdist.CudaEngine(local)

    # ---- synthetic .bk image, resident in HBM (every rank holds all loci: tile sharding streams them all)
    dev_raw = torch.empty(n * m + 64, dtype=torch.uint8, device=dev)
    L.check(lib.tpg_synth_fill(eng.ctx.handle, dev_raw.data_ptr(), n, m, 0, SEED, MISSING, 0, eb, len(eb)), eb)
    X = fake_fbm(tpg, n, m)
    bs = tpg.block_size(n)   # pairwise_king's default block.size (R/pairwise_king.R:36); KING does not depend on it
    src = (L.SRC_DEVICE, dev_raw.data_ptr(), dev_raw)
    tiles = world > 1
    shard = dict(shard_mode=L.SHARD_TILES if tiles else L.SHARD_NONE, shard_rank=rank, shard_world=world)
    job, keep = tpg.snp.make_job(X, None, None, L.STAT_KING, bs, device=local, src=src, **shard)
    job8, keep8 = tpg.snp.make_job(X, None, None, L.STAT_KING, bs, device=local, src=src, flags=L.FLAG_GRAM_I8, **shard)
    out_dev = torch.zeros((n, n), dtype=torch.float64, device=dev)   # column-major storage of KING
    res = L.Result(out0=out_dev.data_ptr(), out1=None, aux0=None, aux1=None, out_on_device=1)

    def run_steps(the_job, steps, warmup):
        agg = {"ms_decode": 0.0, "ms_gram": 0.0, "ms_finish": 0.0, "launches": 0, "chunks": 0, "macs": 0, "q4": 0,
               "i8": 0, "chain": 0, "wide": 0}
        for _ in range(warmup):
            L.check(lib.tpg_pairwise_ctx(eng.ctx.handle, C.byref(the_job), C.byref(res), eb, len(eb)), eb)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(eng.stream)
        for _ in range(steps):
            # the public one-call-per-job entry: zero counts, decode, gram, finish, one sync
            L.check(lib.tpg_pairwise_ctx(eng.ctx.handle, C.byref(the_job), C.byref(res), eb, len(eb)), eb)
            st = res.stats
            agg["ms_decode"] += st.ms_decode
            agg["ms_gram"] += st.ms_gram
            agg["ms_finish"] += st.ms_finish
            agg["launches"] += st.n_launches
            agg["chunks"] += st.n_chunks
            agg["macs"] += st.mma_macs
            agg["q4"] += st.n_mxf4_chunks
            agg["i8"] += st.n_i8_chunks
            agg["wide"] += st.n_wide_launches
            agg["chain"] = max(agg["chain"], st.max_chain_loci)
        e1.record(eng.stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, agg

    sampler = ClockSampler(local)
    sampler.start()
    ms_step, agg = run_steps(job, args.steps, args.warmup)
    sampler.stop_flag = True
    sampler.join(2)
    value = n * n * m / (ms_step * 1e-3)
    if os.environ.get("TPG_BENCH_DEBUG"):
        print(f"[rank {rank}] {ms_step:.3f} ms/step, gram {agg['ms_gram'] / args.steps:.2f} ms", file=sys.stderr, flush=True)

    # ---- parity (1): sampled owned pairs against the reference's per-pair arithmetic, exact
    rng = np.random.default_rng(1000 + rank)
    nb = (n + 255) // 256
    cols_view = dev_raw[: n * m].view(m, n)
    pairs, bad = [], 0
    while len(pairs) < args.pairs:
        i, j = (int(x) for x in rng.integers(0, n, 2))
        if len(pairs) % 8 == 0:
            j = i
        bi, bj = sorted((i // 256, j // 256))
        if tiles and tdist.block_owner(bi, bj, nb, world) != rank:
            continue
        pairs.append((i, j))
    for i, j in pairs:
        ref = king_pair(cols_view[:, i].cpu().numpy(), cols_view[:, j].cpu().numpy())
        got, got_t = out_dev[j, i].item(), out_dev[i, j].item()   # element (i, j) lives at [j, i]
        bad += 0 if (same_f64(got, ref) and same_f64(got_t, got)) else 1
    # a fingerprint of this route's result on owned entries, to compare the int8 route and the e2e run with
    fp_idx = [(int(a), int(b)) for a, b in rng.integers(0, n, (4096, 2))]
    if tiles:
        fp_idx = [(i, j) for i, j in fp_idx if tdist.block_owner(*sorted((i // 256, j // 256)), nb, world) == rank]
    fi = torch.tensor([j * n + i for i, j in fp_idx], dtype=torch.int64, device=dev)
    fp_q4 = out_dev.view(-1)[fi].cpu().numpy()

    # ---- the int8 -> s32 route on the same job (north_star's literal kernel), for its roofline figure
    ms_step8, agg8 = run_steps(job8, max(1, min(2, args.steps)), 1)
    fp_i8 = out_dev.view(-1)[fi].cpu().numpy()
    routes_agree = bool(np.array_equal(fp_q4, fp_i8, equal_nan=True))
    flags = torch.tensor([bad, 0 if routes_agree else 1], device=dev)
    if world > 1:
        dist.all_reduce(flags)
    bad_total, routes_bad = int(flags[0].item()), int(flags[1].item())

    # ---- parity (2) + CPU baseline (N = 1): loci block 1 of R's loop on the first individuals, K and N exact
    cpu, parity_first_block = None, None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import blas_baseline as B
        from oracle import oracle as O
        try:
            torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
        except Exception:  # noqa: BLE001
            pass
        sample = int(O.split_len(m, bs)[0, 2])
        n_s = min(n, CPU_SAMPLE_INDIV)
        raw_np = cols_view[:sample, :n_s].contiguous().cpu().numpy()
        B.king_blocks(raw_np[:256], n_s, 256, bs, max_blocks=1)  # BLAS thread warm-up
        Kc, Nc, done, sec = B.king_blocks(raw_np, n_s, sample, bs, max_blocks=1)
        cores = torch.get_num_threads()
        cpu = {"value": n_s * n_s * done / sec, "unit": UNIT, "cores": cores, "kind": "port", "seconds": sec,
               "sample": cpu_sample_desc(n_s, sample, n, m, bs, cores)}
        rows1 = np.arange(1, n_s + 1, dtype=np.int32)
        cols1 = np.arange(1, sample + 1, dtype=np.int32)
        _, Kg, Ng = tpg.snp_king_counts(X, rows1, cols1, block_size=bs, src=src, device=local)
        parity_first_block = bool(np.array_equal(Kg, Kc) and np.array_equal(Ng, Nc))
        del Kg, Ng, Kc, Nc
        lib.tpg_release(local)

    # ---- e2e: .bk FILE -> public API -> host matrix, one process driving all N GPUs (rank 0)
    e2e, streaming = None, None
    if not args.no_e2e:
        path = os.path.join(args.tmpdir, f"tpg_bench_{n}x{m}_{os.getpid()}.bk")
        t_write = write_bk_file(path, dev_raw, n * m) if rank == 0 else 0.0
    samp_cols = None
    if not args.no_e2e and rank == 0:   # genotype columns of a few sampled pairs, kept for the e2e check
        samp_cols = {k: cols_view[:, k].cpu().numpy() for ij in pairs[:16] for k in ij}
    del cols_view, dev_raw, out_dev, src, keep, keep8, job, job8
    lib.tpg_ctx_trim(eng.ctx.handle)
    torch.cuda.empty_cache()
    host_barrier()
    if not args.no_e2e and rank == 0:
        try:
            Xf = tpg.FBMCode256(path, n, m)
            devices = list(range(world))
            # one process drives all N GPUs here, as an R session does; torchrun's LOCAL_WORLD_SIZE would otherwise
            # make this rank's staging pool take only 1/N of the host cores (the other ranks are idle in a host barrier)
            L.config(host_threads=min(32, len(os.sched_getaffinity(0))))
            king_host = np.zeros((n, n), np.float64, order="F")

            def e2e_call():
                t0 = time.perf_counter()
                tpg.snp.run_job(Xf, None, None, L.STAT_KING, bs, out0=king_host, devices=devices, device=devices[0])
                return time.perf_counter() - t0, dict(tpg.snp.last_stats)

            e2e_call()                                   # warm-up: buffers, page cache
            runs = [e2e_call() for _ in range(args.e2e_steps)]
            # (the median: one call in a run has been seen to spend 1.15 s instead of 0.40 s in the hand-back on a busy host)
            order = sorted(range(len(runs)), key=lambda k: runs[k][0])
            sec, st = runs[order[len(runs) // 2]]
            ok_pairs = all(same_f64(king_host[i, j], king_pair(samp_cols[i], samp_cols[j])) and
                           same_f64(king_host[j, i], king_host[i, j]) for i, j in pairs[:16])
            ok_fp = bool(np.array_equal(np.array([king_host[i, j] for i, j in fp_idx]), fp_q4, equal_nan=True))
            e2e = {"value": n * n * m / sec, "unit": UNIT, "ms_per_step": sec * 1e3, "steps": args.e2e_steps,
                   "runs_ms": [round(r[0] * 1e3, 1) for r in runs], "statistic": "median of the timed calls",
                   "h2d_bytes_per_step": int(st["bytes_h2d"]), "d2h_bytes_per_step": 8 * n * n, "host_bytes_per_step": n * m,
                   "source": "FBM .bk file (TPG_SRC_FILE, page cache warm) -> pread + 2-bit pack by the host worker pool -> "
                             "pinned staging -> H2D; result through pinned bounce buffers into pageable host memory",
                   "call": "tpg_pairwise" if world == 1 else f"tpg_pairwise_multi over {world} GPUs from one process "
                           "(loci-sharded, int32 counts summed over NVLink by peer-reading add kernels, finished once)",
                   "stage_ms": {k: st[k] for k in ("ms_read", "ms_h2d", "ms_decode", "ms_gram", "ms_reduce", "ms_finish", "ms_d2h")},
                   "chunks": int(st["n_chunks"]), "packed_chunks": int(st["n_packed_chunks"]),
                   "result_shells": int(st["n_result_shells"]),
                   "result_shells_note": "tpg_config.result_shells (default): the matrix is computed and handed back in nested shells, "
                                         "the D2H of one beside the products of the next; stage_ms.ms_d2h is then the host time of the "
                                         "hand-back loops, most of it overlapped" if st["n_result_shells"] else None,
                   "host_threads": L.config()["host_threads"],
                   "sampled_pairs_exact": bool(ok_pairs), "matches_device_run": ok_fp, "bk_write_s": t_write}
            # disk -> HBM streaming, reported separately from the kernels (north_star): the same call on a COLD
            # file (page cache dropped), read with O_DIRECT and through the page cache
            streaming = {"file_GB": n * m / 1e9, "warm_page_cache_ms": sec * 1e3}
            for mode, key in ((1, "cold_o_direct"), (0, "cold_buffered")):
                drop_page_cache(path)
                L.config(io_mode=mode)
                secs, stc = e2e_call()
                streaming[key + "_ms"] = secs * 1e3
                streaming[key + "_GBps_file_to_result"] = n * m / secs / 1e9
                streaming[key + "_read_stage_ms"] = stc["ms_read"]
                if mode == 1:
                    streaming["o_direct_chunks"] = int(stc["n_direct_io"])
            L.config(reset=True)
            del king_host
        finally:
            try:
                os.unlink(path)
            except OSError:
                pass
            for d in range(world):
                lib.tpg_release(d)
    host_barrier()

    # ---- the other named shapes (extras), at this N
    extras = None
    if not args.no_extras:
        extras = run_extras(args, rank, world, local, eng, gloo)

    if rank != 0:
        return
    pk = peaks()
    steps = args.steps

    def roof(a, ms_total, nsteps, route):
        ops = KING_OPS_PER_PAIRSNP * n * n * m / world            # algorithmic ops of ONE rank's share per step
        launches = max(1, a["chunks"])
        t_launch = ms_total / launches                            # avg launch duration (ms)
        achieved = ops * nsteps / (ms_total * 1e-3) / 1e12
        issued = 2 * a["macs"] / (ms_total * 1e-3) / 1e12
        tag = "mxf4" if route == "q4" else "i8"
        peak = mma.get(f"{tag}_sustained_tops") or ((4.0 if route == "q4" else 2.0) * pk.get("bf16_tflops_sustained", pk["bf16_tflops"]))
        wide = a["wide"] > 0   # tpg_stats.n_wide_launches: the wide-tile kernels (TMA-reduction epilogue) ran
        if route == "q4":
            kern = (("gram_mxf4_wide_kernel: 256 x 480 tiles, cp.reduce.async.bulk.tensor epilogue" if wide else
                     "gram_mxf4_pair_kernel: 256 x 240 tiles, ping-pong accumulators") +
                    " (tcgen05.mma.cta_group::2.kind::mxf4.block_scale, E2M1 x E2M1 -> f32, exact integer sums)")
        else:
            kern = (("gram_i8_wide_kernel: 256 x 512 tiles, cp.reduce.async.bulk.tensor epilogue" if wide else
                     "gram_i8_pair_kernel: 256 x 256 tiles, ping-pong accumulators") +
                    " (tcgen05.mma.cta_group::2.kind::i8, int8 x int8 -> s32)")
        return {"kernel": kern, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic_from_profile().get("gram_mxf4_c4_dram_bytes_per_launch" if route == "q4" else "gram_i8_c4_dram_bytes_per_launch"),
                "per_gpu": True, "avg_launch_ms": t_launch, "launches_per_step": launches // max(1, nsteps),
                "peak_source": f"measured in this run: {mma['how']}; sustained (1 s) figure because a step is a long, power-capped run "
                               f"(burst {mma.get(tag + '_burst_tops')})" if mma.get(f"{tag}_sustained_tops") else
                               f"{4 if route == 'q4' else 2} x bf16 sustained of {pk['source']} (probe failed)",
                "frac_of_nominal": achieved / (9000.0 if route == "q4" else 4500.0),
                "frac_of_bf16_multiple": achieved / ((4.0 if route == "q4" else 2.0) * pk["bf16_tflops"]),
                "issued_incl_tile_padding": issued,
                "note": "algorithmic = 4 tensor ops per pair-SNP (DD + NN upper triangles, HN full), this rank's share of the job"}

    q4 = agg["q4"] > 0
    roofline = roof(agg, agg["ms_gram"], steps, "q4" if q4 else "i8")
    roofline_int8 = roof(agg8, agg8["ms_gram"], max(1, min(2, steps)), "i8")
    roofline_int8["ms_per_step"] = ms_step8
    roofline_int8["value"] = n * n * m / (ms_step8 * 1e-3)
    roofline_int8["frac_of_nominal_int8_peak"] = roofline_int8["achieved"] / 4500.0
    dec_ms = agg["ms_decode"] / max(1, agg["chunks"])
    bytes_per_gt = 2.5 if q4 else 4.0
    dec_bytes = bytes_per_gt * n * (m * steps / max(1, agg["chunks"]))
    roof_dec = {"kernel": "decode_planes_k4_kernel" if q4 else "decode_planes_mn_kernel", "bound": "hbm",
                "achieved": dec_bytes / (dec_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": dec_bytes / (dec_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                "traffic": traffic_from_profile().get("decode_planes_k4_c4_dram_bytes_per_launch"),
                "note": (f"1 B read + {'1.5 (three 4-bit planes)' if q4 else '3'} B written per genotype; avg launch {dec_ms:.3f} ms; "
                         f"peak from {pk['source']}; chunked jobs run decode(c+1) beside gram(c), so the event interval of a "
                         "decode launch includes waiting for SMs the persistent Gram kernel holds — the C2 extra shows the kernel alone")}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "e2m1 x e2m1 -> f32 (exact integer sums, chains <= 2^23 loci)" if q4 else "int8 x int8 -> s32",
        "data": "synthetic",
        "config": {"workload": workload_name(n, m), "n": n, "m": m, "statistic": "pairwise_king",
                   "sharding": "none" if world == 1 else f"upper-triangle output tiles x{world} (1024 x 1024 ownership regions, every rank "
                                                         "streams all loci, no data-path collective)",
                   "l2": "inputs larger than L2 (25 GB raw + 37.5 GB of 4-bit planes per step vs 126 MB), no flush needed",
                   "tensor_ops_per_pair_snp": KING_OPS_PER_PAIRSNP},
        "clocks": sampler.summary(),
        "e2e": e2e,
        "gpu_launches": agg["launches"],
        "roofline": roofline, "roofline_int8": roofline_int8, "roofline_decode": roof_dec,
        "stage_ms_per_step": {"decode": agg["ms_decode"] / steps, "gram": agg["ms_gram"] / steps, "finish": agg["ms_finish"] / steps},
        "mma_peaks_measured": mma,
        "parity": {"sampled_pairs_per_rank": args.pairs, "sampled_pairs_mismatches": bad_total, "sampled_pairs_exact": bad_total == 0,
                   "int8_and_4bit_routes_agree": routes_bad == 0, "parity_first_block": parity_first_block,
                   "max_chain_loci_4bit": int(agg["chain"]), "chain_limit": L.MXF4_MAX_CHAIN_LOCI,
                   "how": "KING of sampled (owned) pairs recomputed from the raw bytes with the reference's per-pair arithmetic "
                          "(src/snp_king.cpp:45-72, R/snp_king.R:82-91), bit-identical FP64; 4096 sampled entries per rank equal between "
                          "the kind::mxf4 and kind::i8 routes; parity_first_block = K and N_Aa_i of loci block 1 on the CPU-baseline "
                          "sample equal to the CPU port's (N = 1)"},
        "streaming": streaming,
        "cpu_baseline": cpu,
        "extras": extras,
    }
    emit(line)


def traffic_from_profile():
    """dram bytes per launch from the committed ncu captures, if present (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return {}


# ---------------------------------------------------------------------------- extras: the other named shapes

def c2_file_jobs(args, tpg, L, lib, eb, eng, dev, local):
    """configs[1] (929 x 650,000) from a .bk FILE to host matrices, wall clock per call: streamed (pread + 2-bit pack
    + H2D, as every first job runs), and with tpg_config.cache_genotypes — the first whole-file job leaves the packed
    image in HBM (155 MB) and the jobs an R session runs next (another statistic, a QC'd subset of individuals, a
    thinned set of loci) gather from it on the device."""
    import torch
    n, m = 929, 650_000
    raw = torch.empty(n * m + 64, dtype=torch.uint8, device=dev)
    L.check(lib.tpg_synth_fill(eng.ctx.handle, raw.data_ptr(), n, m, 0, 20260002, 0.02, 0, eb, len(eb)), eb)
    path = os.path.join(args.tmpdir, f"tpg_bench_c2_{os.getpid()}.bk")
    write_bk_file(path, raw, n * m)
    del raw
    rng = np.random.default_rng(2)
    rows = (np.sort(rng.permutation(n)[:800]) + 1).astype(np.int32)          # 800 of the 929 individuals kept by a QC step
    cols = (np.arange(0, m, 5) + 1).astype(np.int32)                          # every 5th locus
    out = {}
    try:
        X = tpg.FBMCode256(path, n, m)

        def wall(f, reps=5):
            f()
            best = 1e30
            for _ in range(reps):
                t0 = time.perf_counter()
                f()
                best = min(best, time.perf_counter() - t0)
            return best * 1e3, dict(tpg.snp.last_stats)

        jobs = {"king_all": lambda: tpg.snp_king(X, device=local),
                "ibs_all": lambda: tpg.snp_ibs(X, device=local),
                "king_800_rows": lambda: tpg.snp_king(X, rows, device=local),
                "king_every_5th_locus": lambda: tpg.snp_king(X, None, cols, device=local)}
        L.config(cache_genotypes=0)
        ref = {k: f() for k, f in jobs.items()}
        for k, f in jobs.items():
            ms, st = wall(f)
            out[k] = {"streamed_ms": ms, "h2d_bytes": st["bytes_h2d"]}
        L.config(cache_genotypes=1)
        jobs["king_all"]()                                                    # fills the cache while it streams
        for k, f in jobs.items():
            same = bool(np.array_equal(f(), ref[k], equal_nan=True))
            ms, st = wall(f)
            out[k].update(from_genotype_cache_ms=ms, cached=int(st["n_geno_cached"]), h2d_bytes_cached=st["bytes_h2d"], same_result=same)
        out["note"] = "wall clock per call, host matrix included; best of 5"
    finally:
        L.config(cache_genotypes=0)
        try:
            os.unlink(path)
        except OSError:
            pass
        lib.tpg_ctx_trim(eng.ctx.handle)
        lib.tpg_release(local)
    return out


def run_extras(args, rank, world, local, eng, gloo):
    """C2 / C3 / C5 of BASELINE.json on the same GPUs.  N = 1: one tpg_pairwise_ctx call per step.  N > 1:
    loci-sharded STRONG scaling — every rank decodes and multiplies its slice of the loci, ONE NCCL all-reduce
    of the int32 count matrices the statistic needs, finish."""
    import torch
    import torch.distributed as dist

    import tidypopgen_b200 as tpg
    from tidypopgen_b200 import _lib as L
    from tidypopgen_b200 import dist as tdist

    lib, eb, dev = L.lib(), L.errbuf(), f"cuda:{local}"
    out = {}

    def timed(name, n_tot, m_tot, seed, missing, mode, stat, ops, rows=None, cols=None, steps=5, many=None, note=None):
        raw = torch.empty(n_tot * m_tot + 64, dtype=torch.uint8, device=dev)
        L.check(lib.tpg_synth_fill(eng.ctx.handle, raw.data_ptr(), n_tot, m_tot, 0, seed, missing, mode, eb, len(eb)), eb)
        X = fake_fbm(tpg, n_tot, m_tot)
        src = (L.SRC_DEVICE, raw.data_ptr(), raw)
        n = n_tot if rows is None else len(rows)
        m = m_tot if cols is None else len(cols)
        bs = tpg.block_size(n_tot)
        sharded = world > 1 and many is None
        job, keep = tpg.snp.make_job(X, rows, cols, stat, bs, device=local, src=src,
                                     flags=L.FLAG_ASYNC if sharded else 0,
                                     shard_mode=L.SHARD_LOCI if sharded else L.SHARD_NONE, shard_rank=rank, shard_world=world)
        outs = [torch.zeros((n, n), dtype=torch.float64, device=dev) for _ in range(len(many) if many else 1)]
        res = (L.Result * len(outs))()
        for k, o in enumerate(outs):
            res[k].out0, res[k].out_on_device = o.data_ptr(), 1
        counts = eng.new_counts(n) if sharded else None
        st_codes = np.ascontiguousarray(many if many else [stat], dtype=np.int32)
        agg = {"ms_decode": 0.0, "ms_gram": 0.0, "chunks": 0}

        def step():
            if many:
                L.check(lib.tpg_pairwise_many(C.byref(job), st_codes.ctypes.data, len(many), res, None, 0, L.SHARD_NONE, eb, len(eb)), eb)
                s = res[0].stats
            elif not sharded:
                L.check(lib.tpg_pairwise_ctx(eng.ctx.handle, C.byref(job), C.byref(res[0]), eb, len(eb)), eb)
                s = res[0].stats
            else:
                with torch.cuda.stream(eng.stream):
                    counts.zero_()
                s0 = L.Stats()
                res[0].stats = L.Stats()
                L.check(lib.tpg_accumulate(eng.ctx.handle, C.byref(job), counts.data_ptr(), C.byref(s0), eb, len(eb)), eb)
                with torch.cuda.stream(eng.stream):
                    tdist.all_reduce_counts(counts, stat, n)
                fin = L.Job.from_buffer_copy(job)
                fin.shard_mode, fin.shard_world, fin.shard_rank = L.SHARD_NONE, 1, 0
                L.check(lib.tpg_finish(eng.ctx.handle, C.byref(fin), counts.data_ptr(), C.byref(res[0]), eb, len(eb)), eb)
                s = res[0].stats
                s.n_chunks = s0.n_chunks
            agg["ms_decode"] += s.ms_decode
            agg["ms_gram"] += s.ms_gram
            agg["chunks"] += s.n_chunks

        for _ in range(2):
            step()
        agg.update(ms_decode=0.0, ms_gram=0.0, chunks=0)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        stream = eng.stream if not many else torch.cuda.current_stream()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3 / steps
        ms = e0.elapsed_time(e1) / steps if not many else wall   # (tpg_pairwise_many runs on the cached context's own stream)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if rank == 0:
            nplanes = 3 if ops != 2 else 2
            d = {"n": n, "m": m, "n_total": n_tot, "m_total": m_tot, "ms_per_step": ms, "value": n * n * m / (ms * 1e-3), "unit": UNIT,
                 "sharding": "none" if not sharded else f"loci x{world} (strong), one NCCL all-reduce of the needed int32 count matrices",
                 "decode_ms_per_step": agg["ms_decode"] / steps, "gram_ms_per_step": agg["ms_gram"] / steps,
                 "gram_alg_TOPS_per_gpu": ops * n * n * m / world / (agg["ms_gram"] / steps * 1e-3) / 1e12 if agg["ms_gram"] else None,
                 # 1 B read per staged genotype (a gathered job stages whole column spans) + half a byte per plane written
                 "decode_GBps_per_gpu": (1.0 * (n_tot if rows is not None else n) + 0.5 * nplanes * n) * m / world
                                        / (agg["ms_decode"] / steps * 1e-3) / 1e9 if agg["ms_decode"] else None}
            if note:
                d["note"] = note
            out[name] = d
        del raw, outs, counts
        lib.tpg_ctx_trim(eng.ctx.handle)
        lib.tpg_release(local)
        torch.cuda.empty_cache()

    # C2: HGDP shape (configs[1])
    timed("C2_king_929x650k", 929, 650_000, 20260002, 0.02, 0, L.STAT_KING, 4, steps=10)
    if world == 1:
        timed("C2_king_ibs_as_one_pass", 929, 650_000, 20260002, 0.02, 0, L.STAT_KING, 5, steps=10,
              many=[L.STAT_KING, L.STAT_IBS_PROPORTION, L.STAT_AS],
              note="tpg_pairwise_many: KING + IBS + allele sharing of configs[1] from ONE decode and four products (2.5 MAC per pair-SNP); "
                   "wall clock per call; value counts the pair-SNPs once")
        timed("C2_ibs_929x650k", 929, 650_000, 20260002, 0.02, 0, L.STAT_IBS_PROPORTION, 3, steps=10)
        timed("C2_allele_sharing_929x650k", 929, 650_000, 20260002, 0.02, 0, L.STAT_AS, 2, steps=10)
    if world == 1 and not args.no_e2e:
        out["C2_file_jobs"] = c2_file_jobs(args, tpg, L, lib, eb, eng, dev, local)
    # C3: 1000G shape, imputed bytes, GRM (configs[2])
    timed("C3_grm_2504x1M_imputed", 2504, 1_000_000, 20260003, 0.01, 1, L.STAT_GRM, 2, steps=5)
    # C5: mixed ploidy, 10 % missing, row / column subsets (configs[4])
    rng = np.random.default_rng(5)
    n5, m5 = 10_000, 1_000_000
    rows = (rng.permutation(n5)[: n5 * 7 // 10] + 1).astype(np.int32)
    tail = rng.permutation(m5)[: m5 // 20]
    cols = (np.concatenate([np.arange(0, m5, 3), tail]) + 1).astype(np.int32)
    timed("C5_ibs_7000_shuffled_rows_x_383k_gathered_loci", n5, m5, 20260005, 0.10, 2, L.STAT_IBS_PROPORTION, 3, rows=rows, cols=cols, steps=3,
          note="decode_ms includes gather_pack_kernel (row span staged in shared memory, gathered there, written 2-bit packed)")
    timed("C5_ibs_10000x1M_all", n5, m5, 20260005, 0.10, 2, L.STAT_IBS_PROPORTION, 3, steps=3)
    return out if rank == 0 else None


_REAL_STDOUT = None


def quiet_stdout():
    """Everything incidental (NCCL's version banner, library chatter) goes to stderr; the one
    JSON line is written to the real stdout by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    line = json.dumps(obj) + "\n"
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line.encode())
    else:
        sys.stdout.write(line)
        sys.stdout.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--indiv", type=int, default=N_IND, help="development only: a reduced shape is labelled as such")
    ap.add_argument("--loci", type=int, default=M_LOCI)
    ap.add_argument("--pairs", type=int, default=64, help="sampled pairs per rank for the exact parity check")
    ap.add_argument("--e2e-steps", type=int, default=3, help="timed end-to-end calls; the median is reported")
    ap.add_argument("--tmpdir", default=os.environ.get("TPG_BENCH_TMPDIR", tempfile.gettempdir()))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        from tidypopgen_b200 import dist as tdist
        tdist.init_process_group("nccl")
    elif args.gpus > 1:
        sys.exit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    run_b200(args, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
