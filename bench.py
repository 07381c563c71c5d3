"""bench.py — pairwise_king pair-SNP ops/s (n^2 * m / s) on BASELINE.json configs[1]
(HGDP-shape synthetic: 929 individuals x 650,000 SNPs, 2% missing), on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step = one whole pairwise_king job: decode_planes_k4 + gram_mxf4_pair (tcgen05 kind::mxf4 on exact
4-bit operands; TPG_GRAM_I8=1 = the int8 -> s32 kernels) + finish_king.
`value` is measured with the .bk bytes resident in HBM; `e2e` goes through the public API
with HOST (pinned) buffers, H2D and D2H inside the timed region.  N > 1 is loci-sharded
(weak scaling: every rank owns its own 650,000 loci, global m = N * 650,000): staged C-ABI
accumulate, ONE NCCL all-reduce of the int32 count matrices, finish.  Prints one JSON line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_IND, M_LOCI, MISSING, SEED = 929, 650_000, 0.02, 20260002
WORKLOAD = "HGDP-shape synthetic 929 x 650000, 2% missing, pairwise_king (BASELINE.json configs[1])"
METRIC = "pairwise_king pair-SNP ops/sec (n^2*m)"
UNIT = "pair-SNP/s"
SYNC_PATH = os.environ.get("TPG_BENCH_SYNC") == "1"
KING_INT8_OPS_PER_PAIRSNP = 4.0   # 2.0 MAC: DD (upper) + NN (upper) + HN (full) — DESIGN.md "Algebra"
DECODE_BYTES_PER_GENOTYPE = 4.0   # 1 read + 3 plane bytes written


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
        flag = lambda name: "Active" if r & getattr(n, name, 0) else "Not Active"  # noqa: E731
        self.rows.append([str(self.index), str(sm), str(mx), "0",
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
            time.sleep(0.02 if self.nvml is not None else 0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({nm for r in self.rows if len(r) >= 8 for nm, v in zip(names, r[4:8]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "NVML" if self.nvml is not None else "nvidia-smi"}


def measure_int8_library_peak(local):
    """Dense int8 x int8 -> s32 GEMM through torch._int_mm (cuBLASLt) at 8192^3, best of 10 with CUDA
    events: a measured int8 tensor-core figure to read the gram kernel against, beside the
    2 x bf16 peak of MEASURED_PEAKS.json (which has no int8 entry)."""
    import torch
    try:
        dev = f"cuda:{local}"
        a = torch.randint(-1, 2, (8192, 8192), dtype=torch.int8, device=dev)
        b = torch.randint(-1, 2, (8192, 8192), dtype=torch.int8, device=dev)
        for _ in range(3):
            torch._int_mm(a, b)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return {"tops": 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12, "how": "torch._int_mm 8192^3 (cuBLASLt), best of 10"}
    except Exception as ex:  # noqa: BLE001
        return {"tops": None, "how": f"torch._int_mm unavailable: {ex}"} if False else None


def traffic_from_profile():
    """dram bytes per gram_i8 launch from the committed ncu capture, if one exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return {}


# ---------------------------------------------------------------------------- reference arm

def synth_numpy(n, m, seed):
    """Same law as include/tpg_synth.h (p ~ U(0.05, 0.5), Binomial(2, p), iid missing -> 3),
    drawn with numpy so the reference arm touches none of this repo's native code."""
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.05, 0.5, size=(m, 1))
    g = (rng.random((m, n)) < p).astype(np.uint8) + (rng.random((m, n)) < p).astype(np.uint8)
    g[rng.random((m, n)) < MISSING] = 3
    return g


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
    bs = O.block_size(N_IND)                   # pairwise_king passes block_size(nrow(x)): 144,475
    iv = O.split_len(M_LOCI, bs)               # 5 blocks of 130,000 loci
    sample = int(iv[0, 2])
    raw = synth_numpy(N_IND, sample, SEED)
    t0 = time.perf_counter()
    B.king_blocks(raw, N_IND, sample, bs, max_blocks=1)
    first = time.perf_counter() - t0
    if first > 8.0:                            # keep the whole run within minutes on a small host
        sample = max(4096, int(sample * 4.0 / first))
        raw = raw[:sample]
    for _ in range(max(0, args.warmup - 1)):
        B.king_blocks(raw, N_IND, sample, bs, max_blocks=1)
    times = []
    for _ in range(args.steps):
        K, N, done, sec = B.king_blocks(raw, N_IND, sample, bs, max_blocks=1)
        t1 = time.perf_counter()
        B.king_finish(K, N)
        times.append(sec + (time.perf_counter() - t1) * (sample / M_LOCI))
    sec = float(np.mean(times))
    value = N_IND * N_IND * sample / sec
    cores = torch.get_num_threads()
    desc = (f"1 loci block of {sample} of 650000 loci (R block.size {bs}), serial-equivalent numpy decode + "
            f"FP64 dgemm via MKL on {cores} threads; rate is per pair-SNP so it extrapolates linearly in m")
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ---------------------------------------------------------------------------- B200 arm

def run_b200(args, rank, world, local):
    import torch
    import torch.distributed as dist

    import tidypopgen_b200 as tpg
    from tidypopgen_b200 import _lib as L
    from tidypopgen_b200 import dist as tdist

    torch.cuda.set_device(local)
    eng = tdist.CudaEngine(local)
    lib = L.lib()
    n, m = N_IND, M_LOCI
    m_global = m * world

    # ---- synthetic .bk image of this rank's loci, resident in HBM
    dev_raw = torch.empty(n * m + 64, dtype=torch.uint8, device=f"cuda:{local}")
    eb = L.errbuf()
    L.check(lib.tpg_synth_fill(eng.ctx.handle, dev_raw.data_ptr(), n, m, rank * m, SEED, MISSING, 0, eb, len(eb)), eb)
    X = tpg.FBMCode256.__new__(tpg.FBMCode256)
    X.backingfile, X.nrow, X.ncol, X.code256 = "<synthetic>", n, m, None
    bs = tpg.block_size(n)
    job, keep = tpg.snp.make_job(X, None, None, L.STAT_KING, bs, device=local, flags=L.FLAG_ASYNC if world > 1 else 0,
                                 src=(L.SRC_DEVICE, dev_raw.data_ptr(), dev_raw))
    counts = eng.new_counts(n)
    out_dev = torch.zeros((n, n), dtype=torch.float64, device=f"cuda:{local}")
    res = L.Result(out0=out_dev.data_ptr(), out1=None, aux0=None, aux1=None, out_on_device=1)
    agg = {"ms_decode": 0.0, "ms_gram": 0.0, "ms_finish": 0.0, "launches": 0, "chunks": 0, "macs": 0, "q4": 0}

    def step(record: bool):
        if world == 1:
            # the public one-call-per-job entry: memset, decode, gram, finish, one sync
            L.check(lib.tpg_pairwise_ctx(eng.ctx.handle, C.byref(job), C.byref(res), eb, len(eb)), eb)
            st, n_l = res.stats, res.stats.n_launches
        else:
            # zero fill, accumulate (TPG_FLAG_ASYNC), all-reduce and finish queue up on the library's
            # stream; the only host wait is at the end of tpg_finish
            with torch.cuda.stream(eng.stream):
                counts.zero_()
            st = L.Stats()
            res.stats = L.Stats()  # tpg_finish adds the collected stage times to what it finds
            L.check(lib.tpg_accumulate(eng.ctx.handle, C.byref(job), counts.data_ptr(), C.byref(st), eb, len(eb)), eb)
            if SYNC_PATH:   # A/B switch: host waits around the all-reduce (the round's first scheme)
                eng.stream.synchronize()
                dist.all_reduce(counts)
                torch.cuda.current_stream().synchronize()
            else:
                with torch.cuda.stream(eng.stream):
                    dist.all_reduce(counts)
            L.check(lib.tpg_finish(eng.ctx.handle, C.byref(job), counts.data_ptr(), C.byref(res), eb, len(eb)), eb)
            st.ms_decode, st.ms_gram, st.ms_finish = res.stats.ms_decode, res.stats.ms_gram, res.stats.ms_finish
            n_l = st.n_launches + res.stats.n_launches + 1  # + the counts fill
        if record:
            agg["ms_decode"] += st.ms_decode
            agg["ms_gram"] += st.ms_gram
            agg["ms_finish"] += st.ms_finish
            agg["launches"] += n_l
            agg["chunks"] += st.n_chunks
            agg["macs"] += st.mma_macs
            agg["q4"] += st.n_mxf4_chunks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.stream)
    for _ in range(args.steps):
        step(True)
    e1.record(eng.stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
    if os.environ.get("TPG_BENCH_DEBUG"):
        print(f"[rank {rank}] {float(ms.item()) / args.steps:.3f} ms/step (device events)", file=sys.stderr, flush=True)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / args.steps
    value = n * n * m_global / (ms_step * 1e-3)

    # ---- e2e: host (pinned) .bk bytes -> public API -> host result, copies inside the timed region
    host_raw = torch.empty(n * m, dtype=torch.uint8).pin_memory()
    host_raw.copy_(dev_raw[: n * m])
    src_h = (L.SRC_HOST, host_raw.data_ptr(), host_raw)

    job_h, keep_h = tpg.snp.make_job(X, None, None, L.STAT_KING, bs, device=local, src=src_h,
                                     flags=L.FLAG_ASYNC if world > 1 else 0)
    e2e_bytes = [0]
    king_host = np.zeros((n, n), np.float64, order="F")
    res_h = L.Result(out0=king_host.ctypes.data, out1=None, aux0=None, aux1=None, out_on_device=0)

    def e2e_step():
        if world == 1:  # the call a user makes: snp_king() on host bytes -> host matrix
            return tpg.snp_king(X, block_size=bs, src=src_h, device=local)
        # N ranks, each with its own host-resident loci: accumulate, all-reduce, finish to host
        with torch.cuda.stream(eng.stream):
            counts.zero_()
        st_h = L.Stats()
        L.check(lib.tpg_accumulate(eng.ctx.handle, C.byref(job_h), counts.data_ptr(), C.byref(st_h), eb, len(eb)), eb)
        with torch.cuda.stream(eng.stream):
            dist.all_reduce(counts)
        L.check(lib.tpg_finish(eng.ctx.handle, C.byref(job_h), counts.data_ptr(), C.byref(res_h), eb, len(eb)), eb)
        e2e_bytes[0] += st_h.bytes_h2d
        return king_host

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    e2e_h2d = 0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        king_e2e = e2e_step()
        if world > 1:
            e2e_h2d, e2e_bytes[0] = e2e_h2d + e2e_bytes[0], 0
        if world == 1:
            e2e_h2d += tpg.snp.last_stats["bytes_h2d"]
            e2e_packed = (tpg.snp.last_stats["n_packed_chunks"], tpg.snp.last_stats["n_chunks"])
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / e2e_steps], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    sampler.stop_flag = True
    sampler.join(2)
    e2e_value = n * n * m_global / (float(e2e_ms.item()) * 1e-3)
    same = bool(np.array_equal(king_e2e, out_dev.t().cpu().numpy(), equal_nan=True))

    # ---- disk -> HBM streaming, reported separately from the kernels (north_star): the same job
    # from the .bk FILE (page cache warm after the first pass), per-stage figures from tpg_stats
    streaming = None
    if world == 1 and not args.no_streaming:
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".bk", prefix="tpg_bench_", delete=True) as tf:
            host_raw.numpy().tofile(tf.name)
            Xf = tpg.FBMCode256(tf.name, n, m)
            tpg.snp_king(Xf, block_size=bs, device=local)           # cold-ish pass (file just written)
            t0 = time.perf_counter()
            king_file = tpg.snp_king(Xf, block_size=bs, device=local)
            sec = time.perf_counter() - t0
            stt = dict(tpg.snp.last_stats)
            streaming = {"source": "FBM .bk file via mmap -> 2-bit packed into pinned staging by the host worker pool -> H2D, page cache warm",
                         "ms_total": sec * 1e3, "value": n * n * m / sec, "unit": UNIT,
                         "GBps_file_to_hbm": n * m / sec / 1e9, "ms_staging_sum": stt["ms_read"], "ms_h2d_sum": stt["ms_h2d"],
                         "ms_decode_sum": stt["ms_decode"], "ms_gram_sum": stt["ms_gram"], "chunks": stt["n_chunks"],
                         "matches_device_run": bool(np.array_equal(king_file, out_dev.t().cpu().numpy(), equal_nan=True))}

    if rank != 0:
        return
    pk = peaks()
    steps = args.steps
    gram_ms = agg["ms_gram"] / max(1, agg["chunks"])          # per launch
    dec_ms = agg["ms_decode"] / max(1, agg["chunks"])
    loci_per_launch = m * steps / max(1, agg["chunks"])
    gram_ops = KING_INT8_OPS_PER_PAIRSNP * n * n * loci_per_launch
    int8_peak = 2.0 * pk["bf16_tflops"]
    int8_lib = measure_int8_library_peak(local)  # cuBLASLt int8 GEMM on this GPU, outside the timed region
    tr = traffic_from_profile()
    q4 = agg["q4"] > 0                         # the step's chunks ran the 4-bit route (tpg_stats.n_mxf4_chunks)
    pair = os.environ.get("TPG_GRAM_CTAS", "2") != "1"
    achieved = gram_ops / (gram_ms * 1e-3) / 1e12
    issued = 2 * agg["macs"] / steps / (agg["ms_gram"] / steps * 1e-3) / 1e12
    if q4:
        # The contraction runs as E2M1 x E2M1 -> f32 (exact for these integer sums): the
        # roofline is the dense 4-bit tensor peak.  MEASURED_PEAKS.json has bf16 only; the 4-bit MMA
        # issues 4x the bf16 MACs per cycle (nominal 9 vs 2.25 PFLOP/s), so peak = 4 x measured bf16.
        peak = 4.0 * pk["bf16_tflops"]
        roof = {"kernel": "gram_mxf4_pair_kernel (tcgen05.mma.cta_group::2.kind::mxf4.block_scale)", "bound": "tensor",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": tr.get("gram_mxf4_pair_dram_bytes_per_launch"),
                "note": ("4-bit tensor ops (2 per MAC); algorithmic = 4 op per pair-SNP (DD+NN upper, HN full); "
                         f"peak = 4 x measured bf16 burst from {pk['source']} (nominal 4-bit dense 9000); "
                         f"issued incl. tile padding = {issued:.0f} TOP/s; the same products on the int8 kernel "
                         "(TPG_GRAM_I8=1) reach 0.84 of 2 x bf16; "
                         f"avg launch {gram_ms:.3f} ms over {agg['chunks'] // steps} launch(es)/step"),
                "int8_library_peak": int8_lib,
                "frac_of_int8_library_peak": (achieved / int8_lib["tops"]) if int8_lib else None}
    else:
        roof = {"kernel": "gram_i8_pair_kernel (tcgen05.mma.cta_group::2)" if pair else "gram_i8_kernel (cta_group::1)",
                "bound": "tensor", "achieved": achieved, "peak": int8_peak, "unit": "TFLOP/s", "frac": achieved / int8_peak,
                "traffic": tr.get("gram_i8_pair_dram_bytes_per_launch" if pair else "gram_i8_dram_bytes_per_launch"),
                "note": ("int8 tensor ops (2 per MAC); algorithmic = 4 op per pair-SNP (DD+NN upper, HN full); "
                         f"peak = 2 x measured bf16 burst from {pk['source']} (nominal int8 dense 4500); "
                         f"issued incl. tile padding = {issued:.0f} TOP/s; "
                         f"avg launch {gram_ms:.3f} ms over {agg['chunks'] // steps} launch(es)/step"),
                "int8_library_peak": int8_lib,
                "frac_of_int8_library_peak": (achieved / int8_lib["tops"]) if int8_lib else None}
    bytes_per_gt = 2.5 if q4 else DECODE_BYTES_PER_GENOTYPE
    dec_bytes = bytes_per_gt * n * loci_per_launch
    roof_dec = {"kernel": "decode_planes_k4_kernel" if q4 else "decode_planes_mn_kernel", "bound": "hbm",
                "achieved": dec_bytes / (dec_ms * 1e-3) / 1e9,
                "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": dec_bytes / (dec_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                "traffic": tr.get("decode_planes_k4_dram_bytes_per_launch" if q4 else "decode_planes_dram_bytes_per_launch"),
                "note": (f"1 B read + {'1.5 (three 4-bit planes)' if q4 else '3'} B written per genotype; "
                         f"avg launch {dec_ms:.3f} ms; peak from {pk['source']}")}

    # ---- CPU baseline on this host: the reference path restated (oracle/blas_baseline.py)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import blas_baseline as B
        from oracle import oracle as O
        iv = O.split_len(m, bs)
        sample = int(iv[0, 2])
        raw_np = host_raw[: n * sample].numpy().reshape(sample, n)
        B.king_blocks(raw_np[:8192], n, 8192, bs, max_blocks=1)  # BLAS thread warm-up
        _, _, done, sec = B.king_blocks(raw_np, n, sample, bs, max_blocks=1)
        cpu = {"value": n * n * done / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"first loci block ({done} of {m} loci, R block.size {bs}); numpy decode + FP64 dgemm (MKL) "
                         f"in {sec:.2f} s; per-pair-SNP rate extrapolates linearly in m"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "e2m1 x e2m1 -> f32 (exact integer sums)" if agg["q4"] else "int8 x int8 -> s32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "n": n, "m_per_gpu": m, "m_global": m_global,
                   "sharding": "none" if world == 1 else f"loci x{world}, one NCCL all-reduce of int32 counts",
                   "l2": "inputs larger than L2 (604 MB raw + 1 GB of 4-bit planes (2 GB int8) per step vs 126 MB), no flush needed",
                   "tensor_ops_per_pair_snp": KING_INT8_OPS_PER_PAIRSNP},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": e2e_h2d // e2e_steps, "d2h_bytes_per_step": 8 * n * n,
                "host_bytes_per_step": n * m,
                "transport": (f"{e2e_packed[0]} of {e2e_packed[1]} chunks cross PCIe 2-bit packed by the host worker pool, "
                              "the rest are DMA'd raw from the caller's pinned buffer (hybrid schedule); both kinds decode "
                              "to 4-bit planes; result through pinned bounce buffers") if world == 1 else
                             "per rank: hybrid packed / raw chunks, staged accumulate + all-reduce + finish on one stream",
                "ms_per_step": float(e2e_ms.item()), "steps": e2e_steps, "matches_device_run": same},
        "gpu_launches": agg["launches"],
        "roofline": roof, "roofline_decode": roof_dec,
        "stage_ms_per_step": {"decode": agg["ms_decode"] / steps, "gram": agg["ms_gram"] / steps,
                              "finish": agg["ms_finish"] / steps},
        "streaming": streaming,
        "cpu_baseline": cpu,
    }
    emit(line)


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-streaming", action="store_true")
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
