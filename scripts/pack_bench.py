"""Host-side 2-bit packer: correctness against numpy and throughput per thread count.
    python scripts/pack_bench.py [--pinned]"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tidypopgen_b200", "libtpg_b200.so"))
lib.tpg_pk2_stride.restype = C.c_int64
lib.tpg_pk2_stride.argtypes = [C.c_int64]
lib.tpg_pk2_pack_host.restype = None
lib.tpg_pk2_pack_host.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]

def ref_pack(raw, nrow, ncol, r0, span):
    st = lib.tpg_pk2_stride(span)
    g = np.minimum(raw.reshape(ncol, nrow)[:, r0:r0 + span], 3).astype(np.uint8)
    pad = np.full((ncol, st * 4), 3, np.uint8)
    pad[:, :span] = g
    q = pad.reshape(ncol, st, 4)
    return (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)

rng = np.random.default_rng(0)
for nrow, ncol, r0, span in [(929, 300, 0, 929), (131, 77, 5, 101), (7, 6, 0, 7), (1024, 50, 128, 512), (4097, 33, 1, 4095)]:
    raw = rng.integers(0, 256, nrow * ncol).astype(np.uint8)
    raw[rng.random(raw.size) < 0.8] %= 4
    st = lib.tpg_pk2_stride(span)
    out = np.zeros((ncol, st), np.uint8)
    lib.tpg_pk2_pack_host(raw.ctypes.data + r0, nrow, span, ncol, out.ctypes.data)
    assert np.array_equal(out, ref_pack(raw, nrow, ncol, r0, span)), (nrow, ncol, r0, span)
print("pk2 pack == numpy reference")

nrow, ncol = 929, 650000
pinned = "--pinned" in sys.argv
if pinned:
    import torch
    src_t = torch.empty(nrow * ncol, dtype=torch.uint8).pin_memory()
    src = src_t.numpy()
    dst_t = torch.empty(ncol * lib.tpg_pk2_stride(nrow), dtype=torch.uint8).pin_memory()
    dst = dst_t.numpy()
else:
    src = np.empty(nrow * ncol, np.uint8)
    dst = np.empty(ncol * lib.tpg_pk2_stride(nrow), np.uint8)
src[:] = rng.integers(0, 4, src.size, dtype=np.uint8)
for th in (1, 2, 4, 8, 16, 32):
    if th > 2 * (os.cpu_count() or 1):
        break
    os.environ["TPG_HOST_THREADS"] = str(th)
    lib.tpg_pk2_pack_host(src.ctypes.data, nrow, nrow, ncol, dst.ctypes.data)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        lib.tpg_pk2_pack_host(src.ctypes.data, nrow, nrow, ncol, dst.ctypes.data)
        best = min(best, time.perf_counter() - t0)
    print(f"threads={th:2d} pack {src.size/1e6:.0f} MB in {best*1e3:.2f} ms = {src.size/best/1e9:.1f} GB/s (pinned={pinned})", flush=True)
