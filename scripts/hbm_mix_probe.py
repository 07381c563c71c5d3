"""What a write-heavy stream can reach on this GPU: pure fill, pure read (sum), copy, and a 1 B read :
3 B write mix (the decode kernel's ratio), with torch ops and CUDA events.  Context for decode_planes'
roofline fraction (the MEASURED_PEAKS figure is a 1:1 copy)."""
import json, torch
dev = "cuda:0"
def timeit(fn, reps=10):
    for _ in range(3): fn()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
N = 1 << 31
a = torch.empty(N, dtype=torch.uint8, device=dev)
b = torch.empty(N, dtype=torch.uint8, device=dev)
out = {}
out["fill_GBps"] = N / timeit(lambda: a.zero_()) / 1e6
out["copy_GBps(read+write)"] = 2 * N / timeit(lambda: b.copy_(a)) / 1e6
x = torch.empty(N // 4, dtype=torch.int32, device=dev)
out["read_GBps"] = (N) / timeit(lambda: x.sum()) / 1e6
src = torch.empty(N // 4, dtype=torch.uint8, device=dev)
dst = torch.empty((3, N // 4), dtype=torch.uint8, device=dev)
out["mix_1r3w_GBps"] = N / timeit(lambda: dst.copy_(src.unsqueeze(0).expand(3, -1))) / 1e6
print(json.dumps(out))
