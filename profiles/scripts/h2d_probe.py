"""Does a pinned->device cudaMemcpyAsync keep its rate while bandwidth-bound kernels run on another stream?"""
import time, torch
dev = torch.device("cuda", 0)
n = 640 * 1024 * 1024 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device=dev)
a = torch.empty(1 << 28, dtype=torch.float32, device=dev); b = torch.empty_like(a)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def h2d(reps=4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s1):
        e0.record()
        for _ in range(reps): d.copy_(h, non_blocking=True)
        e1.record()
    return e0, e1
for load in ("idle", "copy kernels (HBM-bound)", "matmul (compute-bound)"):
    torch.cuda.synchronize()
    stop = False
    with torch.cuda.stream(s2):
        if load.startswith("copy"):
            for _ in range(60): b.copy_(a)
        elif load.startswith("matmul"):
            m = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
            for _ in range(40): m2 = m @ m
    e0, e1 = h2d()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    print(f"{load:28s}: H2D of 640 MiB in {ms:7.2f} ms = {0.671 / ms * 1e3:6.1f} GB/s")
