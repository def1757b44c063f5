import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from musicfpaugment_b200 import lib, synth
import bench
dev = torch.device("cuda", 0)
ctx = lib.Context(0)
p = lib.afp_defaults()
B, T = 10000, 64000
x = synth.music_like(B, seed=1234, device=dev, chunk=32)
ir = synth.impulse_responses(B, length=8000, seed=2000, device=dev)
prm, pr = bench.aug_param_array(lib, B, 4000, 8000)
noise = synth.rms_noise(B, seed=3000, device=dev)
x_host = torch.empty(B, T, dtype=torch.float32).pin_memory(); x_host.copy_(x)
rows = torch.empty(4_000_000, 2, dtype=torch.int32).pin_memory(); offs = torch.empty(B + 1, dtype=torch.int64).pin_memory()
for _ in range(2):
    ctx.augment_fingerprint_host(x_host, prm, 1, p, ir=ir, noise=noise, rows=rows, offsets=offs)
torch.cuda.synchronize()
os.environ["MFPA_DEBUG_PIPE"] = "1"
t0 = time.perf_counter()
ctx.augment_fingerprint_host(x_host, prm, 1, p, ir=ir, noise=noise, rows=rows, offsets=offs)
torch.cuda.synchronize()
print("wall", round((time.perf_counter() - t0) * 1e3, 2), "ms", file=sys.stderr)
# plain fingerprint host for comparison
del os.environ["MFPA_DEBUG_PIPE"]
for _ in range(2):
    t0 = time.perf_counter(); ctx.fingerprint_host(x_host, 1, p, rows=rows, offsets=offs); torch.cuda.synchronize()
    print("fingerprint_host wall", round((time.perf_counter() - t0) * 1e3, 2), file=sys.stderr)
