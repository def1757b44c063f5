"""Probe of the host pipeline of mfpa_augment_fingerprint_host: stage times inside the pipeline vs standalone."""
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
g = torch.Generator(device=dev); g.manual_seed(3000)
bank = torch.randn(1 << 24, generator=g, device=dev)
src = np.random.default_rng(3100).integers(0, (1 << 24) - T, B)
pieces = np.zeros(B, dtype=lib.NOISE_PIECE_DTYPE)
pieces["src_a"], pieces["src_b"], pieces["query"], pieces["dst"], pieces["len"] = src, -1, np.arange(B), 0, T
noise = ctx.noise_assemble(bank, pieces, B, T)
x_host = torch.empty(B, T, dtype=torch.float32).pin_memory(); x_host.copy_(x)
rows = torch.empty(4_000_000, 2, dtype=torch.int32).pin_memory(); offs = torch.empty(B + 1, dtype=torch.int64).pin_memory()

def run(tag, **kw):
    ctx.augment_fingerprint_host(x_host, prm, 1, p, ir=ir, rows=rows, offsets=offs, **kw)
    torch.cuda.synchronize()
    ctx.set_option(lib.OPT_STAGE_TIMES, 1)
    t0 = time.perf_counter()
    ctx.augment_fingerprint_host(x_host, prm, 1, p, ir=ir, rows=rows, offsets=offs, **kw)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    st, n = ctx.stage_times()
    ctx.set_option(lib.OPT_STAGE_TIMES, 0)
    print(tag, round(ms, 2), "ms; per-chunk stage ms (mean of", n, "chunks):", {k: round(v, 3) for k, v in st.items()}, "sum", round(sum(st.values()), 3))

run("pieces ", noise_bank=bank, pieces=pieces)
run("rows   ", noise=noise)
os.environ["MFPA_CHUNK_MB"] = "2560"
run("1 chunk", noise=noise)
