"""Time conv variants for the UNet layer shapes (GPU box)."""
import sys, torch
sys.path.insert(0, '.')
from musicfpaugment_b200 import lib
ctx = lib.Context(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
layers = [  # name, H, W, cin, cout
    ("inc.3", 257, 251, 64, 64), ("u4.0", 257, 251, 128, 64),
    ("d1.0", 128, 125, 64, 128), ("d1.3", 128, 125, 128, 128), ("u3.0", 128, 125, 256, 128),
    ("d2.0", 64, 62, 128, 256), ("d2.3", 64, 62, 256, 256), ("u2.0", 64, 62, 512, 256),
    ("d3.0", 32, 31, 256, 512), ("d3.3", 32, 31, 512, 512), ("u1.0", 32, 31, 1024, 512),
    ("d4.0", 16, 15, 512, 1024), ("d4.3", 16, 15, 1024, 1024),
]
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
def variants(H, W, cin, cout):
    bn = 256 if cout >= 256 else cout
    v = [dict(bn=bn, mt=2), dict(bn=bn, mt=1)]
    for wh in sorted({W + 2, (W + 1) // 2 + 2, (W + 2) // 3 + 2, 32, 34, 18, 66}):
        if wh > 256 or wh < 10: continue
        for mt in (1, 2):
            for st in (2, 3):
                v.append(dict(bn=bn, mt=mt, halo_wh=wh, stages=st))
                if cout == 64:
                    v.append(dict(bn=bn, mt=mt, halo_wh=wh, stages=st, wres=1))
    return v
for name, H, W, cin, cout in layers:
    if only and name not in only: continue
    x = torch.randn(N, H, W, cin, device='cuda').to(torch.bfloat16)
    w = (torch.randn(cout, 9, cin, device='cuda') / (9 * cin) ** 0.5).to(torch.bfloat16)
    sc = torch.ones(cout, device='cuda'); sh = torch.zeros(cout, device='cuda')
    out = torch.empty(N, H, W, cout, dtype=torch.bfloat16, device='cuda')
    gflop = 2 * 9 * cin * cout * H * W * N / 1e9
    res = []
    for kw in variants(H, W, cin, cout):
        try:
            for _ in range(2): lib.conv_bf16(ctx, x, w, sc, sh, out=out, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): lib.conv_bf16(ctx, x, w, sc, sh, out=out, **kw)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            res.append((ms, kw))
        except Exception as ex:
            pass
    res.sort(key=lambda r: r[0])
    print(name, H, W, cin, cout, f"{gflop:.1f} GFLOP")
    for ms, kw in res[:6]:
        print(f"   {ms:7.3f} ms {gflop/ms:7.0f} TF/s  {kw}")
    base = [r for r in res if 'halo_wh' not in r[1]]
    for ms, kw in base: print(f"   base {ms:7.3f} ms {gflop/ms:7.0f} TF/s  {kw}")
    sys.stdout.flush()
