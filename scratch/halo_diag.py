import os, sys, torch
sys.path.insert(0, '.')
from musicfpaugment_b200 import lib
import numpy as np
ctx = lib.Context(0)
torch.backends.cudnn.allow_tf32 = False
def ref_conv(x, w):
    xf = x.float().permute(0,3,1,2); cout,_,cin = w.shape
    wf = w.float().reshape(cout,3,3,cin).permute(0,3,1,2)
    return torch.nn.functional.conv2d(xf, wf, padding=1).permute(0,2,3,1)
N,H,W,cin,cout = 1, 20, 30, 64, 64
g = torch.Generator(device='cuda').manual_seed(1)
x = torch.randn(N,H,W,cin, device='cuda', generator=g).to(torch.bfloat16)
wfull = (torch.randn(cout,9,cin, device='cuda', generator=g)/8).to(torch.bfloat16)
one = torch.ones(cout, device='cuda'); zero = torch.zeros(cout, device='cuda')
for wh in (32, 18):
    for tap in range(9):
        w = torch.zeros_like(wfull); w[:,tap] = wfull[:,tap]
        out = lib.conv_bf16(ctx, x, w, one, zero, relu=False, taps=9, bn=64, mt=1, halo_wh=wh)
        torch.cuda.synchronize()
        ref = ref_conv(x, w)
        err = (out.float()-ref).abs()
        bad = (err > 0.02*ref.abs()+1e-2)
        # which pixels bad
        bp = bad.any(dim=-1)[0]
        print(f"bo={os.environ.get('MFPA_HALO_BO')} wh={wh} tap={tap} maxerr={float(err.max()):.3f} badpix={int(bp.sum())}/{H*W}", "rows bad:", sorted(set(torch.nonzero(bp)[:,0].tolist()))[:8], "cols bad:", sorted(set(torch.nonzero(bp)[:,1].tolist()))[:10])
