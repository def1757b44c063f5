"""UNet denoiser (BASELINE.json configs[3]; training/unet.py:75-108): the tcgen05 implicit-GEMM
convolutions against torch fp32 on the same bf16-rounded operands, and the whole network against the
pinned fp32 oracle (oracle/unet_torch.py) within a bf16 tolerance stated in each test."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _conv_ref(x, w, scale, shift, relu, taps):
    """fp32 reference of one layer on the bf16-rounded operands.  x [N,H,W,C] bf16, w [Cout,taps,Cin] bf16."""
    xf = x.float().permute(0, 3, 1, 2)
    cout, _, cin = w.shape
    k = 3 if taps == 9 else 1
    wf = w.float().reshape(cout, k, k, cin).permute(0, 3, 1, 2)
    y = torch.nn.functional.conv2d(xf, wf, padding=k // 2)
    y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1)


CASES = [
    # N, H, W, cin, cout, taps, bn, mt, stages
    (2, 20, 19, 64, 64, 9, 64, 1, 2),
    (2, 20, 19, 64, 64, 9, 64, 2, 4),
    (1, 33, 37, 128, 128, 9, 128, 2, 3),
    (1, 16, 15, 128, 256, 9, 256, 2, 3),
    (1, 16, 15, 256, 256, 9, 256, 1, 0),
    (3, 9, 50, 64, 128, 9, 64, 1, 0),
    (2, 17, 16, 192, 64, 1, 64, 2, 0),
    (1, 40, 40, 64, 256, 1, 128, 2, 0),
]


@pytest.mark.parametrize("N,H,W,cin,cout,taps,bn,mt,stages", CASES)
def test_conv_gemm_matches_torch(mfpa_ctx, N, H, W, cin, cout, taps, bn, mt, stages):
    from musicfpaugment_b200 import lib

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + H * 10 + cin)
    x = torch.randn(N, H, W, cin, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, taps, cin, device="cuda", generator=g) / (taps * cin) ** 0.5).to(torch.bfloat16)
    scale = 0.5 + torch.rand(cout, device="cuda", generator=g)
    shift = 0.2 * torch.randn(cout, device="cuda", generator=g)
    # write into a channel slice of a wider buffer (the skip-concatenation layout)
    out = torch.full((N, H, W, cout + 64), 7.0, dtype=torch.bfloat16, device="cuda")
    lib.conv_bf16(mfpa_ctx, x, w, scale, shift, relu=True, taps=taps, out=out, coff=64, bn=bn, mt=mt, stages=stages)
    torch.cuda.synchronize()
    ref = _conv_ref(x, w, scale, shift, True, taps)
    got = out[..., 64:].float()
    # bf16 output rounding: 2^-8 relative, plus fp32 accumulation-order noise
    err = (got - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 1e-3
    assert bool((err <= tol).all()), f"max err {float(err.max())} at ref {float(ref.abs().max())}"
    assert bool((out[..., :64] == 7.0).all()), "channels outside the slice were overwritten"


HALO_CASES = [
    # N, H, W, cin, cout, bn, mt, halo_wh, wres
    (2, 20, 19, 64, 64, 64, 1, 18, 0),
    (2, 20, 19, 64, 64, 64, 2, 18, 1),
    (1, 33, 70, 128, 64, 64, 2, 18, 1),
    (1, 33, 70, 128, 64, 64, 1, 32, 0),
    (1, 40, 125, 64, 128, 128, 2, 127, 0),
    (1, 17, 251, 64, 64, 64, 2, 86, 1),
    (1, 17, 251, 128, 256, 256, 1, 66, 0),
    (2, 16, 15, 256, 256, 256, 1, 17, 0),
]


@pytest.mark.parametrize("N,H,W,cin,cout,bn,mt,wh,wres", HALO_CASES)
def test_conv_halo_kernel_matches_torch(mfpa_ctx, N, H, W, cin, cout, bn, mt, wh, wres):
    """The halo-reuse 3x3 kernel (one halo load per 64-channel chunk, nine shifted operand views)."""
    from musicfpaugment_b200 import lib

    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(H * 100 + W + cin)
    x = torch.randn(N, H, W, cin, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, 9, cin, device="cuda", generator=g) / (9 * cin) ** 0.5).to(torch.bfloat16)
    scale = 0.5 + torch.rand(cout, device="cuda", generator=g)
    shift = 0.2 * torch.randn(cout, device="cuda", generator=g)
    out = torch.full((N, H, W, cout), 7.0, dtype=torch.bfloat16, device="cuda")
    lib.conv_bf16(mfpa_ctx, x, w, scale, shift, relu=True, taps=9, out=out, bn=bn, mt=mt, halo_wh=wh, wres=wres)
    torch.cuda.synchronize()
    ref = _conv_ref(x, w, scale, shift, True, 9)
    err = (out.float() - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 1e-3
    assert bool((err <= tol).all()), f"max err {float(err.max())} at ref {float(ref.abs().max())}"


def test_unet_forward_matches_fp32_oracle(mfpa_ctx):
    """Whole network vs the fp32 oracle.  Tolerance: 23 bf16-rounded layers -> <= 3 % of the output's
    dynamic range at the worst pixel, <= 0.5 % rms."""
    from musicfpaugment_b200 import lib
    from oracle.unet_torch import seeded_unet

    torch.backends.cudnn.allow_tf32 = False
    net = seeded_unet(0).cuda()
    den = lib.UNetDenoiser(mfpa_ctx, net.state_dict(), max_chunk=2)
    rng = np.random.default_rng(5)
    for (B, H, W) in ((3, 257, 251), (2, 257, 250), (1, 64, 48)):
        x = torch.from_numpy(rng.random((B, H, W), dtype=np.float32) ** 4).cuda()
        got = den.forward(x)
        with torch.no_grad():
            ref = net(x[:, None])[:, 0]
        torch.cuda.synchronize()
        span = float(ref.max() - ref.min())
        err = (got - ref).abs()
        assert float(err.max()) <= 0.03 * span, (B, H, W, float(err.max()), span)
        assert float(err.pow(2).mean().sqrt()) <= 0.005 * span, (float(err.pow(2).mean().sqrt()), span)
    den.close()


def test_unet_golden_vector(mfpa_ctx):
    """tests/golden/unet.npz: output of the REAL reference UNet (manual_seed(0) weights)."""
    import os

    from musicfpaugment_b200 import lib
    from oracle.unet_torch import seeded_unet

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "unet.npz"))
    net = seeded_unet(0, randomize_bn=False)
    den = lib.UNetDenoiser(mfpa_ctx, net.state_dict())
    got = den.forward(torch.from_numpy(g["x"]).cuda()).cpu().numpy()
    span = float(g["y"].max() - g["y"].min())
    assert np.abs(got - g["y"]).max() <= 0.03 * span
    den.close()


def test_denoise_mag_in_place_and_peaks(mfpa_ctx):
    """find_peaks with denoising (peak_extractor.py:263-269): stft -> /max -> UNet -> picker.  The in-place
    frame-major entry must equal the plain [B,H,W] entry, and the picker must accept its output."""
    from musicfpaugment_b200 import lib, synth
    from oracle.unet_torch import seeded_unet

    net = seeded_unet(0).cuda()
    den = lib.UNetDenoiser(mfpa_ctx, net.state_dict())
    x = synth.music_like(3, seed=11, device=torch.device("cuda"))
    mag, qmax = mfpa_ctx.stft_mag(x, 1)
    plain = den.forward((mag[:, :, :257] / qmax[:, None, None]).permute(0, 2, 1).contiguous())
    den.denoise_mag(mag, qmax)
    torch.cuda.synchronize()
    assert torch.equal(mag[:, :, :257].permute(0, 2, 1), plain)
    p = lib.afp_defaults()
    rec, npk = mfpa_ctx.audfprint_peaks(mag, None, x.shape[1], 1, p)
    assert int(npk.min()) > 0
    den.close()
