"""TEST INFRASTRUCTURE ONLY — fp32 torch restatement of the reference denoiser
``UNet(n_channels=1, n_classes=1, rate, bilinear=False)`` (training/unet.py:8-108).

Parity status: pinned.  ``tests/golden/unet.npz`` holds the output of the REAL
reference class (imported from /root/reference by ``oracle/make_golden_unet.py``)
for ``torch.manual_seed(0)`` initial weights and a seeded input; this module must
reproduce it bit for bit on the CPU with the same seed, which requires the same
parameter creation order and the same ``state_dict`` keys
(``inc.double_conv.0.weight`` … ``outc.conv.bias``) — the keys are also the
order in which ``musicfpaugment_b200.lib.unet_param_blob`` flattens a checkpoint.

Structure followed (file:line of the reference):
  two 3x3 conv(bias=False)+BatchNorm+ReLU per block        training/unet.py:15-22
  down step = MaxPool2d(2) then a block                     :32-34
  up step   = ConvTranspose2d(c, c/2, 2, stride 2), zero-pad the upsampled map at the
              bottom/right to the skip's size, cat([skip, up]), block   :49-65
  head      = 1x1 conv with bias                            :69-72
  dropout between stages is inactive in eval mode           :83,99-103
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

WIDTHS = (64, 128, 256, 512, 1024)


class _Block(nn.Module):
    """(3x3 conv, BatchNorm, ReLU) twice; attribute name kept for the state_dict keys."""

    def __init__(self, c_in: int, c_out: int):
        super().__init__()
        layers = []
        for a, b in ((c_in, c_out), (c_out, c_out)):
            layers += [nn.Conv2d(a, b, 3, padding=1, bias=False), nn.BatchNorm2d(b), nn.ReLU(inplace=True)]
        self.double_conv = nn.Sequential(*layers)

    def forward(self, x):
        return self.double_conv(x)


_block = _Block


class OracleUNet(nn.Module):
    def __init__(self, rate: float = 0.05):
        super().__init__()
        self.dropout = nn.Dropout(rate)
        self.inc = _block(1, WIDTHS[0])
        for i in range(1, 5):
            stage = nn.Module()
            stage.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), _block(WIDTHS[i - 1], WIDTHS[i]))
            setattr(self, f"down{i}", stage)
        for i in range(1, 5):
            c = WIDTHS[5 - i]
            stage = nn.Module()
            stage.up = nn.ConvTranspose2d(c, c // 2, kernel_size=2, stride=2)
            stage.conv = _block(c, c // 2)
            setattr(self, f"up{i}", stage)
        head = nn.Module()
        head.conv = nn.Conv2d(WIDTHS[0], 1, kernel_size=1)
        self.outc = head

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        skips = [self.inc(x)]
        for i in range(1, 5):
            y = getattr(self, f"down{i}").maxpool_conv(skips[-1])
            skips.append(self.dropout(y))
        y = skips.pop()
        for i in range(1, 5):
            stage = getattr(self, f"up{i}")
            skip = skips.pop()
            y = stage.up(y)
            dh, dw = skip.shape[2] - y.shape[2], skip.shape[3] - y.shape[3]
            y = F.pad(y, [dw // 2, dw - dw // 2, dh // 2, dh - dh // 2])
            y = stage.conv(torch.cat([skip, y], dim=1))
            if i == 1:
                y = self.dropout(y)
        return self.outc.conv(y)


def seeded_unet(seed: int = 0, randomize_bn: bool = True) -> OracleUNet:
    """Random-init network (BASELINE.json configs[3]: "random init").  With ``randomize_bn`` the
    BatchNorm affine parameters and running statistics get seeded non-trivial values so the folded
    scale/shift path is exercised."""
    torch.manual_seed(seed)
    net = OracleUNet().eval()
    if randomize_bn:
        g = torch.Generator().manual_seed(seed + 1)
        for m in net.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.weight.data = 0.8 + 0.4 * torch.rand(m.num_features, generator=g)
                m.bias.data = 0.1 * torch.randn(m.num_features, generator=g)
                m.running_mean.data = 0.1 * torch.randn(m.num_features, generator=g)
                m.running_var.data = 0.8 + 0.4 * torch.rand(m.num_features, generator=g)
    return net
