"""Mask network with the reference ``UNet``'s state_dict layout (addvisor.py:27-84) and the fused
mask head.

The U-Net body (4 encoder blocks, dilated bottleneck, 4 transposed-conv decoders with skip
concatenation) is ordinary cuDNN work that the north star leaves to torch; it is re-declared here
only so that checkpoints trained with the reference (``e1.block.0.weight`` ... ``mask_head.0.bias``,
optionally ``module.``-prefixed by DDP, LMAC_metrics.py:23-25) load unchanged.  The part on our hot
path is the last op, ``mask_head`` = 1x1 conv (32 -> 1) + sigmoid (addvisor.py:57-60,82), which
runs as one streaming sm_100a kernel (``adv_mask_head``): 32 channel planes in, one mask plane out.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def _double_conv(cin, cout, kernel=(3, 3), stride=(1, 1), padding=(1, 1)):
    # addvisor.py:12-25: Conv-BN-LeakyReLU(0.2) twice; only the first conv takes the custom geometry
    return nn.Sequential(
        nn.Conv2d(cin, cout, kernel, stride=stride, padding=padding), nn.BatchNorm2d(cout),
        nn.LeakyReLU(0.2, inplace=True),
        nn.Conv2d(cout, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.LeakyReLU(0.2, inplace=True))


class ConvBlock(nn.Module):
    def __init__(self, in_ch, out_ch, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1)):
        super().__init__()
        self.block = _double_conv(in_ch, out_ch, kernel_size, stride, padding)

    def forward(self, x):
        return self.block(x)


class UNet(nn.Module):
    """x [B,1,F',T'] (F' % 16 == 0, T' % 4 == 0) -> mask [B,1,F',T'] in [0,1]."""

    ENC = (("e1", 1, 32, (5, 3), (2, 1), (2, 1)), ("e2", 32, 64, (5, 3), (2, 1), (2, 1)),
           ("e3", 64, 128, (3, 3), (2, 2), (1, 1)), ("e4", 128, 256, (3, 3), (2, 2), (1, 1)))
    DEC = (("up4", "d4", 512, 256, 128, (2, 2)), ("up3", "d3", 256, 128, 64, (2, 2)),
           ("up2", "d2", 128, 64, 32, (2, 1)), ("up1", "d1", 64, 32, 1, (2, 1)))

    def __init__(self):
        super().__init__()
        for name, cin, cout, k, s, p in self.ENC:
            setattr(self, name, ConvBlock(cin, cout, kernel_size=k, stride=s, padding=p))
        self.bottleneck = nn.Sequential(
            nn.Conv2d(256, 512, 3, padding=2, dilation=2), nn.BatchNorm2d(512), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(512, 512, 3, padding=4, dilation=4), nn.BatchNorm2d(512), nn.LeakyReLU(0.2, inplace=True))
        for up, dec, cin, cout, skip, k in self.DEC:
            setattr(self, up, nn.ConvTranspose2d(cin, cout, kernel_size=k, stride=k))
            setattr(self, dec, ConvBlock(cout + skip, cout))
        self.mask_head = nn.Sequential(nn.Conv2d(32, 1, kernel_size=1), nn.Sigmoid())

    def body(self, x):
        """Everything up to the mask head's input y1 [B,32,F',T'] (addvisor.py:62-80)."""
        skips = [x]
        h = x
        for name, *_ in self.ENC:
            h = getattr(self, name)(h)
            skips.append(h)
        h = self.bottleneck(skips.pop())
        for up, dec, *_ in self.DEC:
            h = getattr(self, dec)(torch.cat([getattr(self, up)(h), skips.pop()], dim=1))
        return h

    def forward(self, x):
        y1 = self.body(x)
        conv = self.mask_head[0]
        if y1.is_cuda and not (torch.is_grad_enabled() and y1.requires_grad):
            return ops.mask_head(y1, conv.weight, conv.bias)   # fused 1x1 conv + sigmoid kernel
        return self.mask_head(y1)                              # autograd (training) keeps torch's ops


def load_checkpoint(model, state_dict):
    """Strip DDP's ``module.`` prefix like LMAC_metrics.py:23-25 and load."""
    if any(k.startswith("module.") for k in state_dict):
        state_dict = {k.replace("module.", ""): v for k, v in state_dict.items()}
    model.load_state_dict(state_dict)
    return model


def seeded_init(model, seed=0):
    """Fill every conv / transposed-conv weight and bias of ``model`` from a generator seeded per state_dict KEY
    (He-scaled normal weights, N(0, 0.01^2) biases; BatchNorm keeps its identity affine and unit running stats).
    The result depends only on the key names and shapes, not on module construction order, so the reference's
    ``UNet`` and this module's re-declaration get bit-identical weights from the same seed (there is no trained
    checkpoint offline; BASELINE configs[0] uses ``seed=0``)."""
    import zlib
    sd = model.state_dict()
    with torch.no_grad():
        for key in sorted(sd):
            t = sd[key]
            if not t.is_floating_point() or "running_" in key or t.dim() == 0:
                continue
            g = torch.Generator().manual_seed((int(seed) << 20) ^ zlib.crc32(key.encode()))
            if t.dim() >= 3:     # Conv2d / ConvTranspose2d weight
                fan_in = t[0].numel() if t.dim() == 4 else t.numel()
                t.copy_(torch.randn(t.shape, generator=g) * (2.0 / max(fan_in, 1)) ** 0.5)
            elif key.endswith(".bias") and not _is_bn(model, key):
                t.copy_(0.01 * torch.randn(t.shape, generator=g))
    return model


def _is_bn(model, key):
    mod = model
    for part in key.split(".")[:-1]:
        mod = getattr(mod, part) if not part.isdigit() else mod[int(part)]
    return isinstance(mod, nn.BatchNorm2d)
