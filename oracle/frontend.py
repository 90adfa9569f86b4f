"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the two small conditioning modules the reference's stage-2 driver
runs once per image pair before the denoising loop:

* `ImageProjModel_p` — defined by the reference itself (/root/reference/stage2_batchtest_inpaint_model.py:48-66:
  Linear(1536, 768) -> GELU -> Dropout -> LayerNorm(768) -> Linear(768, 1024) -> Dropout) and applied to the DINOv2
  tokens at :169.  tests/test_frontend.py executes the reference's own class (extracted from that file) and requires
  bit-equality with this restatement.
* `ControlNetConditioningEmbedding(320, 3, (16, 32, 96, 256))` — diffusers 0.24.0 `models/controlnet.py`, instantiated
  at stage2_batchtest_inpaint_model.py:101 as `pose_proj` and applied to the source|target pose canvas at :179.
  diffusers is not vendored: restated from the published module (conv_in, three [conv, stride-2 conv] pairs, conv_out,
  SiLU after every conv but the last); PARITY UNPINNED against diffusers itself, key names identical.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class ImageProjModel_p(nn.Module):
    def __init__(self, in_dim, hidden_dim, out_dim, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.LayerNorm(hidden_dim), nn.Linear(hidden_dim, out_dim), nn.Dropout(dropout))

    def forward(self, x):
        return self.net(x)


class ControlNetConditioningEmbedding(nn.Module):
    def __init__(self, conditioning_embedding_channels: int, conditioning_channels: int = 3,
                 block_out_channels=(16, 32, 96, 256)):
        super().__init__()
        self.conv_in = nn.Conv2d(conditioning_channels, block_out_channels[0], kernel_size=3, padding=1)
        self.blocks = nn.ModuleList([])
        for i in range(len(block_out_channels) - 1):
            cin, cout = block_out_channels[i], block_out_channels[i + 1]
            self.blocks.append(nn.Conv2d(cin, cin, kernel_size=3, padding=1))
            self.blocks.append(nn.Conv2d(cin, cout, kernel_size=3, padding=1, stride=2))
        self.conv_out = nn.Conv2d(block_out_channels[-1], conditioning_embedding_channels, kernel_size=3, padding=1)
        nn.init.zeros_(self.conv_out.weight)   # diffusers zero_module(): trained checkpoints overwrite it
        nn.init.zeros_(self.conv_out.bias)

    def forward(self, conditioning):
        e = F.silu(self.conv_in(conditioning))
        for block in self.blocks:
            e = F.silu(block(e))
        return self.conv_out(e)


def make_frontend(seed: int = 0, pose_channels=(16, 32, 96, 256), out_channels: int = 320, in_dim: int = 1536,
                  hidden_dim: int = 768, out_dim: int = 1024):
    """Seeded random instances (conv_out made non-zero so that the whole stack is exercised)."""
    torch.manual_seed(seed)
    proj = ImageProjModel_p(in_dim, hidden_dim, out_dim).eval()
    pose = ControlNetConditioningEmbedding(out_channels, 3, tuple(pose_channels)).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        fan = pose.conv_out.weight[0].numel()
        pose.conv_out.weight.copy_(torch.randn(pose.conv_out.weight.shape, generator=g) * fan ** -0.5)
        pose.conv_out.bias.copy_(0.05 * torch.randn(pose.conv_out.bias.shape, generator=g))
        proj.net[3].weight.add_(0.1 * torch.randn(hidden_dim, generator=g))
        proj.net[3].bias.add_(0.1 * torch.randn(hidden_dim, generator=g))
    return proj, pose
