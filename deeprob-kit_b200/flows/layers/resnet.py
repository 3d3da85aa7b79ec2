"""Residual conv conditioner (interface and state_dict layout of deeprob/flows/layers/resnet.py:9-90).
The convolutions are library calls (cuDNN through torch); the hand-written part of the 2D couplings is the
transform / log-det / batch-norm bijector around them (csrc/flows.cu)."""
import torch
from torch import nn

from ...torch.utils import WeightNormConv2d


class ResidualBlock(nn.Module):
    def __init__(self, n_channels: int):
        super().__init__()
        conv = lambda: WeightNormConv2d(n_channels, n_channels, kernel_size=3, padding=1, bias=False)  # noqa: E731
        self.block = nn.Sequential(
            nn.BatchNorm2d(n_channels), nn.ReLU(inplace=True), conv(),
            nn.BatchNorm2d(n_channels), nn.ReLU(inplace=True), conv()
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x + self.block(x)


class ResidualNetwork(nn.Module):
    def __init__(self, in_channels: int, mid_channels: int, out_channels: int, n_blocks: int):
        if n_blocks <= 0:
            raise ValueError("The number of residual blocks must be positve")
        super().__init__()
        self.blocks = nn.ModuleList()
        self.skips = nn.ModuleList()
        self.in_conv = WeightNormConv2d(in_channels, mid_channels, kernel_size=3, padding=1, bias=False)
        self.in_skip = WeightNormConv2d(mid_channels, mid_channels, kernel_size=1, padding=0, bias=True)
        for _ in range(n_blocks):
            self.blocks.append(ResidualBlock(mid_channels))
            self.skips.append(WeightNormConv2d(mid_channels, mid_channels, kernel_size=1, padding=0, bias=True))
        self.out_network = nn.Sequential(
            nn.BatchNorm2d(mid_channels), nn.ReLU(inplace=True),
            WeightNormConv2d(mid_channels, out_channels, kernel_size=1, padding=0, bias=True)
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = self.in_conv(x)
        acc = self.in_skip(h)
        for block, skip in zip(self.blocks, self.skips):
            h = block(h)
            acc = acc + skip(h)
        return self.out_network(acc)
