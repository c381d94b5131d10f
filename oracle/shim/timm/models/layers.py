import collections.abc
from itertools import repeat

import torch.nn as nn


class DropPath(nn.Module):
    """Stochastic depth; identity in eval mode (the only mode the oracle runs)."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert not self.training, "oracle shim: DropPath is inference-only"
        return x


def to_2tuple(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return tuple(x)
    return tuple(repeat(x, 2))


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)
