"""Host-side mirror of reference models/net.py: FeatureNet, Pipeline (same constructor, forward
signature, output dict keys and state_dict keys -- checkpoints saved by the reference's train.py
load with `load_state_dict`, with or without the DataParallel 'module.' prefix stripped).

Pipeline.forward(imgs, proj_matrices, depth_min, depth_max)            reference net.py:78
    test mode  -> {"depths_upsampled", "confidence_upsampled"}         reference net.py:125-128
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .estimator import IterMVS

Tensor = torch.Tensor


class _ConvBN(nn.Module):
    """conv(no bias) + BatchNorm (+ReLU): key layout `<name>.conv.weight`, `<name>.bn.*` (module.py:6-29)."""

    def __init__(self, cin, cout, stride=1, relu=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)
        self.bn = nn.BatchNorm2d(cout)
        self._relu = relu

    def forward(self, x):
        y = self.bn(self.conv(x))
        return F.relu(y, inplace=True) if self._relu else y


class _ResBlock(nn.Module):
    """module.py:32-50."""

    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.conv1 = _ConvBN(cin, cout, stride=stride, relu=True)
        self.conv2 = _ConvBN(cout, cout, relu=False)
        self.downsample = None if stride == 1 else _ConvBN(cin, cout, stride=stride, relu=False)

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        if self.downsample is not None:
            x = self.downsample(x)
        return F.relu(x + y, inplace=True)


class FeatureNet(nn.Module):
    """net.py:7-66 -- FPN feature extractor.  Outside the hot path named by the north star (SURVEY
    section 8 row f-1): convolutions run on the stock cuDNN path; what this class adds is batching
    all views of all reference views into one pass in eval mode (the reference loops over views,
    net.py:56-65) and emitting the channels-last pyramids the fused kernels consume."""

    def __init__(self, test=False):
        super().__init__()
        self.test = test
        self.conv1 = _ConvBN(3, 8)
        self.layer1 = nn.Sequential(_ResBlock(8, 16, stride=2), _ResBlock(16, 16))
        self.layer2 = nn.Sequential(_ResBlock(16, 32, stride=2), _ResBlock(32, 32))
        self.layer3 = nn.Sequential(_ResBlock(32, 48, stride=2), _ResBlock(48, 48))
        self.output3 = nn.Conv2d(48, 48, 3, stride=1, padding=1)
        self.output2 = nn.Conv2d(48, 32, 3, stride=1, padding=1)
        self.output1 = nn.Conv2d(48, 16, 3, stride=1, padding=1)
        self.inner1 = nn.Conv2d(16, 48, 1, stride=1, padding=0, bias=True)
        self.inner2 = nn.Conv2d(32, 48, 1, stride=1, padding=0, bias=True)
        self.inner3 = nn.Conv2d(48, 48, 1, stride=1, padding=0, bias=True)   # unused in forward, as in net.py:25

    def _pyramid(self, x: Tensor) -> Dict[str, Tensor]:
        f1 = self.layer1(self.conv1(x))
        f2 = self.layer2(f1)
        f3 = self.layer3(f2)
        out3 = self.output3(f3)
        intra = F.interpolate(f3, scale_factor=2, mode="bilinear") + self.inner2(f2)
        out2 = self.output2(intra)
        intra = F.interpolate(intra, scale_factor=2, mode="bilinear") + self.inner1(f1)
        return {"level3": out3, "level2": out2, "level1": self.output1(intra)}

    def forward_batched(self, x: Tensor) -> Dict[str, Tensor]:
        """x [B,V,3,H,W] -> NCHW pyramids [B*V,C_l,H_l,W_l] (all views in one pass)."""
        b, v, _, h, w = x.shape
        if self.training and self.test:
            # BatchNorm in training mode uses per-call batch statistics: keep the reference's per-view calls
            per = [self._pyramid(x[:, i]) for i in range(v)]
            return {k: torch.stack([p[k] for p in per], dim=1).flatten(0, 1) for k in per[0]}
        return self._pyramid(x.reshape(b * v, 3, h, w))

    def forward(self, x: Tensor):
        """Reference return format (net.py:35-66): dict level -> sequence of per-view [B,C,H,W]."""
        b, v = x.shape[:2]
        pyr = self.forward_batched(x)
        return {k: list(torch.unbind(t.view(b, v, *t.shape[1:]), dim=1)) for k, t in pyr.items()}


class Pipeline(nn.Module):
    """net.py:68-128."""

    def __init__(self, iteration=4, test=False):
        super().__init__()
        self.feature_dim = [8, 16, 32, 48]
        self.hidden_dim = 32
        self.test = test
        self.feature_net = FeatureNet(test=test)
        self.iter_mvs = IterMVS(iteration, self.feature_dim[2], self.hidden_dim, test)

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts checkpoints with the DataParallel 'module.' prefix (train.py:153-157) too."""
        if state_dict and all(k.startswith("module.") for k in state_dict):
            state_dict = {k[7:]: v for k, v in state_dict.items()}
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def forward(self, imgs, proj_matrices, depth_min, depth_max):
        if not self.test:
            raise NotImplementedError("itermvs_b200.Pipeline(test=False): training forward is not built in this round "
                                      "(DESIGN.md, 'next'); use test=True")
        x = imgs["level_0"]
        if not x.is_cuda:
            raise RuntimeError("itermvs_b200.Pipeline: inputs must be CUDA tensors (there is no CPU path)")
        b, v = x.shape[:2]
        pyr = self.feature_net.forward_batched(x.float())
        fea = {}
        for k, t in pyr.items():
            n, c, h, w = t.shape
            fea[k] = ops.nchw_to_nhwc(t).view(b, v, h, w, c)
        ref2 = pyr["level2"].view(b, v, *pyr["level2"].shape[1:])[:, 0].contiguous()
        projs = [ops._chk(proj_matrices[f"level_{l}"].float(), "proj_matrices") for l in (1, 2, 3)]   # net.py:96-98
        flag = ops.NanFlag(x.device)
        depth, depth_up, conf, conf_up = self.iter_mvs.forward_packed(
            fea["level1"], fea["level2"], fea["level3"], ref2, projs[0], projs[1], projs[2],
            ops._chk(depth_min.float(), "depth_min"), ops._chk(depth_max.float(), "depth_max"), nan_flag=flag)
        self._last_nan_flag = flag            # checked lazily by callers that synchronise (tests, eval loop)
        return {"depths_upsampled": depth_up, "confidence_upsampled": conf_up}


def full_loss(*args, **kwargs):
    raise NotImplementedError("itermvs_b200.full_loss: the training loss (net.py:131-190) stays with the reference's "
                              "PyTorch implementation; it is outside the hot path (SURVEY.md section 2, row 3)")
