"""Host-side mirror of reference models/net.py: FeatureNet, Pipeline (same constructor, forward
signature, output dict keys and state_dict keys -- checkpoints saved by the reference's train.py
load with `load_state_dict`, with or without the DataParallel 'module.' prefix stripped).

Pipeline.forward(imgs, proj_matrices, depth_min, depth_max)            reference net.py:78
    test=True  -> {"depths_upsampled", "confidence_upsampled"}         reference net.py:125-128
    test=False -> {"depths": {"combine","probability","initial"}, "depths_upsampled", "confidences",
                   "confidence_upsampled"}  (net.py:115-120): eval() = forward only on the inference kernels
                   (validation, train.py:257); train() = differentiable training path (training.py)
full_loss(...)                                                         reference net.py:131-190
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import _pack
from . import ops
from .estimator import IterMVS, _cached_pack, _wants_grad

Tensor = torch.Tensor


class _ConvBN(nn.Module):
    """conv(no bias) + BatchNorm: key layout `<name>.conv.weight`, `<name>.bn.*` (module.py:6-29).
    Parameter / buffer holder; the kernels consume the BN-folded weights."""

    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)
        self.bn = nn.BatchNorm2d(cout)


class _ResBlock(nn.Module):
    """module.py:32-50 (holder)."""

    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.conv1 = _ConvBN(cin, cout, stride=stride)
        self.conv2 = _ConvBN(cout, cout)
        self.downsample = None if stride == 1 else _ConvBN(cin, cout, stride=stride)


class FeatureNet(nn.Module):
    """net.py:7-66 -- FPN feature extractor, eval-mode semantics (BatchNorm running statistics),
    on the tensor-core convolution kernels: all views of all reference views in one pass, BN / ReLU /
    residual / FPN adds fused into the convolutions, pyramids emitted channels-last."""

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
        self._ws = {}                 # workspace slot -> (shape key, buffer)
        self.workspace_slot = 0

    def _packed(self, device) -> _pack.PackedFeatureNet:
        def build():
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            return _pack.PackedFeatureNet(sd, device)
        # the cache key covers parameters and buffers (BatchNorm running statistics): estimator._pack_key
        return _cached_pack(self, device, build)

    def forward_nhwc(self, x: Tensor):
        """x [B,V,3,H,W] -> channels-last pyramids (fea1 [B,V,H/2,W/2,16], fea2 [...,32], fea3 [...,48])."""
        if self.training:
            raise NotImplementedError("itermvs_b200.FeatureNet.forward_nhwc is the inference kernel path (BatchNorm folded "
                                      "with running statistics); in train() mode use Pipeline.forward, which runs "
                                      "itermvs_b200/training.py (batch statistics, autograd)")
        # uint8 = the raw 8-bit image: normalised on the device exactly as the reference's loaders do (2 * x / 255. - 1, dtu_yao_eval.py:63-64)
        u8 = x.dtype == torch.uint8
        if u8:
            if not x.is_cuda:
                raise RuntimeError("itermvs_b200: imgs must be a CUDA tensor (there is no CPU path)")
            x = x.contiguous()
        else:
            x = ops._chk(x.float(), "imgs")
        b, v, c, h, w = x.shape
        assert c == 3
        dev = x.device
        n = b * v
        key = (n, h, w, str(dev))
        held = self._ws.get(self.workspace_slot)
        if held is None or held[0] != key:
            nbytes = _lib.lib().imvs_featurenet_workspace_bytes(n, h, w)
            if nbytes == 0:
                raise ValueError(f"FeatureNet: unsupported shape {tuple(x.shape)}")
            held = (key, torch.empty(nbytes, dtype=torch.uint8, device=dev))
            self._ws[self.workspace_slot] = held
        ws = held[1]
        f1 = torch.empty(b, v, h // 2, w // 2, 16, device=dev)
        f2 = torch.empty(b, v, h // 4, w // 4, 32, device=dev)
        f3 = torch.empty(b, v, h // 8, w // 8, 48, device=dev)
        fwd = _lib.lib().imvs_featurenet_forward_u8 if u8 else _lib.lib().imvs_featurenet_forward
        _lib.check(fwd(self._packed(dev).ref, x.data_ptr(), f1.data_ptr(), f2.data_ptr(), f3.data_ptr(), ws.data_ptr(), ws.numel(),
                       n, h, w, ops._stream()), "featurenet_forward")
        return f1, f2, f3

    def forward(self, x: Tensor):
        """Reference return format (net.py:35-66): dict level -> list of per-view NCHW [B,C,H,W]."""
        f1, f2, f3 = self.forward_nhwc(x)
        out = {}
        for name, f in (("level1", f1), ("level2", f2), ("level3", f3)):
            out[name] = [f[:, i].permute(0, 3, 1, 2).contiguous() for i in range(f.shape[1])]
        return out


class Pipeline(nn.Module):
    """net.py:68-128."""

    def __init__(self, iteration=4, test=False):
        super().__init__()
        self.feature_dim = [8, 16, 32, 48]
        self.hidden_dim = 32
        self.test = test
        self.feature_net = FeatureNet(test=test)
        self.iter_mvs = IterMVS(iteration, self.feature_dim[2], self.hidden_dim, test)
        self._last_nan_flag = None

    def set_workspace_slot(self, slot: int) -> None:
        """Workspaces are cached per slot; forwards that may overlap on the device (two CUDA graphs replayed on
        different streams) must run under different slots."""
        self.feature_net.workspace_slot = slot
        self.iter_mvs.workspace_slot = slot

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts checkpoints with the DataParallel 'module.' prefix (train.py:153-157) too."""
        if state_dict and all(k.startswith("module.") for k in state_dict):
            state_dict = {k[7:]: v for k, v in state_dict.items()}
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _forward_all_predictions(self, imgs, proj_matrices, depth_min, depth_max):
        """test=False output structure (net.py:100-120), forward only: what train.py's validation pass
        (train.py:250-262, model.eval() under no_grad) and full_loss consume."""
        features = self.feature_net(imgs["level_0"])                      # reference format: level -> list of views
        ref_feature = {k: v[0] for k, v in features.items()}
        src_features = {k: v[1:] for k, v in features.items()}
        ref_proj, src_projs = {}, {}
        for l in (1, 2, 3):
            pm = torch.unbind(proj_matrices[f"level_{l}"].float(), 1)
            ref_proj[f"level{l}"], src_projs[f"level{l}"] = pm[0], list(pm[1:])
        depths, depths_upsampled, confidences, confidence_upsampled = self.iter_mvs(
            ref_feature, src_features, ref_proj, src_projs, depth_min.float(), depth_max.float())
        return {"depths": depths, "depths_upsampled": depths_upsampled, "confidences": confidences,
                "confidence_upsampled": confidence_upsampled}

    def forward(self, imgs, proj_matrices, depth_min, depth_max):
        x = imgs["level_0"]
        if not x.is_cuda:
            raise RuntimeError("itermvs_b200.Pipeline: inputs must be CUDA tensors (there is no CPU path)")
        if not self.test and (self.training or _wants_grad(self)):
            # train.py:205 (model.train(); BatchNorm batch statistics; gradients): fused plane sweep with its CUDA
            # backward + the convolution stacks under torch autograd -- itermvs_b200/training.py
            from . import training
            return training.pipeline_train_forward(self, imgs, proj_matrices, depth_min, depth_max)
        if not self.test:
            return self._forward_all_predictions(imgs, proj_matrices, depth_min, depth_max)
        f1, f2, f3 = self.feature_net.forward_nhwc(x)
        projs = [ops._chk(proj_matrices[f"level_{l}"].float(), "proj_matrices") for l in (1, 2, 3)]   # net.py:96-98
        flag = ops.NanFlag(x.device)
        depth, depth_up, conf, conf_up = self.iter_mvs.forward_packed(
            f1, f2, f3, projs[0], projs[1], projs[2],
            ops._chk(depth_min.float(), "depth_min"), ops._chk(depth_max.float(), "depth_max"), nan_flag=flag)
        self._last_nan_flag = flag            # checked lazily by callers that synchronise (tests, eval loop)
        return {"depths_upsampled": depth_up, "confidence_upsampled": conf_up}


def _masked_l1(a: Tensor, b: Tensor, mask: Tensor) -> Tensor:
    return F.l1_loss(a[mask], b[mask], reduction="mean")


def full_loss(depths, depths_upsampled, confidences, depths_gt, mask, depth_min, depth_max, regress=True):
    """Training / validation loss of the reference (models/net.py:131-190), same signature and value.

    Host-side PyTorch (it is not on the hot path: a handful of reductions over quarter-resolution maps); runs on
    whatever device the predictions live on.  Terms, with N predictions (init + one per update) and 256 bins:
      * 0.8^N   * 256 * L1(normalized initial depth, normalized gt)                        [level_2 mask]
      * 0.8^(N-1-i) * cross-entropy of prediction i's 256-bin distribution against the one-hot gt bin
        (D = probability.size(1) bins over the clamped normalized gt; probabilities clamped at 1e-5)
      * if regress: 0.8^(N-1-i) * 256 * L1 on the pixels whose gt bin lies within +-4 of the arg-max bin,
        and 0.8^(N-1-i) * BCE-with-logits of the confidence logit against [|nd - nd_gt| < 0.002]
      * 256 * L1(normalized upsampled depth, normalized gt)                                [level_0 mask]
    """
    radius, out_num_samples = 4, 256
    probs = depths["probability"]
    num_sample = probs[0].size(1)
    m0, m2 = mask["level_0"] > 0.5, mask["level_2"] > 0.5
    gt0, gt2 = depths_gt["level_0"], depths_gt["level_2"]
    batch = gt2.size(0)
    inv_min = (1.0 / depth_min).view(batch, 1, 1, 1)
    inv_max = (1.0 / depth_max).view(batch, 1, 1, 1)
    nd_gt = ops.depth_normalization(gt2, inv_min, inv_max)
    gt_bin = torch.floor(torch.clamp(nd_gt, min=0, max=1) * (num_sample - 1) * m2.float()).long()
    onehot = torch.zeros_like(probs[0]).scatter_(1, gt_bin, 1)
    n_pred = len(depths["combine"])

    nd = ops.depth_normalization(depths["initial"][0], inv_min, inv_max)
    loss = 0.8 ** n_pred * out_num_samples * _masked_l1(nd, nd_gt, m2)
    bce = nn.BCEWithLogitsLoss()
    for i in range(n_pred):
        weight = 0.8 ** (n_pred - i - 1)
        p = torch.clamp(probs[i], min=1e-5)
        ce = -torch.sum(onehot * torch.log(p), dim=1, keepdim=True)
        loss = loss + weight * torch.mean(ce[m2])
        if not regress:
            continue
        with torch.no_grad():
            top = torch.argmax(p, dim=1, keepdim=True).float()
            near = (gt_bin >= top - radius) & (gt_bin <= top + radius)
        nd = ops.depth_normalization(depths["combine"][i], inv_min, inv_max)
        sel = m2 & near
        if torch.sum(sel) > 0:
            loss = loss + weight * out_num_samples * _masked_l1(nd, nd_gt, sel)
        target = (torch.abs(nd[m2].detach() - nd_gt[m2]) < 0.002).float()
        loss = loss + weight * bce(confidences[i][m2], target)

    nd_gt0 = ops.depth_normalization(gt0, inv_min, inv_max)
    nd_up = ops.depth_normalization(depths_upsampled[0], inv_min, inv_max)
    return loss + out_num_samples * _masked_l1(nd_up, nd_gt0, m0)
