"""Host-side mirror of reference models/itermvs.py + ConvGRU of models/module.py.

Same class names, constructor arguments, forward signatures, return values and state_dict keys as
the reference (DepthInitialization, Evaluation, Update, IterMVS, PixelViewWeight, CorrNet,
ConvGRU), so `Pipeline`, `train.py` / `eval.py` and the shipped checkpoints stay drop-in -- but
every forward runs the sm_100a kernels through the C ABI (no cuDNN / ATen call chains, no CPU path).

The nn.Conv2d / nn.ConvTranspose2d members are parameter holders only (names + shapes +
initialisation identical to the reference); their own forward is never called on the inference
path.  Parameters are re-packed to the kernel layout lazily whenever their version counters change.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from . import ops
from ._pack import PackedWeights

Tensor = torch.Tensor
_LEVEL_DIM = {"level1": 16, "level2": 32, "level3": 48}

# packed-weight caches live OUTSIDE the modules (ctypes structs must not be deep-copied / pickled
# with a module): module -> {"key": versions, "val": packed}
_PACK_CACHE: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _cache_of(mod) -> dict:
    c = _PACK_CACHE.get(mod)
    if c is None:
        c = {}
        _PACK_CACHE[mod] = c
    return c


# ------------------------------------------------------------------------------ holders -------
class _ConvHolder(nn.Module):
    """`<name>.conv.weight` key layout of the reference's ConvReLU wrapper (module.py:15-21)."""

    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)


def _stack_views(ref: Tensor, srcs: Sequence[Tensor]) -> Tensor:
    """reference-view + source-view NCHW tensors -> one channels-last pyramid [B,V,H,W,C]."""
    b, c, h, w = ref.shape
    v = 1 + len(srcs)
    out = torch.empty(b, v, h, w, c, device=ref.device, dtype=torch.float32)
    for i, t in enumerate([ref, *srcs]):
        if t.shape != ref.shape:
            raise ValueError("all views of a level must have the same shape")
        # out[:, i] is strided over the batch; transpose each batch element into place
        for bi in range(b):
            ops.nchw_to_nhwc(t[bi:bi + 1], out[bi, i:i + 1])
    return out


def _stack_proj(ref_proj: Tensor, src_projs: Sequence[Tensor]) -> Tensor:
    return torch.stack([ref_proj, *src_projs], dim=1).float().contiguous()


class _WeightOwner:
    """Mixin: lazily packed kernel weights for the parameters under `self._weight_root()`."""

    def _weight_root(self):            # the IterMVS-shaped module whose state_dict is packed
        raise NotImplementedError

    def packed(self, device) -> PackedWeights:
        root = self._weight_root()
        params = dict(root.named_parameters())
        key = (str(device),) + tuple((k, p._version, p.data_ptr()) for k, p in params.items())
        cache = _cache_of(root)
        if cache.get("key") != key:
            cache["key"] = key
            cache["val"] = PackedWeights({k: p.detach() for k, p in params.items()}, device)
        return cache["val"]


# --------------------------------------------------------------------------- sub-modules -------
class DepthInitialization(nn.Module):
    """itermvs.py:6-19: D inverse-depth-uniform hypotheses.  Inside IterMVS.forward the fused
    plane-sweep kernel generates them itself; this module exists for API parity."""

    def __init__(self, num_sample):
        super().__init__()
        self.num_sample = num_sample

    def forward(self, inverse_depth_min, inverse_depth_max, height, width, device):
        batch = inverse_depth_min.size(0)
        index = torch.arange(0, self.num_sample, 1, device=device, dtype=torch.float32).view(1, -1, 1, 1)
        normalized = index.repeat(batch, 1, height, width) / (self.num_sample - 1)
        return 1.0 / (inverse_depth_max + normalized * (inverse_depth_min - inverse_depth_max))


class PixelViewWeight(nn.Module):
    """itermvs.py:333-350.  x [B,G,N,H,W] -> [B,1,H,W]."""

    def __init__(self, G):
        super().__init__()
        self.conv = nn.Sequential(_ConvHolder(G, 16), nn.Conv2d(16, 1, 1, stride=1, padding=0))

    def forward(self, x: Tensor) -> Tensor:
        b, g, n, h, w = x.shape
        vol = ops._chk(x, "x").permute(0, 2, 3, 4, 1).contiguous().view(b, 1, n, h * w, g)
        wts = _StandaloneWeights.for_pvw(self, x.device)
        logits = torch.empty(b, n, h * w, device=x.device)
        vw3 = torch.empty(b, 1, h, w, device=x.device)
        vw2 = torch.empty(b, 1, 2 * h, 2 * w, device=x.device)
        _lib.check(_lib.lib().imvs_pixel_view_weight(C.byref(wts.struct), vol.data_ptr(), logits.data_ptr(), vw3.data_ptr(),
                                                     vw2.data_ptr(), b, 1, n, h, w, ops._stream()), "pixel_view_weight")
        return vw3


class CorrNet(nn.Module):
    """itermvs.py:352-381.  x [B,G,N,H,W] -> [B,N,H,W]."""

    def __init__(self, G):
        super().__init__()
        self.conv0 = _ConvHolder(G, 8)
        self.conv1 = _ConvHolder(8, 16, stride=2)
        self.conv2 = _ConvHolder(16, 32, stride=2)
        self.conv3 = nn.ConvTranspose2d(32, 16, 3, padding=1, output_padding=1, stride=2, bias=False)
        self.conv4 = nn.ConvTranspose2d(16, 8, 3, padding=1, output_padding=1, stride=2, bias=False)
        self.conv5 = nn.Conv2d(8, 1, 3, stride=1, padding=1)

    def forward(self, x: Tensor) -> Tensor:
        b, g, n, h, w = x.shape
        vol = ops._chk(x, "x").permute(0, 2, 3, 4, 1).contiguous()          # [B,N,H,W,8]
        wts = _StandaloneWeights.for_corrnet(self, x.device)
        out = torch.empty(b, n, h, w, device=x.device)
        scratch = torch.empty(_lib.lib().imvs_corrnet_scratch_floats(b * n, h, w), device=x.device)
        _lib.check(_lib.lib().imvs_corrnet(wts.sets, 1, 1, 1, vol.data_ptr(), out.data_ptr(), h * w, scratch.data_ptr(),
                                           b * n, h, w, ops._stream()), "corrnet")
        return out


class _StandaloneWeights:
    """Packed weights for a sub-module used on its own (tests / drop-in use of a single operator)."""

    def __init__(self):
        self.keep = []

    @staticmethod
    def _cached(mod, device, build):
        key = (str(device),) + tuple((k, p._version, p.data_ptr()) for k, p in mod.named_parameters())
        cache = _cache_of(mod)
        if cache.get("key") != key:
            cache["key"], cache["val"] = key, build()
        return cache["val"]

    @classmethod
    def for_corrnet(cls, mod: "CorrNet", device):
        from ._pack import pack_conv, pack_tconv, _vec

        def build():
            self = cls()
            t = {"conv0": pack_conv(mod.conv0.conv.weight.to(device)), "conv1": pack_conv(mod.conv1.conv.weight.to(device)),
                 "conv2": pack_conv(mod.conv2.conv.weight.to(device)), "conv3": pack_tconv(mod.conv3.weight.to(device)),
                 "conv4": pack_tconv(mod.conv4.weight.to(device)), "conv5": pack_conv(mod.conv5.weight.to(device)),
                 "conv5_b": _vec(mod.conv5.bias.to(device))}
            self.keep = t
            self.sets = (_lib.CorrNetWeights * 3)()
            for j in range(3):
                for k, v in t.items():
                    setattr(self.sets[j], k, v.data_ptr())
            return self
        return cls._cached(mod, device, build)

    @classmethod
    def for_pvw(cls, mod: "PixelViewWeight", device):
        from ._pack import pack_conv, _vec

        def build():
            self = cls()
            t = {"pvw_conv0": pack_conv(mod.conv[0].conv.weight.to(device)), "pvw_conv1": _vec(mod.conv[1].weight.to(device)),
                 "pvw_conv1_b": _vec(mod.conv[1].bias.to(device))}
            self.keep = t
            self.struct = _lib.Weights()
            for k, v in t.items():
                setattr(self.struct, k, v.data_ptr())
            return self
        return cls._cached(mod, device, build)

    @classmethod
    def for_update(cls, mod: "Update", device):
        from ._pack import pack_conv, _vec

        def build():
            self = cls()
            g = lambda p: p.to(device)
            t = {
                "gru_zr": pack_conv(torch.cat([g(mod.gru.convz.weight), g(mod.gru.convr.weight)], 0)),
                "gru_zr_b": _vec(torch.cat([g(mod.gru.convz.bias), g(mod.gru.convr.bias)], 0)),
                "gru_q": pack_conv(g(mod.gru.convq.weight)), "gru_q_b": _vec(g(mod.gru.convq.bias)),
                "head_conv0": pack_conv(torch.cat([g(mod.depth_head[0].weight), g(mod.confidence_head[0].weight)], 0)),
                "head_fc1": pack_conv(g(mod.depth_head[2].weight)), "head_fc2": pack_conv(g(mod.depth_head[4].weight)),
                "head_fc2_b": _vec(g(mod.depth_head[4].bias)), "conf_fc": _vec(g(mod.confidence_head[2].weight)),
                "conf_fc_b": _vec(g(mod.confidence_head[2].bias)),
                "hinit_conv0": pack_conv(g(mod.hidden_init_head[0].weight)),
                "hinit_fc": pack_conv(g(mod.hidden_init_head[2].weight)), "hinit_fc_b": _vec(g(mod.hidden_init_head[2].bias)),
            }
            self.keep = t
            self.struct = _lib.Weights()
            for k, v in t.items():
                setattr(self.struct, k, v.data_ptr())
            return self
        return cls._cached(mod, device, build)

    @classmethod
    def for_gru(cls, mod: "ConvGRU", device):
        from ._pack import pack_conv, _vec

        def build():
            self = cls()
            g = lambda p: p.to(device)
            t = {"gru_zr": pack_conv(torch.cat([g(mod.convz.weight), g(mod.convr.weight)], 0)),
                 "gru_zr_b": _vec(torch.cat([g(mod.convz.bias), g(mod.convr.bias)], 0)),
                 "gru_q": pack_conv(g(mod.convq.weight)), "gru_q_b": _vec(g(mod.convq.bias))}
            self.keep = t
            self.struct = _lib.Weights()
            for k, v in t.items():
                setattr(self.struct, k, v.data_ptr())
            return self
        return cls._cached(mod, device, build)


class ConvGRU(nn.Module):
    """module.py:52-66.  forward(h, x) -> new h.  The fused kernels are specialised for the
    estimator's shapes (hidden 32, input 11)."""

    def __init__(self, hidden_dim, input_dim, kernel_size=3):
        super().__init__()
        if (hidden_dim, input_dim, kernel_size) != (32, 11, 3):
            raise ValueError("itermvs_b200.ConvGRU is specialised for hidden_dim=32, input_dim=11, kernel_size=3 "
                             "(the only configuration the reference instantiates, itermvs.py:243)")
        c = hidden_dim + input_dim
        self.convz = nn.Conv2d(c, hidden_dim, kernel_size, padding=kernel_size - 1, dilation=2)
        self.convr = nn.Conv2d(c, hidden_dim, kernel_size, padding=kernel_size - 1, dilation=2)
        self.convq = nn.Conv2d(c, hidden_dim, kernel_size, padding=kernel_size - 1, dilation=2)

    def forward(self, h: Tensor, x: Tensor, _weights=None) -> Tensor:
        h = ops._chk(h, "h").clone()
        x = ops._chk(x, "x")
        b, _, hh, ww = h.shape
        wts = _weights if _weights is not None else _StandaloneWeights.for_gru(self, h.device).struct
        scratch = torch.empty(2 * h.numel(), device=h.device)
        _lib.check(_lib.lib().imvs_conv_gru(C.byref(wts), h.data_ptr(), x.data_ptr(), scratch.data_ptr(), b, hh, ww,
                                            ops._stream()), "conv_gru")
        return h


class Evaluation(nn.Module, _WeightOwner):
    """itermvs.py:22-126: correlation of all depth samples for each pixel."""

    def __init__(self):
        super().__init__()
        self.G = 8
        self.pixel_view_weight = PixelViewWeight(self.G)
        self.corr_conv1 = nn.ModuleList([CorrNet(self.G), CorrNet(self.G), CorrNet(self.G)])

    def _sets(self, device, idx):
        arr = (_lib.CorrNetWeights * 3)()
        keep = []
        for j, i in enumerate(idx):
            sw = _StandaloneWeights.for_corrnet(self.corr_conv1[i], device)
            keep.append(sw)
            for f, _t in _lib.CorrNetWeights._fields_:
                setattr(arr[j], f, getattr(sw.sets[0], f))
        return arr, keep

    def forward(self, ref_feature, src_features, ref_proj, src_projs, depth_sample, inverse_depth_min=None,
                inverse_depth_max=None, view_weights=None):
        L = _lib.lib()
        st = ops._stream()
        if view_weights is None:
            # ---- init branch (itermvs.py:36-82)
            ref3 = ops._chk(ref_feature["level3"], "ref_feature")
            dev = ref3.device
            b, _, h3, w3 = ref3.shape
            srcs = src_features["level3"]
            s = len(srcs)
            d = depth_sample.shape[1]
            fea3 = _stack_views(ref3, srcs)
            flag = ops.NanFlag(dev)
            rt3 = ops.compose_projections(_stack_proj(ref_proj["level3"], src_projs["level3"]).to(dev), flag)
            ds = ops._chk(depth_sample, "depth_sample")
            corr = torch.empty(b, s, d, h3 * w3, 8, device=dev)
            _lib.check(L.imvs_warpcorr_init(fea3.data_ptr(), rt3.data_ptr(), None, None, ds.data_ptr(), corr.data_ptr(),
                                            b, s + 1, h3, w3, d, st), "warpcorr_init")
            pw = _StandaloneWeights.for_pvw(self.pixel_view_weight, dev)
            logits = torch.empty(b, s, d, h3 * w3, device=dev)
            vw3 = torch.empty(b, s, h3, w3, device=dev)
            vw2 = torch.empty(b, s, 2 * h3, 2 * w3, device=dev)
            _lib.check(L.imvs_pixel_view_weight(C.byref(pw.struct), corr.data_ptr(), logits.data_ptr(), vw3.data_ptr(),
                                                vw2.data_ptr(), b, s, d, h3, w3, st), "pixel_view_weight")
            agg = torch.empty(b, d, h3 * w3, 8, device=dev)
            _lib.check(L.imvs_aggregate_init(corr.data_ptr(), vw3.data_ptr(), agg.data_ptr(), b, s, d, h3 * w3, st),
                       "aggregate_init")
            sets, _keep = self._sets(dev, (2, 2, 2))
            out = torch.empty(b, d, h3, w3, device=dev)
            scratch = torch.empty(L.imvs_corrnet_scratch_floats(b * d, h3, w3), device=dev)
            _lib.check(L.imvs_corrnet(sets, 1, 1, 1, agg.data_ptr(), out.data_ptr(), h3 * w3, scratch.data_ptr(), b * d, h3, w3,
                                      st), "corrnet")
            flag.raise_if_set()
            # itermvs.py:74-81 (only consumed by the training loss)
            probability = torch.softmax(out, dim=1)
            index = torch.arange(0, d, 1, device=dev, dtype=torch.float32).view(1, d, 1, 1)
            nd = torch.sum(index * probability, dim=1, keepdim=True) / (d - 1.0)
            depth = ops.depth_unnormalization(nd, inverse_depth_min, inverse_depth_max)
            depth = torch.nn.functional.interpolate(depth, scale_factor=2, mode="bilinear")
            return vw2, out, depth
        # ---- iteration branch (itermvs.py:84-126)
        ref2 = ops._chk(ref_feature["level2"], "ref_feature")
        dev = ref2.device
        b, _, h2, w2 = ref2.shape
        s = len(src_features["level2"])
        feas, rts = [], []
        flag = ops.NanFlag(dev)
        for l in (1, 2, 3):
            k = f"level{l}"
            feas.append(_stack_views(ops._chk(ref_feature[k], "ref_feature"), src_features[k]))
            rts.append(ops.compose_projections(_stack_proj(ref_proj[k], src_projs[k]).to(dev), flag))
        smp = [ops._chk(depth_sample[f"level{l}"], "depth_sample") for l in (1, 2, 3)]
        vw = ops._chk(view_weights, "view_weights")
        agg = torch.empty(b, 10, h2 * w2, 8, device=dev)
        _lib.check(L.imvs_warpcorr_iter(feas[0].data_ptr(), feas[1].data_ptr(), feas[2].data_ptr(), rts[0].data_ptr(),
                                        rts[1].data_ptr(), rts[2].data_ptr(), None, 0, vw.data_ptr(), None, None,
                                        smp[0].data_ptr(), smp[1].data_ptr(), smp[2].data_ptr(), agg.data_ptr(),
                                        b, s + 1, h2, w2, st), "warpcorr_iter")
        sets, _keep = self._sets(dev, (0, 1, 2))
        out = torch.empty(b, 10, h2, w2, device=dev)
        scratch = torch.empty(L.imvs_corrnet_scratch_floats(b * 10, h2, w2), device=dev)
        _lib.check(L.imvs_corrnet(sets, 10, 4, 8, agg.data_ptr(), out.data_ptr(), 10 * h2 * w2, scratch.data_ptr(), b * 10,
                                  h2, w2, st), "corrnet")
        flag.raise_if_set()
        return out


class Update(nn.Module):
    """itermvs.py:129-220."""

    def __init__(self, input_dim, hidden_dim, num_sample):
        super().__init__()
        self.G = 4
        self.hidden_dim = hidden_dim
        self.out_num_samples = 256
        self.radius = 4
        self.gru = ConvGRU(hidden_dim, input_dim)
        self.depth_head = nn.Sequential(
            nn.Conv2d(hidden_dim, 32, 3, stride=1, padding=2, dilation=2, bias=False), nn.ReLU(inplace=True),
            nn.Conv2d(32, 64, 1, stride=1, padding=0, dilation=1, bias=False), nn.ReLU(inplace=True),
            nn.Conv2d(64, self.out_num_samples, 1, stride=1, padding=0, dilation=1))
        self.confidence_head = nn.Sequential(
            nn.Conv2d(hidden_dim, 32, 3, stride=1, padding=2, dilation=2, bias=False), nn.ReLU(inplace=True),
            nn.Conv2d(32, 1, 1, stride=1, padding=0, dilation=1))
        self.hidden_init_head = nn.Sequential(
            nn.Conv2d(num_sample, 64, 3, stride=1, padding=1, dilation=1, bias=False), nn.ReLU(inplace=True),
            nn.Conv2d(64, hidden_dim, 1, stride=1, padding=0, dilation=1))
        self.return_probability: Optional[bool] = None    # None: materialise the 256-bin tensor only in training mode

    def _w(self, device):
        return _StandaloneWeights.for_update(self, device).struct

    def _want_prob(self):
        return self.training if self.return_probability is None else self.return_probability

    def hidden_init(self, corr: Tensor) -> Tensor:
        corr = ops._chk(corr, "corr")
        b, d, h3, w3 = corr.shape
        hidden = torch.empty(b, self.hidden_dim, 2 * h3, 2 * w3, device=corr.device)
        scratch = torch.empty(b * 96 * h3 * w3, device=corr.device)
        _lib.check(_lib.lib().imvs_hidden_init(C.byref(self._w(corr.device)), corr.data_ptr(), hidden.data_ptr(),
                                               scratch.data_ptr(), b, d, h3, w3, ops._stream()), "hidden_init")
        return hidden

    def _heads(self, hidden: Tensor, want_conf: bool):
        hidden = ops._chk(hidden, "hidden")
        b, _, h, w = hidden.shape
        dev = hidden.device
        nd = torch.empty(b, 1, h, w, device=dev)
        prob = torch.empty(b, self.out_num_samples, h, w, device=dev) if self._want_prob() else None
        conf = torch.empty(b, 1, h, w, device=dev) if want_conf else None
        conf0 = torch.empty(b, 1, h, w, device=dev) if want_conf else None
        scratch = torch.empty(b * 64 * h * w, device=dev)
        _lib.check(_lib.lib().imvs_depth_head(C.byref(self._w(dev)), hidden.data_ptr(), nd.data_ptr(), h * w, ops._p(prob),
                                              ops._p(conf), ops._p(conf0), None, None, None, scratch.data_ptr(), b, h, w,
                                              ops._stream()), "depth_head")
        return nd, prob, conf, conf0

    def conf_init(self, hidden):
        _, _, conf, conf0 = self._heads(hidden, True)
        return conf, conf0

    def depth_init(self, hidden):
        nd, prob, _, _ = self._heads(hidden, False)
        return nd, prob

    def forward(self, hidden, normalized_depth, corr, confidence=None, confidence_flag=False):
        x = torch.cat([normalized_depth, corr], dim=1)
        hidden = self.gru(hidden, x, _weights=self._w(hidden.device))
        nd, prob, conf, conf0 = self._heads(hidden, confidence_flag)
        return hidden, nd, prob, conf, conf0


# -------------------------------------------------------------------------------- IterMVS ------
class IterMVS(nn.Module, _WeightOwner):
    """itermvs.py:223-329."""

    def __init__(self, iteration, feature_dim, hidden_dim, test=False):
        super().__init__()
        self.iteration = iteration
        self.hidden_dim = hidden_dim
        self.corr_sample = 10
        self.interval_scale = 1.0 / 256
        self.corr_interval = {
            "level1": torch.FloatTensor([-2, -2.0 / 3, 2.0 / 3, 2]).view(1, 4, 1, 1),
            "level2": torch.FloatTensor([-8, -8.0 / 3, 8.0 / 3, 8]).view(1, 4, 1, 1),
            "level3": torch.FloatTensor([-32, 32]).view(1, 2, 1, 1),
        }
        self.num_sample = 32
        self.test = test
        self.depth_initialization = DepthInitialization(self.num_sample)
        self.evaluation = Evaluation()
        self.update = Update(1 + 10, hidden_dim, self.num_sample)
        self.upsample = nn.Sequential(
            nn.Conv2d(feature_dim, 64, 3, stride=1, padding=1, dilation=1, bias=False), nn.ReLU(inplace=True),
            nn.Conv2d(64, 16 * 9, 1, stride=1, padding=0, dilation=1, bias=False))
        self._workspaces: Dict[tuple, Tensor] = {}

    def _weight_root(self):
        return self

    def _workspace(self, pb: _lib.Problem, device) -> Tensor:
        key = (pb.B, pb.V, pb.H, pb.W, pb.D, pb.iterations, str(device))
        ws = self._workspaces.get(key)
        if ws is None:
            nbytes = _lib.lib().imvs_forward_workspace_bytes(C.byref(pb))
            if nbytes == 0:
                _lib.check(1, "forward_workspace_bytes")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._workspaces = {key: ws}            # one live workspace per module
        return ws

    def forward_packed(self, fea1: Tensor, fea2: Tensor, fea3: Tensor, ref_fea2_planar: Tensor, proj1: Tensor,
                       proj2: Tensor, proj3: Tensor, depth_min: Tensor, depth_max: Tensor, out=None, nan_flag=None):
        """Test-mode forward on channels-last pyramids [B,V,H_l,W_l,C_l] (what FeatureNet emits).
        One C call enqueues the whole estimator.  Returns (depth, depth_up, conf, conf_up)."""
        b, v, h1, w1, _ = fea1.shape
        dev = fea1.device
        d = int(self.update.hidden_init_head[0].weight.shape[1])
        pb = _lib.Problem(b, v, 2 * h1, 2 * w1, d, self.iteration)
        ws = self._workspace(pb, dev)
        H, W = 2 * h1, 2 * w1
        if out is None:
            out = (torch.empty(b, 1, H // 4, W // 4, device=dev), torch.empty(b, 1, H, W, device=dev),
                   torch.empty(b, 1, H // 4, W // 4, device=dev), torch.empty(b, 1, H, W, device=dev))
        wts = self.packed(dev)
        _lib.check(_lib.lib().imvs_itermvs_forward(
            C.byref(pb), wts.ref, fea1.data_ptr(), fea2.data_ptr(), fea3.data_ptr(), ref_fea2_planar.data_ptr(),
            proj1.data_ptr(), proj2.data_ptr(), proj3.data_ptr(), depth_min.data_ptr(), depth_max.data_ptr(),
            ws.data_ptr(), ws.numel(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(),
            nan_flag.ptr() if nan_flag is not None else None, ops._stream()), "itermvs_forward")
        return out

    def forward(self, ref_feature, src_features, ref_proj, src_projs, depth_min, depth_max):
        if not self.test:
            raise NotImplementedError(
                "itermvs_b200.IterMVS: the training-mode forward (test=False) is not built in this round; "
                "see DESIGN.md ('out of scope / next').  Inference (test=True) is the supported path.")
        dev = ref_feature["level2"].device
        feas = [_stack_views(ops._chk(ref_feature[f"level{l}"], "ref_feature"), src_features[f"level{l}"]) for l in (1, 2, 3)]
        projs = [_stack_proj(ref_proj[f"level{l}"], src_projs[f"level{l}"]).to(dev) for l in (1, 2, 3)]
        flag = ops.NanFlag(dev)
        out = self.forward_packed(feas[0], feas[1], feas[2], ops._chk(ref_feature["level2"], "ref_feature"),
                                  projs[0], projs[1], projs[2], ops._chk(depth_min.float(), "depth_min"),
                                  ops._chk(depth_max.float(), "depth_max"), nan_flag=flag)
        flag.raise_if_set()
        return out
