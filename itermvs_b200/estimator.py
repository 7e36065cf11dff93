"""Host-side mirror of reference models/itermvs.py + ConvGRU of models/module.py.

Same class names, constructor arguments, forward signatures, return values and state_dict keys as
the reference (DepthInitialization, Evaluation, Update, IterMVS, PixelViewWeight, CorrNet,
ConvGRU), so `Pipeline`, `train.py` / `eval.py` and the shipped checkpoints stay drop-in -- but
every forward runs the sm_100a kernels through the C ABI (no cuDNN / ATen call chains, no CPU path).

The nn.Conv2d / nn.ConvTranspose2d members hold the parameters (names + shapes + initialisation identical
to the reference); their own forward is never called on the inference path.  In train() mode with autograd on, every
module here runs its differentiable implementation instead (itermvs_b200/training.py).  Parameters are re-packed to the kernel layout lazily whenever their version counters change.
The single-operator modules accept and return the reference's NCHW tensors (converted at the
boundary); `IterMVS.forward_packed` is the zero-copy channels-last path `Pipeline` uses.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from . import _pack
from . import ops

Tensor = torch.Tensor
XCH = 16          # stored channels of the GRU input x (IMVS_XCH)

# packed-weight caches live OUTSIDE the modules (ctypes structs must not be deep-copied / pickled
# with a module): module -> {"key": versions, "val": packed}
_PACK_CACHE: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _pack_key(mod: nn.Module, device):
    """Version / storage of every parameter AND buffer (BatchNorm running statistics are buffers): a change of any of
    them -- optimizer step, load_state_dict, BN re-calibration, a broadcast -- invalidates the packed copy."""
    named = list(mod.named_parameters()) + list(mod.named_buffers())
    return (str(device),) + tuple((k, t._version, t.data_ptr()) for k, t in named)


def invalidate_packs(root: nn.Module = None) -> None:
    """Drop cached packed weights (all, or those of `root`'s submodules).  Needed only after writes that bypass
    autograd's version counter (`tensor.data` assignment, raw pointer writes)."""
    if root is None:
        _PACK_CACHE.clear()
        return
    for m in root.modules():
        _PACK_CACHE.pop(m, None)


def _cached_pack(mod: nn.Module, device, build):
    key = _pack_key(mod, device)
    cache = _PACK_CACHE.get(mod)
    if cache is None:
        cache = {}
        _PACK_CACHE[mod] = cache
    if cache.get("key") != key:
        cache["key"], cache["val"] = key, build()
    return cache["val"]


def _sd(mod: nn.Module) -> Dict[str, Tensor]:
    return {k: p.detach() for k, p in mod.named_parameters()}


def _nhwc(x: Tensor) -> Tensor:
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x: Tensor) -> Tensor:
    return x.permute(0, 3, 1, 2).contiguous()


def _wants_grad(mod: nn.Module) -> bool:
    return torch.is_grad_enabled() and any(p.requires_grad for p in mod.parameters())


def _differentiable(mod: nn.Module) -> bool:
    """train() mode with autograd on: the module runs its differentiable implementation (itermvs_b200/training.py) --
    torch autograd over the parameter-holding convolutions, the fused plane sweep with its CUDA backward; eval() or
    no_grad: the inference kernels."""
    return mod.training and torch.is_grad_enabled()


class _ConvHolder(nn.Module):
    """`<name>.conv.weight` key layout of the reference's ConvReLU wrapper (module.py:15-21)."""

    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)


def _stack_views(ref: Tensor, srcs: Sequence[Tensor]) -> Tensor:
    """reference-view + source-view NCHW tensors -> one channels-last pyramid [B,V,H,W,C]."""
    for t in srcs:
        if t.shape != ref.shape:
            raise ValueError("all views of a level must have the same shape")
    return torch.stack([ops._chk(t, "feature").permute(0, 2, 3, 1) for t in (ref, *srcs)], dim=1).contiguous()


def _stack_proj(ref_proj: Tensor, src_projs: Sequence[Tensor]) -> Tensor:
    return torch.stack([ref_proj, *src_projs], dim=1).float().contiguous()


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


class _Packed(_pack._Holder):
    def __init__(self):
        super().__init__()
        self.struct = _lib.Weights()

    @property
    def ref(self):
        return C.byref(self.struct)


class PixelViewWeight(nn.Module):
    """itermvs.py:333-350.  x [B,G,N,H,W] -> [B,1,H,W]."""

    def __init__(self, G):
        super().__init__()
        self.conv = nn.Sequential(_ConvHolder(G, 16), nn.Conv2d(16, 1, 1, stride=1, padding=0))

    def _packed(self, device):
        def build():
            p = _Packed()
            _pack.fill_pvw(p, p.struct, _sd(self), "", device)
            return p
        return _cached_pack(self, device, build)

    def forward(self, x: Tensor) -> Tensor:
        if _differentiable(self):
            from . import training
            return training.pixel_view_weight(self, ops._chk(x, "x"))
        b, g, n, h, w = x.shape
        vol = ops._chk(x, "x").permute(0, 2, 3, 4, 1).contiguous()          # [B,N,H,W,8] == [B][S=1][D][P][8]
        logits = torch.empty(b, n, h * w, device=x.device)
        vw3 = torch.empty(b, 1, h, w, device=x.device)
        vw2 = torch.empty(b, 1, 2 * h, 2 * w, device=x.device)
        _lib.check(_lib.lib().imvs_pixel_view_weight(self._packed(x.device).ref, vol.data_ptr(), logits.data_ptr(), vw3.data_ptr(),
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

    def _packed(self, device):
        def build():
            h = _pack._Holder()
            h.sets = (_lib.CorrNetWeights * 3)()
            _pack.fill_corrnet(h, h.sets[0], _sd(self), "", device)
            h.sets[1] = h.sets[0]
            h.sets[2] = h.sets[0]
            return h
        return _cached_pack(self, device, build)

    def forward(self, x: Tensor) -> Tensor:
        if _differentiable(self):
            from . import training
            return training.corr_net(self, ops._chk(x, "x"))
        b, g, n, h, w = x.shape
        vol = ops._chk(x, "x").permute(0, 2, 3, 4, 1).contiguous()          # [B*N][P][8]
        out = torch.empty(b * n, h, w, device=x.device)
        scratch = torch.empty(_lib.lib().imvs_corrnet_scratch_floats(b * n, h, w), device=x.device)
        _lib.check(_lib.lib().imvs_corrnet(self._packed(x.device).sets, 1, 1, 1, vol.data_ptr(), out.data_ptr(), h * w, 1,
                                           scratch.data_ptr(), b * n, h, w, ops._stream()), "corrnet")
        return out.view(b, n, h, w)


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

    def _packed(self, device):
        def build():
            p = _Packed()
            _pack.fill_gru(p, p.struct, _sd(self), "", device)
            return p
        return _cached_pack(self, device, build)

    def forward_nhwc(self, h: Tensor, x16: Tensor, wref=None) -> Tensor:
        """h [B,H,W,32] (updated IN PLACE and returned), x16 [B,H,W,16]."""
        b, hh, ww, _ = h.shape
        scratch = torch.empty(4 * h.numel(), device=h.device)
        _lib.check(_lib.lib().imvs_conv_gru(wref if wref is not None else self._packed(h.device).ref, h.data_ptr(), x16.data_ptr(),
                                            scratch.data_ptr(), b, hh, ww, ops._stream()), "conv_gru")
        return h

    def forward(self, h: Tensor, x: Tensor) -> Tensor:
        if _differentiable(self):
            from . import training
            return training.conv_gru(self, ops._chk(h, "h"), ops._chk(x, "x"))
        hn = _nhwc(ops._chk(h, "h"))
        x = ops._chk(x, "x")
        b, c, hh, ww = x.shape
        x16 = torch.zeros(b, hh, ww, XCH, device=x.device)
        x16[..., :c] = x.permute(0, 2, 3, 1)
        return _nchw(self.forward_nhwc(hn, x16))


class Evaluation(nn.Module):
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
            pk = self.corr_conv1[i]._packed(device)
            keep.append(pk)
            arr[j] = pk.sets[0]
        return arr, keep

    def forward(self, ref_feature, src_features, ref_proj, src_projs, depth_sample, inverse_depth_min=None,
                inverse_depth_max=None, view_weights=None):
        if _differentiable(self):
            from . import training
            return training.evaluation_forward(self, ref_feature, src_features, ref_proj, src_projs, depth_sample,
                                               inverse_depth_min, inverse_depth_max, view_weights)
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
            logits = torch.empty(b, s, d, h3 * w3, device=dev)
            vw3 = torch.empty(b, s, h3, w3, device=dev)
            vw2 = torch.empty(b, s, 2 * h3, 2 * w3, device=dev)
            _lib.check(L.imvs_pixel_view_weight(self.pixel_view_weight._packed(dev).ref, corr.data_ptr(), logits.data_ptr(),
                                                vw3.data_ptr(), vw2.data_ptr(), b, s, d, h3, w3, st), "pixel_view_weight")
            agg = torch.empty(b, d, h3 * w3, 8, device=dev)
            _lib.check(L.imvs_aggregate_init(corr.data_ptr(), vw3.data_ptr(), agg.data_ptr(), b, s, d, h3 * w3, st),
                       "aggregate_init")
            sets, _keep = self._sets(dev, (2, 2, 2))
            out = torch.empty(b, d, h3, w3, device=dev)          # planar: pixel stride 1, slice stride H*W
            scratch = torch.empty(L.imvs_corrnet_scratch_floats(b * d, h3, w3), device=dev)
            _lib.check(L.imvs_corrnet(sets, 1, 1, 1, agg.data_ptr(), out.data_ptr(), h3 * w3, 1, scratch.data_ptr(), b * d, h3, w3,
                                      st), "corrnet")
            flag.raise_if_set()
            # itermvs.py:74-81 (only consumed by the training loss): softmax expectation -> depth -> bilinear x2, one C call
            dmin = (1.0 / inverse_depth_min.reshape(b).float()).contiguous()
            dmax = (1.0 / inverse_depth_max.reshape(b).float()).contiguous()
            depth = torch.empty(b, 1, 2 * h3, 2 * w3, device=dev)
            scr = torch.empty(b * h3 * w3, device=dev)
            _lib.check(L.imvs_init_depth(out.data_ptr(), d * h3 * w3, h3 * w3, 1, dmin.data_ptr(), dmax.data_ptr(), scr.data_ptr(),
                                         depth.data_ptr(), b, d, h3, w3, st), "init_depth")
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
                                        rts[1].data_ptr(), rts[2].data_ptr(), None, 0, 1, vw.data_ptr(), None, None,
                                        smp[0].data_ptr(), smp[1].data_ptr(), smp[2].data_ptr(), agg.data_ptr(),
                                        b, s + 1, h2, w2, st), "warpcorr_iter")
        sets, _keep = self._sets(dev, (0, 1, 2))
        out = torch.empty(b, h2, w2, 10, device=dev)             # channels-last: pixel stride 10
        scratch = torch.empty(L.imvs_corrnet_scratch_floats(b * 10, h2, w2), device=dev)
        _lib.check(L.imvs_corrnet(sets, 10, 4, 8, agg.data_ptr(), out.data_ptr(), 10 * h2 * w2, 10, scratch.data_ptr(), b * 10,
                                  h2, w2, st), "corrnet")
        flag.raise_if_set()
        return _nchw(out)


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

    def _packed(self, device):
        def build():
            p = _Packed()
            _pack.fill_update(p, p.struct, _sd(self), "", device)
            return p
        return _cached_pack(self, device, build)

    def _want_prob(self):
        return self.training if self.return_probability is None else self.return_probability

    def hidden_init(self, corr: Tensor) -> Tensor:
        if _differentiable(self):
            from . import training
            return training.update_hidden_init(self, ops._chk(corr, "corr"))
        corr = _nhwc(ops._chk(corr, "corr"))
        b, h3, w3, d = corr.shape
        hidden = torch.empty(b, 2 * h3, 2 * w3, self.hidden_dim, device=corr.device)
        scratch = torch.empty(b * 96 * h3 * w3, device=corr.device)
        _lib.check(_lib.lib().imvs_hidden_init(self._packed(corr.device).ref, corr.data_ptr(), hidden.data_ptr(),
                                               scratch.data_ptr(), b, d, h3, w3, ops._stream()), "hidden_init")
        return _nchw(hidden)

    def _heads_nhwc(self, hidden: Tensor, want_conf: bool):
        b, h, w, _ = hidden.shape
        dev = hidden.device
        nd = torch.empty(b, 1, h, w, device=dev)
        prob = torch.empty(b, self.out_num_samples, h, w, device=dev) if self._want_prob() else None
        conf = torch.empty(b, 1, h, w, device=dev) if want_conf else None
        conf0 = torch.empty(b, 1, h, w, device=dev) if want_conf else None
        scratch = torch.empty(b * 384 * h * w, device=dev)
        _lib.check(_lib.lib().imvs_depth_head(self._packed(dev).ref, hidden.data_ptr(), nd.data_ptr(), h * w, 1, ops._p(prob),
                                              ops._p(conf), ops._p(conf0), None, None, None, scratch.data_ptr(), b, h, w,
                                              ops._stream()), "depth_head")
        return nd, prob, conf, conf0

    def conf_init(self, hidden):
        if _differentiable(self):
            conf0 = self.confidence_head(ops._chk(hidden, "hidden"))
            return torch.sigmoid(conf0), conf0
        _, _, conf, conf0 = self._heads_nhwc(_nhwc(ops._chk(hidden, "hidden")), True)
        return conf, conf0

    def depth_init(self, hidden):
        if _differentiable(self):
            from . import training
            return training.update_depth(self, ops._chk(hidden, "hidden"))
        nd, prob, _, _ = self._heads_nhwc(_nhwc(ops._chk(hidden, "hidden")), False)
        return nd, prob

    def forward(self, hidden, normalized_depth, corr, confidence=None, confidence_flag=False):
        if _differentiable(self):
            from . import training
            return training.update_forward(self, ops._chk(hidden, "hidden"), normalized_depth, corr, confidence_flag)
        x = torch.cat([normalized_depth, corr], dim=1)
        b, c, hh, ww = x.shape
        x16 = torch.zeros(b, hh, ww, XCH, device=x.device)
        x16[..., :c] = ops._chk(x, "x").permute(0, 2, 3, 1)
        hn = self.gru.forward_nhwc(_nhwc(ops._chk(hidden, "hidden")), x16, wref=self._packed(x.device).ref)
        nd, prob, conf, conf0 = self._heads_nhwc(hn, confidence_flag)
        return _nchw(hn), nd, prob, conf, conf0


# -------------------------------------------------------------------------------- IterMVS ------
class IterMVS(nn.Module):
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
        self._workspaces: Dict[int, tuple] = {}      # workspace slot -> (shape key, buffer): one live buffer per slot
        self.workspace_slot = 0                        # concurrent replays (graph.StreamingPipeline) use one slot each

    def packed(self, device) -> _pack.PackedWeights:
        return _cached_pack(self, device, lambda: _pack.PackedWeights(_sd(self), device))

    def _workspace(self, pb: _lib.Problem, device) -> Tensor:
        key = (pb.B, pb.V, pb.H, pb.W, pb.D, pb.iterations, str(device))
        held = self._workspaces.get(self.workspace_slot)
        if held is None or held[0] != key:
            nbytes = _lib.lib().imvs_forward_workspace_bytes(C.byref(pb))
            if nbytes == 0:
                _lib.check(1, "forward_workspace_bytes")
            held = (key, torch.empty(nbytes, dtype=torch.uint8, device=device))
            self._workspaces[self.workspace_slot] = held
        return held[1]

    def forward_packed(self, fea1: Tensor, fea2: Tensor, fea3: Tensor, proj1: Tensor, proj2: Tensor, proj3: Tensor,
                       depth_min: Tensor, depth_max: Tensor, out=None, nan_flag=None):
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
            C.byref(pb), wts.ref, fea1.data_ptr(), fea2.data_ptr(), fea3.data_ptr(),
            proj1.data_ptr(), proj2.data_ptr(), proj3.data_ptr(), depth_min.data_ptr(), depth_max.data_ptr(),
            ws.data_ptr(), ws.numel(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(),
            nan_flag.ptr() if nan_flag is not None else None, ops._stream()), "itermvs_forward")
        return out

    def _upsample_outputs(self, ref2: Tensor, nd: Tensor, conf: Optional[Tensor], depth_min: Tensor, depth_max: Tensor):
        """itermvs.py:262-264 + 310-314 / 321-324 in one call: weight net on the level-2 reference feature, softmax
        over the 9 taps, convex x4 upsampling of `nd`, depth_unnormalization, bilinear x4 of the confidence."""
        b, _, h, w = ref2.shape
        dev = ref2.device
        fea = _nhwc(ref2)
        depth_up = torch.empty(b, 1, 4 * h, 4 * w, device=dev)
        conf_up = torch.empty(b, 1, 4 * h, 4 * w, device=dev) if conf is not None else None
        scratch = torch.empty(b * 64 * h * w, device=dev)
        nd = ops._chk(nd, "normalized_depth")
        _lib.check(_lib.lib().imvs_upsample_outputs(self.packed(dev).ref, fea.data_ptr(), h * w * fea.shape[-1], nd.data_ptr(), h * w, 1,
                                                    ops._p(conf), depth_min.data_ptr(), depth_max.data_ptr(), depth_up.data_ptr(),
                                                    ops._p(conf_up), scratch.data_ptr(), b, h, w, ops._stream()), "upsample_outputs")
        return depth_up, conf_up

    def _forward_all_predictions(self, ref_feature, src_features, ref_proj, src_projs, depth_min, depth_max):
        """The test=False structure of itermvs.py:253-329 -- every intermediate prediction (initial depth, one
        depth / probability / confidence logit per update) -- as a FORWARD pass on the CUDA operators.  This is what
        train.py's validation loop (train.py:257, model.eval() under no_grad) and full_loss consume.  The inference
        kernels carry no gradient; the differentiable path is the train() mode (itermvs_b200/training.py)."""
        depths = {"combine": [], "probability": [], "initial": []}
        confidences, depths_upsampled, confidence_upsampled = [], [], None
        ref2 = ops._chk(ref_feature["level2"], "ref_feature")
        dev = ref2.device
        batch, _, height, width = ref2.shape
        dmin, dmax = ops._chk(depth_min.float(), "depth_min"), ops._chk(depth_max.float(), "depth_max")
        inv_min, inv_max = (1.0 / dmin).view(batch, 1, 1, 1), (1.0 / dmax).view(batch, 1, 1, 1)
        upd = self.update
        saved, upd.return_probability = upd.return_probability, True
        try:
            samples0 = self.depth_initialization(inv_min, inv_max, height // 2, width // 2, dev)
            view_weights, corr, depth = self.evaluation(ref_feature, src_features, ref_proj, src_projs, samples0, inv_min, inv_max)
            depths["initial"].append(depth)
            hidden = upd.hidden_init(corr)
            nd, prob, conf, conf0 = upd._heads_nhwc(_nhwc(hidden), True)          # depth_init + conf_init (itermvs.py:276-279)
            depths["combine"].append(ops.depth_unnormalization(nd, inv_min, inv_max))
            depths["probability"].append(prob)
            confidences.append(conf0)
            for it in range(self.iteration):
                samples = {}
                for lvl in ("level1", "level2", "level3"):                        # itermvs.py:289-293
                    ns = torch.clamp(nd + self.corr_interval[lvl].to(dev) * self.interval_scale, min=0, max=1)
                    samples[lvl] = ops.depth_unnormalization(ns, inv_min, inv_max)
                corr = self.evaluation(ref_feature, src_features, ref_proj, src_projs, samples, view_weights=view_weights)
                hidden, nd, prob, conf, conf0 = upd(hidden, nd, corr, confidence_flag=True)
                depths["combine"].append(ops.depth_unnormalization(nd, inv_min, inv_max))
                depths["probability"].append(prob)
                confidences.append(conf0)
                if it == self.iteration - 1:
                    depth_up, confidence_upsampled = self._upsample_outputs(ref2, nd, conf, dmin, dmax)
                    depths_upsampled.append(depth_up)
        finally:
            upd.return_probability = saved
        return depths, depths_upsampled, confidences, confidence_upsampled

    def forward(self, ref_feature, src_features, ref_proj, src_projs, depth_min, depth_max):
        # differentiable path: train() mode, or eval() with autograd on (fine-tuning with frozen BatchNorm statistics works in
        # the reference; the modules of training.py follow `m.training`, so eval() uses running statistics there too)
        if not self.test and (self.training or _wants_grad(self)):
            # itermvs.py:253-329 as train.py drives it: fused plane sweep with CUDA backward + torch autograd (training.py)
            from . import training
            levels = ("level1", "level2", "level3")
            feas = [torch.stack([ref_feature[k], *src_features[k]], dim=1).permute(0, 1, 3, 4, 2).contiguous() for k in levels]
            projs = [_stack_proj(ref_proj[k], src_projs[k]).to(feas[0].device) for k in levels]
            return training.itermvs_train_forward(self, feas[0], feas[1], feas[2], projs, depth_min.float(), depth_max.float())
        if not self.test:
            return self._forward_all_predictions(ref_feature, src_features, ref_proj, src_projs, depth_min, depth_max)
        dev = ref_feature["level2"].device
        feas = [_stack_views(ops._chk(ref_feature[f"level{l}"], "ref_feature"), src_features[f"level{l}"]) for l in (1, 2, 3)]
        projs = [_stack_proj(ref_proj[f"level{l}"], src_projs[f"level{l}"]).to(dev) for l in (1, 2, 3)]
        flag = ops.NanFlag(dev)
        out = self.forward_packed(feas[0], feas[1], feas[2], projs[0], projs[1], projs[2],
                                  ops._chk(depth_min.float(), "depth_min"), ops._chk(depth_max.float(), "depth_max"), nan_flag=flag)
        flag.raise_if_set()
        return out
