"""Pack the estimator's parameters (reference state_dict names) into the kernel layouts of
include/itermvs_b200.h: conv weights [Cin][k*k][Cout], transposed convs [Cin][9][Cout].

The packed tensors are plain device tensors owned by a `PackedWeights` object; `struct` is the
`imvs_weights` C struct pointing at them.  Repack whenever parameters change (training) -- the
modules in estimator.py key the cache on the parameters' `_version` counters.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib

Tensor = torch.Tensor


def pack_conv(w: Tensor) -> Tensor:
    """nn.Conv2d weight [Cout,Cin,k,k] -> [Cin][k*k][Cout]."""
    co, cin, kh, kw = w.shape
    return w.detach().float().permute(1, 2, 3, 0).reshape(cin, kh * kw, co).contiguous()


def pack_tconv(w: Tensor) -> Tensor:
    """nn.ConvTranspose2d weight [Cin,Cout,k,k] -> [Cin][k*k][Cout]."""
    cin, co, kh, kw = w.shape
    return w.detach().float().permute(0, 2, 3, 1).reshape(cin, kh * kw, co).contiguous()


def _vec(t: Tensor) -> Tensor:
    return t.detach().float().reshape(-1).contiguous()


_CORR_FIELDS = ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv5_b")


class PackedWeights:
    """sd: {name: tensor} with names relative to the IterMVS module
    ('evaluation.pixel_view_weight.conv.0.conv.weight', 'update.gru.convz.weight', ...)."""

    def __init__(self, sd: Dict[str, Tensor], device: torch.device):
        g = lambda k: sd[k].to(device)
        keep = {}
        ev, up = "evaluation.", "update."
        keep["pvw_conv0"] = pack_conv(g(ev + "pixel_view_weight.conv.0.conv.weight"))
        keep["pvw_conv1"] = _vec(g(ev + "pixel_view_weight.conv.1.weight"))
        keep["pvw_conv1_b"] = _vec(g(ev + "pixel_view_weight.conv.1.bias"))
        for i in range(3):
            p = f"{ev}corr_conv1.{i}."
            keep[f"c{i}.conv0"] = pack_conv(g(p + "conv0.conv.weight"))
            keep[f"c{i}.conv1"] = pack_conv(g(p + "conv1.conv.weight"))
            keep[f"c{i}.conv2"] = pack_conv(g(p + "conv2.conv.weight"))
            keep[f"c{i}.conv3"] = pack_tconv(g(p + "conv3.weight"))
            keep[f"c{i}.conv4"] = pack_tconv(g(p + "conv4.weight"))
            keep[f"c{i}.conv5"] = pack_conv(g(p + "conv5.weight"))
            keep[f"c{i}.conv5_b"] = _vec(g(p + "conv5.bias"))
        keep["gru_zr"] = pack_conv(torch.cat([g(up + "gru.convz.weight"), g(up + "gru.convr.weight")], 0))
        keep["gru_zr_b"] = _vec(torch.cat([g(up + "gru.convz.bias"), g(up + "gru.convr.bias")], 0))
        keep["gru_q"] = pack_conv(g(up + "gru.convq.weight"))
        keep["gru_q_b"] = _vec(g(up + "gru.convq.bias"))
        keep["head_conv0"] = pack_conv(torch.cat([g(up + "depth_head.0.weight"), g(up + "confidence_head.0.weight")], 0))
        keep["head_fc1"] = pack_conv(g(up + "depth_head.2.weight"))
        keep["head_fc2"] = pack_conv(g(up + "depth_head.4.weight"))
        keep["head_fc2_b"] = _vec(g(up + "depth_head.4.bias"))
        keep["conf_fc"] = _vec(g(up + "confidence_head.2.weight"))
        keep["conf_fc_b"] = _vec(g(up + "confidence_head.2.bias"))
        keep["hinit_conv0"] = pack_conv(g(up + "hidden_init_head.0.weight"))
        keep["hinit_fc"] = pack_conv(g(up + "hidden_init_head.2.weight"))
        keep["hinit_fc_b"] = _vec(g(up + "hidden_init_head.2.bias"))
        keep["ups_conv0"] = pack_conv(g("upsample.0.weight"))
        keep["ups_fc"] = pack_conv(g("upsample.2.weight"))
        self.tensors = keep
        self.num_sample = int(sd[up + "hidden_init_head.0.weight"].shape[1])
        s = _lib.Weights()
        for name in ("pvw_conv0", "pvw_conv1", "pvw_conv1_b", "gru_zr", "gru_zr_b", "gru_q", "gru_q_b", "head_conv0",
                     "head_fc1", "head_fc2", "head_fc2_b", "conf_fc", "conf_fc_b", "hinit_conv0", "hinit_fc",
                     "hinit_fc_b", "ups_conv0", "ups_fc"):
            setattr(s, name, keep[name].data_ptr())
        for i in range(3):
            for f in _CORR_FIELDS:
                setattr(s.corrnet[i], f, keep[f"c{i}.{f}"].data_ptr())
        self.struct = s

    @property
    def ref(self):
        return C.byref(self.struct)

    def corrnet_sets(self, a: int, b: int, c: int):
        """A C array of three CorrNet weight sets (indices into corr_conv1) for imvs_corrnet."""
        arr = (_lib.CorrNetWeights * 3)()
        for j, i in enumerate((a, b, c)):
            for f in _CORR_FIELDS:
                setattr(arr[j], f, getattr(self.struct.corrnet[i], f))
        return arr
