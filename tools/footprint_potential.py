#!/usr/bin/env python
"""How much register-level reuse the iteration plane sweep leaves on the table (round-2 design aid; geometry only, no GPU,
no kernel): for the benchmark scene (640x512, 4 source views, plane, hypotheses around the true depth) count the bytes of
bilinear taps the current thread mapping delivers to registers (every (pixel, hypothesis, view, tap) separately) and the
bytes of DISTINCT source texels per (pixel, view) and per (4-pixel item, view) -- the floors of a kernel that forms the
group dot products <ref(p), src(q)> once per (pixel, texel) and combines them with the bilinear weights.

    python tools/footprint_potential.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import itermvs_oracle as O
from itermvs_b200.synthetic import make_sample, plane_depth_map
W,H,S=640,512,4
s=make_sample(W,H,n_src=S,batch=1,seed=0,scene="plane")
dmin,dmax=s["depth_min"].float(),s["depth_max"].float()
inv_min=(1/dmin).view(1,1,1,1); inv_max=(1/dmax).view(1,1,1,1)
d2=F.interpolate(torch.from_numpy(plane_depth_map(W,H).astype(np.float32))[None,None],scale_factor=.25,mode="nearest")
for noise in (0.0, 0.02):
    nd=((1/d2-inv_max)/(inv_min-inv_max))
    nd=(nd+noise*torch.randn_like(nd)).clamp(0,1)
    h2,w2=H//4,W//4
    tot_now=tot_px=tot_item=0
    print("nd noise",noise)
    for l,C in ((1,16),(2,32),(3,48)):
        ds=O.iteration_depth_samples(nd,l,inv_min,inv_max)   # [1,R,h2,w2]
        R=ds.shape[1]
        pm=s["proj_matrices"][f"level_{l}"].float()
        hf,wf=(h2*2,w2*2) if l==1 else ((h2,w2) if l==2 else (h2//2,w2//2))
        now=px_u=item_u=0
        for v in range(S):
            proj=O.compose_projection(pm[:,v+1],pm[:,0])
            u,vv=O.warp_sample_positions(proj,ds,hf,wf)     # [1,R,P]
            x0=torch.floor(u).long().clamp(-1,wf-1)[0]; y0=torch.floor(vv).long().clamp(-1,hf-1)[0]   # [R,P]
            # clamp like the kernel: base in [0,W-2]
            xb=x0.clamp(0,wf-2); yb=y0.clamp(0,hf-2)
            taps=torch.stack([yb*wf+xb, yb*wf+xb+1, (yb+1)*wf+xb, (yb+1)*wf+xb+1],0)   # [4,R,P]
            P=taps.shape[-1]
            t=taps.reshape(4*R,P).T.contiguous()                # [P,4R]
            ts,_=t.sort(dim=1)
            uniq_px=(1+(ts[:,1:]!=ts[:,:-1]).sum(1)).sum().item()
            ti=t.reshape(h2,w2//4,4*4*R)                         # items: 4 consecutive px of a row
            tis,_=ti.sort(dim=2)
            uniq_item=(1+(tis[:,:,1:]!=tis[:,:,:-1]).sum(2)).sum().item()
            now+=4*R*P; px_u+=uniq_px; item_u+=uniq_item
        b=C*4
        print(f"  level {l}: taps now {now*b/1e6:.1f} MB, unique per (px,view) {px_u*b/1e6:.1f} MB ({px_u/now:.2f}), unique per (4px item,view) {item_u*b/1e6:.1f} MB ({item_u/now:.2f})")
        tot_now+=now*b; tot_px+=px_u*b; tot_item+=item_u*b
    print(f"  total: now {tot_now/1e6:.1f} MB -> per-px footprint {tot_px/1e6:.1f} MB ({tot_px/tot_now:.2f}) -> per-item footprint {tot_item/1e6:.1f} MB ({tot_item/tot_now:.2f})")
