#!/usr/bin/env python
"""One training step at the BENCHMARK configuration (640x512, 4 source views, D=32, 4 iterations, batch 1) compared with
the same step of the reference itself (build container only: needs /root/reference; ~1 minute of CPU).

The reference's Pipeline.train() forward + full_loss + backward runs on the CPU; itermvs_b200's training path runs with
its fused plane-sweep kernels (forward and backward, 148 persistent blocks) executed through tests/cusim.  Prints both
losses, the agreement of the arg-max bins of every prediction and the worst per-parameter gradient-norm deviation.
Result recorded in profiles/model_training_step_r01.txt.

    python tools/full_size_training_check.py
"""
import sys, os, time, ctypes as C, numpy as np, torch, warnings
import torch.nn.functional as F
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'cusim'))
import cusim_build, itermvs_b200
from itermvs_b200 import _lib, training
from itermvs_b200.synthetic import make_sample, plane_depth_map
torch.set_num_threads(8)
W,H,NS,IT=640,512,4,4
w={k: torch.from_numpy(v) for k,v in np.load(os.path.join(ROOT, 'tests', 'golden', 'dtu_weights.npz')).items()}
s=make_sample(W,H,n_src=NS,batch=1,seed=0,scene="plane")
d0=torch.from_numpy(plane_depth_map(W,H).astype(np.float32))[None,None]
gt={"level_0":d0,"level_2":F.interpolate(d0,scale_factor=.25,mode="nearest")}; mask={k: torch.ones_like(v) for k,v in gt.items()}
sys.path.insert(0,'/root/reference')
from models.net import Pipeline as RP, full_loss as rloss
ref=RP(iteration=IT,test=False); ref.load_state_dict(w,strict=True); ref.train()
t=time.time()
out=ref(s["imgs"],s["proj_matrices"],s["depth_min"],s["depth_max"])
l=rloss(out["depths"],out["depths_upsampled"],out["confidences"],gt,mask,s["depth_min"],s["depth_max"]); l.backward()
print("reference step", round(time.time()-t,1),"s loss",l.item()); sys.stdout.flush()
rg={k:(p.grad.double().norm().item() if p.grad is not None else 0.0) for k,p in ref.named_parameters()}
rarg=[p.argmax(1) for p in out["depths"]["probability"]]
lib=C.CDLL(cusim_build.build())
for name in ("imvs_compose_projections","imvs_warpcorr_init","imvs_warpcorr_iter","imvs_warpcorr_init_backward","imvs_warpcorr_iter_backward","imvs_last_error"):
    fn=getattr(lib,name); fn.restype,fn.argtypes=_lib._SIGNATURES[name]
training._L=lambda: lib; training._st=lambda: None; training._chk=lambda t,n: t.float().contiguous()
os.environ["CUSIM_SMS"]="148"
m=itermvs_b200.Pipeline(iteration=IT,test=False); m.load_state_dict(w,strict=True); m.train()
t=time.time()
o=training.pipeline_train_forward(m,s["imgs"],s["proj_matrices"],s["depth_min"],s["depth_max"])
l2=itermvs_b200.full_loss(o["depths"],o["depths_upsampled"],o["confidences"],gt,mask,s["depth_min"],s["depth_max"])
print("ours forward", round(time.time()-t,1),"s loss",l2.item()); sys.stdout.flush()
l2.backward()
print("ours step total", round(time.time()-t,1),"s")
agree=[float((a==p.argmax(1)).float().mean()) for a,p in zip(rarg,o["depths"]["probability"])]
print("argmax agreement",agree)
worst=0; tot=np.sqrt(sum(v*v for v in rg.values()))
for k,p in m.named_parameters():
    g=p.grad.double().norm().item() if p.grad is not None else 0.0
    worst=max(worst,abs(g-rg[k])/max(rg[k],1e-3*tot))
print("total grad norm ref",tot,"worst relative grad-norm deviation",worst)
