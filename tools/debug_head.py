import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import itermvs_b200
from oracle import itermvs_oracle as O
dev = torch.device("cuda:0")
z = np.load("tests/golden/dtu_weights.npz"); W = {k: torch.from_numpy(z[k]) for k in z.files}
m = itermvs_b200.Pipeline(4, test=True); m.load_state_dict(W); m = m.to(dev).eval()
k = np.load("tests/golden/stage_kats.npz")
upd = m.iter_mvs.update; upd.return_probability = True
h = torch.from_numpy(k["gru_h"]).to(dev)
nd, prob = upd.depth_init(h)
pr = torch.softmax(torch.from_numpy(k["head_logits"]), 1)
e = (prob.cpu() - pr).abs(); print("random-h prob maxerr", float(e.max()), "at", np.unravel_index(int(e.argmax()), e.shape), "pmax", float(pr.max()))
upd = copy.deepcopy(m.iter_mvs.update)
with torch.no_grad():
    for p in upd.parameters(): p.zero_()
    targets = [0, 1, 2, 3, 4, 5, 100, 250, 251, 252, 253, 254, 255, 17, 64, 200]
    for j, tj in enumerate(targets):
        upd.depth_head[0].weight[j, j, 1, 1] = 1.0
        upd.depth_head[2].weight[j, j, 0, 0] = 1.0
        upd.depth_head[4].weight[tj, j, 0, 0] = 12.0
        upd.depth_head[4].weight[min(tj + 1, 255), j, 0, 0] += 9.0
h = torch.zeros(1, 32, 2, 8)
for j in range(16): h[0, j, j // 8, j % 8] = 1.0
upd.return_probability = True
nd, prob = upd.depth_init(h.to(dev))
w = {"iter_mvs.update." + kk: v.detach().cpu() for kk, v in upd.state_dict().items()}
nd_ref, prob_ref = O.depth_init(w, h)
e = (prob.cpu() - prob_ref).abs()
print("edge prob maxerr", float(e.max()), "at", np.unravel_index(int(e.argmax()), e.shape))
for j in range(16):
    py, px = j // 8, j % 8
    g, r = prob[0, :, py, px].cpu(), prob_ref[0, :, py, px]
    print(j, targets[j], "gpu top", int(g.argmax()), float(g.max()), "ref top", int(r.argmax()), float(r.max()), "nd", float(nd[0,0,py,px]), float(nd_ref[0,0,py,px]))
