"""Phase clock stamps of CTA 0 of the tcgen05 convolution (ConvGRU, TF32 mode): staging / MMA / epilogue cycles.
ncu cannot replay that kernel on this pool's driver, hence the in-kernel stamps.  Needs a GPU."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, itermvs_b200
from itermvs_b200 import _lib
L = _lib.lib(); L.imvs_tc5_debug_clocks.argtypes = [C.c_int, C.c_void_p]; L.imvs_tc5_debug_clocks.restype = C.c_int
dev = torch.device("cuda:0")
z = np.load("tests/golden/dtu_weights.npz"); W = {k: torch.from_numpy(z[k]) for k in z.files}
m = itermvs_b200.Pipeline(4, test=True); m.load_state_dict(W); m = m.to(dev).eval()
upd = m.iter_mvs.update
h = torch.tanh(torch.randn(1, 32, 128, 160)).to(dev); x = (torch.randn(1, 11, 128, 160) * 0.5).to(dev)
_lib.set_conv_passes(1)
for it in range(3): upd.gru(h, x)
L.imvs_tc5_debug_clocks(1, None)
hn = h.permute(0, 2, 3, 1).contiguous(); x16 = torch.zeros(1, 128, 160, 16, device=dev); x16[..., :11] = x.permute(0, 2, 3, 1)
names = ["start", "tile cp.async issued", "cp.async landed", "rounded+synced", "weights landed", "MMAs issued", "MMAs done", "epilogue done", "end"]
for tag in ("warm",):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record(); upd.gru.forward_nhwc(hn.clone(), x16, wref=upd._packed(dev).ref); ev[1].record(); torch.cuda.synchronize()
    out = (C.c_longlong * 16)(); L.imvs_tc5_debug_clocks(1, out)
    t = list(out)[:9]
    print("gru (zr+q) event time us:", ev[0].elapsed_time(ev[1]) * 1e3, " | q-kernel CTA0 phases (cycles since start):")
    for n_, v in zip(names, t): print(f"   {n_:24s} {v - t[0]:8d}")
L.imvs_tc5_debug_clocks(0, None)
