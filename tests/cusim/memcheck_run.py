"""Run the plane-sweep / warping kernels (forward and backward) of the emulated library on guard-page buffers
(tests/cusim/guarded.py), every buffer once flush against a trailing and once against a leading inaccessible page.
Executed in a subprocess by tests/test_cusim_kernels.py::test_kernels_stay_inside_their_buffers: an out-of-bounds
access ends this process with SIGSEGV."""
import ctypes as C
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cusim_build  # noqa: E402
from guarded import guarded  # noqa: E402
from itermvs_b200 import _lib  # noqa: E402
from itermvs_b200.synthetic import make_sample, random_feature_pyramids  # noqa: E402


def main():
    lib = C.CDLL(cusim_build.build())
    for name, (res, args) in _lib._SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    P = lambda t: None if t is None else t.data_ptr()

    def ok(rc):
        assert rc == 0, lib.imvs_last_error().decode()

    os.environ["CUSIM_SMS"] = "3"
    for side in ("end", "start"):
        G = lambda t: guarded(t, side)
        for batch, n_src, w, h, d in ((1, 2, 64, 64, 8), (2, 1, 96, 64, 5), (1, 9, 64, 32, 4)):
            ref, srcs = random_feature_pyramids(w, h, n_src, batch, 3)
            s = make_sample(w, h, n_src=n_src, batch=batch, seed=3, scene="noise")
            feas, rts = [], []
            for l in (1, 2, 3):
                k = f"level{l}"
                feas.append(G(torch.stack([ref[k]] + list(srcs[k]), dim=1).permute(0, 1, 3, 4, 2).contiguous()))
                proj = G(s["proj_matrices"][f"level_{l}"].float().contiguous())
                rt = G(torch.zeros(batch, n_src, 12))
                ok(lib.imvs_compose_projections(P(proj), batch, n_src + 1, P(rt), None, None))
                rts.append(rt)
            h2, w2, h3, w3 = h // 4, w // 4, h // 8, w // 8
            dmin, dmax = G(s["depth_min"].float().repeat(4)[:4 * batch]), G(s["depth_max"].float().repeat(4)[:4 * batch])   # 16-byte multiples
            if batch > 1:       # [b] indexing must see the real values first
                dmin[:batch] = s["depth_min"].float(); dmax[:batch] = s["depth_max"].float()
            g = torch.Generator().manual_seed(1)
            # init: forward, aggregation, backward
            corr = G(torch.zeros(batch, n_src, d, h3 * w3, 8))
            ok(lib.imvs_warpcorr_init(P(feas[2]), P(rts[2]), P(dmin), P(dmax), None, P(corr), batch, n_src + 1, h3, w3, d, None))
            vw3 = G(torch.rand(batch, n_src, h3 * w3, generator=g))
            agg0 = G(torch.zeros(batch, d, h3 * w3, 8))
            ok(lib.imvs_aggregate_init(P(corr), P(vw3), P(agg0), batch, n_src, d, h3 * w3, None))
            gfea3 = G(torch.zeros_like(feas[2]))
            ok(lib.imvs_warpcorr_init_backward(P(feas[2]), P(rts[2]), P(dmin), P(dmax), None, P(G(torch.randn(corr.shape, generator=g))),
                                               P(gfea3), batch, n_src + 1, h3, w3, d, None))
            # iteration: forward and backward, hypotheses generated in the kernels from nd
            nd = G(torch.rand(batch, h2 * w2, generator=g))
            vw = G(torch.rand(batch, n_src, h2 * w2, generator=g))
            agg = G(torch.zeros(batch, 10, h2 * w2, 8))
            ok(lib.imvs_warpcorr_iter(P(feas[0]), P(feas[1]), P(feas[2]), P(rts[0]), P(rts[1]), P(rts[2]), P(nd), h2 * w2, 1, P(vw),
                                      P(dmin), P(dmax), None, None, None, P(agg), batch, n_src + 1, h2, w2, None))
            gf = [G(torch.zeros_like(f)) for f in feas]
            ok(lib.imvs_warpcorr_iter_backward(P(feas[0]), P(feas[1]), P(feas[2]), P(rts[0]), P(rts[1]), P(rts[2]), P(nd), h2 * w2, 1,
                                               P(vw), P(dmin), P(dmax), None, None, None, P(G(torch.randn(agg.shape, generator=g))),
                                               P(gf[0]), P(gf[1]), P(gf[2]), batch, n_src + 1, h2, w2, None))
            # the stand-alone operator in the reference's layouts, forward and backward
            c, dd = 16, 4
            fea = G(torch.randn(batch, c, h2, w2, generator=g))
            sp, rp = G(s["proj_matrices"]["level_2"][:, 1].float().contiguous()), G(s["proj_matrices"]["level_2"][:, 0].float().contiguous())
            dep = G(400.0 + 500.0 * torch.rand(batch, dd, h2, w2, generator=g))
            out = G(torch.zeros(batch, c, dd, h2, w2))
            rt = G(torch.zeros(4 * batch, 12)[:batch * 4])
            ok(lib.imvs_differentiable_warping(P(fea), P(sp), P(rp), P(dep), P(out), batch, c, h2, w2, dd, h2, w2, P(rt), None, None))
            gfea = G(torch.zeros_like(fea))
            ok(lib.imvs_differentiable_warping_backward(P(out), P(sp), P(rp), P(dep), P(gfea), batch, c, h2, w2, dd, h2, w2, P(rt), None, None))
            assert all(torch.isfinite(t).all() for t in (corr, agg0, gfea3, agg, gf[0], gf[1], gf[2], out, gfea))
    print("MEMCHECK-OK")




def whole_forward():
    """imvs_featurenet_forward + imvs_itermvs_forward (every inference kernel incl. the tensor-core engine) with all
    inputs, outputs, weights-independent scratch and workspaces on guard-page buffers (trailing guard)."""
    import numpy as np
    import itermvs_b200
    from itermvs_b200 import ops
    lib = C.CDLL(cusim_build.build())
    for name, (res, args) in _lib._SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib._lib = lib
    ops._chk = lambda t, name: t.contiguous()
    ops._stream = lambda: None
    golden = os.path.join(os.path.dirname(HERE), "golden", "dtu_weights.npz")
    with np.load(golden) as z:
        weights = {k: torch.from_numpy(z[k]) for k in z.files}
    w = h = 32
    n_src, iters = 2, 1
    m = itermvs_b200.Pipeline(iteration=iters, test=True)
    m.load_state_dict(weights, strict=True)
    m.eval()
    s = make_sample(w, h, n_src=n_src, batch=1, seed=4, scene="plane")
    x = guarded(s["imgs"]["level_0"].float().contiguous())
    n = n_src + 1

    def gbytes(nbytes):
        return guarded(torch.zeros((nbytes + 255) // 256 * 256, dtype=torch.uint8))[:nbytes]
    m.feature_net._ws[0] = ((n, h, w, "cpu"), gbytes(lib.imvs_featurenet_workspace_bytes(n, h, w)))
    pb = _lib.Problem(1, n, h, w, 32, iters)
    m.iter_mvs._workspaces[0] = ((1, n, h, w, 32, iters, "cpu"), gbytes(lib.imvs_forward_workspace_bytes(C.byref(pb))))
    with torch.no_grad():
        f1, f2, f3 = m.feature_net.forward_nhwc(x)
        f1, f2, f3 = guarded(f1), guarded(f2), guarded(f3)
        projs = [guarded(s["proj_matrices"][f"level_{l}"].float().contiguous()) for l in (1, 2, 3)]
        dmin, dmax = guarded(s["depth_min"].float().repeat(4)), guarded(s["depth_max"].float().repeat(4))
        out = tuple(guarded(torch.zeros(1, 1, hh, ww)) for hh, ww in ((h // 4, w // 4), (h, w), (h // 4, w // 4), (h, w)))
        m.iter_mvs.forward_packed(f1, f2, f3, projs[0], projs[1], projs[2], dmin, dmax, out=out)
    assert all(torch.isfinite(t).all() for t in out)
    print("MEMCHECK-FORWARD-OK")


if __name__ == "__main__":
    whole_forward() if "--forward" in sys.argv else main()
