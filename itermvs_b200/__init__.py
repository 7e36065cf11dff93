"""itermvs_b200 -- B200 (sm_100a) implementation of the IterMVS hot path behind the reference's
own Python call surface (models/module.py, models/itermvs.py, models/net.py).

    from itermvs_b200 import Pipeline
    model = Pipeline(iteration=4, test=True).cuda().eval()
    model.load_state_dict(torch.load(ckpt)["model"])          # reference checkpoints load as-is
    out = model(imgs, proj_matrices, depth_min, depth_max)     # {"depths_upsampled", "confidence_upsampled"}

Inference: all compute runs in hand-written CUDA kernels (itermvs_b200/csrc) bound through the C ABI declared
in include/itermvs_b200.h.  Training (model.train()): the plane sweep runs on the fused kernels in both directions,
the convolution stacks under torch autograd (itermvs_b200/training.py, itermvs_b200/ddp.py).  There is no CPU fallback.
"""
from .ops import (differentiable_warping, depth_normalization, depth_unnormalization, upsample,  # noqa: F401
                  compose_projections, nchw_to_nhwc, nhwc_to_nchw)
from .estimator import (ConvGRU, CorrNet, DepthInitialization, Evaluation, IterMVS, PixelViewWeight,  # noqa: F401
                        Update)
from .pipeline import FeatureNet, Pipeline, full_loss  # noqa: F401
from .fusion import check_geometric_consistency, filter_depth_view  # noqa: F401
from .ddp import FlatBucketDDP, train_step  # noqa: F401
from . import io  # noqa: F401  (read_pfm / save_pfm / read_cam_file / read_pair_file / load_views)
from . import training  # noqa: F401  (Pipeline.train() path: FusedCorrInit / FusedCorrIter with CUDA backward)

__all__ = ["Pipeline", "FeatureNet", "IterMVS", "Evaluation", "Update", "ConvGRU", "CorrNet", "PixelViewWeight",
           "DepthInitialization", "differentiable_warping", "depth_normalization", "depth_unnormalization", "upsample",
           "compose_projections", "nchw_to_nhwc", "nhwc_to_nchw", "full_loss", "check_geometric_consistency", "filter_depth_view", "FlatBucketDDP", "train_step", "training"]
