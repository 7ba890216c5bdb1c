"""B200-native (sm_100a) implementation of the dense-prediction hot path of
HairongLuo/monocular-depth-estimation-cil: decoder/fusion/head layers, the training loss and the
evaluation reductions, behind the reference's own Python signatures.

The directory name is not a Python identifier; import it through the repo-root shim
(``import depth_b200``) or add this directory to ``sys.path`` to shadow the reference's
``util`` / ``network`` modules.
"""
from . import _lib                      # noqa: F401
from . import util                      # noqa: F401
from . import config                    # noqa: F401
from . import evaluation                # noqa: F401
from .graphs import GraphedTrainStep    # noqa: F401
from .util import (scale_invariant_loss, silog_loss, gradient_loss, edge_aware_loss, combined_loss,  # noqa: F401
                   absolute_relative_error, delta_thres, evaluation_metrics, evaluate_model_sums,
                   per_pixel_scale_invariant_loss, delta_counts)

__all__ = ["util", "scale_invariant_loss", "silog_loss", "gradient_loss", "edge_aware_loss", "combined_loss",
           "absolute_relative_error", "delta_thres", "evaluation_metrics", "evaluate_model_sums",
           "per_pixel_scale_invariant_loss", "delta_counts"]
