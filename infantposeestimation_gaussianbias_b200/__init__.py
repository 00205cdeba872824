"""B200-native heatmap codec (encode / six-term fusion loss / decode) for
MarkJhonBao/InfantPoseEstimation_GaussianBias's hot path.

The arithmetic lives in libgbcodec.so (hand-written sm_100a CUDA behind the C
ABI of include/gbcodec.h); this package is the host-side mirror of the
reference's Python entry points.  Importing it never touches the oracle and
there is no CPU fallback: ops raise if the library or a CUDA device is missing.
"""
from . import _native
from ._native import GbcodecError

__all__ = ["_native", "GbcodecError", "load", "FusionPoseLoss", "generate_heatmaps", "HeatmapGenerator",
           "decode_outputs", "head_decode", "soft_argmax", "decode_heatmaps", "inference", "patch_reference"]


def load():
    """Load libgbcodec.so and register the torch ops; raises if the library is absent."""
    _native.lib()
    from . import ops  # noqa: F401
    return _native.lib()


def __getattr__(name):
    # torch-dependent parts are imported lazily so that `import package` stays cheap
    if name in ("FusionPoseLoss", "decode_outputs", "head_decode", "soft_argmax", "SKELETON", "LOSS_KEYS"):
        from . import fusion_head
        return getattr(fusion_head, name)
    if name in ("generate_heatmaps", "HeatmapGenerator", "generate_heatmaps_clipped", "GenerateTarget"):
        from . import generate_heatmap
        return getattr(generate_heatmap, name)
    if name in ("decode_heatmaps", "inference", "flip_permutation", "HeatmapHeadStep"):
        from . import pose_estimator
        return getattr(pose_estimator, name)
    if name == "patch_reference":
        from .patch import patch_reference
        return patch_reference
    if name in ("FusedPoseLoss", "MorphologyShapeLoss", "OffsetRegressionLoss", "JointsMSELoss", "KeypointMSELoss",
                "CombinedLoss", "build_loss"):
        from . import losses
        return getattr(losses, name)
    if name in ("ops", "postprocess", "sharded", "patch", "losses", "generate_heatmap", "fusion_head", "pose_estimator", "host_step"):
        import importlib
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
