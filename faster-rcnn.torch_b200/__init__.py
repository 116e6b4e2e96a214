"""B200-native Faster R-CNN detection hot path behind the reference's Lua surface.

Host mirror (Python, because this image has no Lua toolchain) of the names the reference exposes on the path:
Rect, Localizer, Anchors, nms, Detector, extract_roi_pooling_input and the vgg_small / vgg_large model factories.
All arithmetic runs in libfrcnn_b200.so (hand-written CUDA for sm_100a) through the C ABI of include/frcnn_b200.h;
PyTorch is used for device memory only.  There is no CPU fallback."""
from ._lib import FrcnnError, declared_functions, ffi, lib  # noqa: F401
from .rect import Rect  # noqa: F401
from .geometry import Anchors, Localizer  # noqa: F401
from .nms import nms, nms_segmented, nms_segmented_dev  # noqa: F401
from .models import Model, create_model, duplo_cfg, find_target_size, imgnet_cfg, vgg_large, vgg_small  # noqa: F401
from .detector import Detector, DetectorPipeline, SpatialAdaptiveMaxPooling, extract_roi_pooling_input, roi_pooling_view  # noqa: F401
from .shard import reduce_timing, shard_frames, shard_segments  # noqa: F401
from .objective import allreduce_gradient, clean_anchors, create_objective, dp_allreduce, dp_init  # noqa: F401
from .optim import rmsprop, rmsprop_step, sync_running_stats  # noqa: F401
from .batch_iterator import FramePrefetcher  # noqa: F401
from . import t7  # noqa: F401,E402  (Torch7 .t7 snapshots: utilities.lua:113-134)
