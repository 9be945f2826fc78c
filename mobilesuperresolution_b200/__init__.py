"""mobilesuperresolution_b200 -- B200-native (sm_100a) forward path of zhuzhui-2000/mobilesuperresolution.

Drop-in mirrors of the reference's nn.Modules for the super-resolution forward hot path; the arithmetic runs in
hand-written CUDA reached through the C ABI in ``include/b200sr.h``.  No Triton, no cuDNN dispatch, no CPU fallback.
"""
from .masks import BinaryConv2d, rounding  # noqa: F401
from .wdsr import BASIC_MODEL, NAS_MODEL_classic, AggregationLayer, Block, Model, WdsrPlan  # noqa: F401
from .nas import NAS_MODEL, BlockBSpeedEstimator  # noqa: F401
from .split import Conv_sep, MyAggregationLayer, Split_Block  # noqa: F401
from .graph import Graphed  # noqa: F401
from .frames import forward_u8_frames, psnr_u8, ssd_u8, u8_to_unit  # noqa: F401
from .video import (BasicVSR, BasicVSR_origin, MotionVectorVSR, SpyNet, flow_warp)  # noqa: F401
from .naive import Naive_model, NaiveBlock  # noqa: F401

__version__ = "0.1.0"


def get_model(params):
    """models/__init__.py:31-32 (``eval(params.model_type)(params)``) without the eval."""
    table = {"BASIC_MODEL": BASIC_MODEL, "NAS_MODEL": NAS_MODEL, "NAS_MODEL_classic": NAS_MODEL_classic}
    return table[params.model_type](params)
