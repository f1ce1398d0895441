"""dlv3p-b200: the DeepLabV3+ encoder head of tf-keras-deeplabv3p-model-set as sm_100a CUDA kernels
behind a C ABI (include/dlv3p.h).  Import name: `dlv3p_b200` (see ../dlv3p_b200/__init__.py; the
directory name carries hyphens and cannot be imported directly)."""
from . import ffi, sharding, keras_bridge, metrics, xception  # noqa: F401
from .ffi import (Context, DeviceBuffer, PinnedBuffer, Dlv3pError, load_library, device_count,  # noqa: F401
                  device_info, f32_to_bf16_bits, bf16_bits_to_f32)
from .head import (ASPPBlock, ASPPLiteBlock, DecoderBlock, DeepLabHead, get_deeplabv3p_head,  # noqa: F401
                   BACKBONE_CHANNELS, MODEL_TYPES, atrous_rates)

from .xception import DeepLabV3PlusXception, get_deeplabv3p_xception  # noqa: F401

__version__ = '0.2.0'
