"""movis_b200 -- B200-native (sm_100a) frame compositor behind the movis Python API.

Covers the per-frame render path of movis (``Composition.__call__``): scene graph, keyframe
engine, the 18 blend modes + matte modes of ``alpha_composite``, GaussianBlur / Glow / DropShadow /
FillColor / HSLShift / ChromaKey, draft preview levels and a device-resident render cache.  Pixel
work runs in hand-written CUDA kernels behind a C ABI (``include/movis_b200.h``); there is no CPU
fallback.  Usage mirrors the reference::

    import movis_b200 as mv
    scene = mv.layer.Composition(size=(3840, 2160), duration=5.0)
    scene.add_layer(mv.layer.Image(rgba_array), scale=1.2, rotation=10, blending_mode='overlay')
    frame = scene(0.5)                       # numpy (H, W, 4) uint8, like movis
    frames = scene.render_frames(times)      # CUDA tensor (N, H, W, 4), batched keyframes
"""
from . import effect  # noqa
from . import layer  # noqa
from .attribute import Attribute, AttributesMixin, transform_to_hashable  # noqa
from .enum import (AttributeType, BlendingMode, Direction, Easing, MatteMode,  # noqa
                   TextAlignment)
from .imgproc import alpha_composite  # noqa
from .layer.protocol import AUDIO_SAMPLING_RATE  # noqa
from .motion import Motion  # noqa
from .ops import crop  # noqa
from .transform import Transform  # noqa
from .util import to_rgb  # noqa

__version__ = "0.1.0"
