from .composition import Composition, LayerItem  # noqa
from .layer_ops import AlphaMatte, LuminanceMatte  # noqa
from .media import Image, ImageSequence  # noqa
from .mixin import TimelineMixin  # noqa
from .protocol import BasicLayer, Layer  # noqa
from .texture import Stripe  # noqa
