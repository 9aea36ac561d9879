from .blur import GaussianBlur, Glow  # noqa
from .color import FillColor, HSLShift  # noqa
from .protocol import BasicEffect, Effect  # noqa
from .style import DropShadow  # noqa
