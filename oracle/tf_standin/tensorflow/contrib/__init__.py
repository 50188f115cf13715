"""tensorflow.contrib of the stand-in (see tensorflow/__init__.py)."""
from . import layers  # noqa: F401
from . import cudnn_rnn  # noqa: F401
from . import rnn  # noqa: F401
