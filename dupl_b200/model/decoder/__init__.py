from .conv_head import LargeFOV  # noqa: F401
