"""Parameter containers for the encoders selectable by `siamese_network(backbone=...)`.
Reference: model/backbone/__init__.py, deit.py:97-109, vit.py:1093."""
from .vit import VisionTransformer, deit_base_patch16_224, vit_base_patch16_224  # noqa: F401
