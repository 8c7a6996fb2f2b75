"""clipcap_b200 — B200-native drop-in for the ClipCap hot path (image -> CLIP ViT-L/14 -> mapper -> GPT-2 decode).

Same top-level surface as the reference package (clipcap/__init__.py:1-2): ``get_encoder``, ``get_encoder_from_model``,
``load``. Every forward goes through libclipcap_b200.so (hand-written sm_100a kernels behind a C ABI); there is no
PyTorch or CPU fallback path.
"""
from clipcap_b200.encoders import get_encoder, get_encoder_from_model  # noqa: F401
from clipcap_b200.model import load  # noqa: F401

__version__ = "0.1.0"
