"""The preprocess encode loop as a caller of the ViT engine: `EncoderMapper` and the `NumpyWriter`
on-disk format of the reference (clipcap/preprocess/mapper.py, writer.py)."""
from clipcap_b200.preprocess.mapper import EncoderMapper  # noqa: F401
from clipcap_b200.preprocess.writer import NumpyWriter, save_config  # noqa: F401
