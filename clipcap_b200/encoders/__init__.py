from clipcap_b200.encoders.base import get_encoder, get_encoder_from_config, get_encoder_from_model  # noqa: F401
from clipcap_b200.encoders.config import EncoderConfig  # noqa: F401
from clipcap_b200.encoders.clip import CLIPModel, ViTImageTower  # noqa: F401
