from clipcap_b200.model.model import ClipCapModel, ClipCapModelPrefixOnly, get_tokenizer  # noqa: F401
from clipcap_b200.model.config import Config, TrainingConfig  # noqa: F401
from clipcap_b200.model.mapper import TransformerMapper, TransformerMapperWindowed, MLPMapper  # noqa: F401
from clipcap_b200.model.lm import GPT2LM  # noqa: F401
from clipcap_b200.model.load import load  # noqa: F401
