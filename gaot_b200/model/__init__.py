from .gaot import GAOT  # noqa: F401
from .layers.magno import MAGNOConfig, MAGNOEncoder, MAGNODecoder  # noqa: F401
from .layers.attn import AttentionConfig, TransformerConfig, Transformer  # noqa: F401
