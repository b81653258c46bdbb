"""Layer sub-modules (mirrors the reference's dgmr/layers/__init__.py:3-5 exports used on the hot path)."""
from .Attention import AttentionLayer
from .ConvGRU import ConvGRU, ConvGRUCell

__all__ = ["AttentionLayer", "ConvGRU", "ConvGRUCell"]
