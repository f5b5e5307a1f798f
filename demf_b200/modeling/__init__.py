"""DeMF model classes registered under the reference's names (demf/modeling/__init__.py):
importing this package populates the registries, as `import demf` does upstream."""
from . import coders, detectors, encoder, heads, layers  # noqa: F401
from .coders import DeMFClassAgnosticBBoxCoder
from .detectors import DeMFVoteNet
from .encoder import DeformableDetrEncoder
from .heads import DeMFVoteHead
from .layers import DeMFTransformerDecoderLayer, PositionEmbeddingLearned

__all__ = ['DeMFVoteNet', 'DeMFVoteHead', 'DeMFTransformerDecoderLayer',
           'PositionEmbeddingLearned', 'DeMFClassAgnosticBBoxCoder', 'DeformableDetrEncoder']
