"""Same import surface as ``deephumor.models`` (deephumor/models/__init__.py:1-25), minus the reference's dead
``TransformerEncoder``."""
from .encoders import ImageEncoder, ImageLabelEncoder, LabelEncoder
from .rnn_models import LSTMDecoder
from .transformers import SelfAttentionTransformerDecoder, TransformerDecoder
from .caption_models import (CaptioningLSTM, CaptioningLSTMWithLabels, CaptioningTransformer,
                             CaptioningTransformerBase)

__all__ = ['ImageEncoder', 'ImageLabelEncoder', 'LSTMDecoder', 'TransformerDecoder',
           'SelfAttentionTransformerDecoder', 'CaptioningLSTM', 'CaptioningLSTMWithLabels',
           'CaptioningTransformerBase', 'CaptioningTransformer']
