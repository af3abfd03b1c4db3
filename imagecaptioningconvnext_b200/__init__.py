"""B200-native (sm_100a) captioning hot path: drop-in modules for sa06840/ImageCaptioningConvNeXt.

``Encoder``, ``DecoderWithAttention`` and ``TransformerDecoder`` keep the reference's nn.Module API
(models/encoder.py, models/decoder.py, models/transformerDecoder.py) and run on hand-written CUDA kernels in
``libccx.so`` (C ABI: include/ccx.h).  No eager / CPU fallback exists.
"""
from .decoder import Attention, DecoderWithAttention  # noqa: F401
from .encoder import Encoder  # noqa: F401
from .transformerDecoder import (PositionalEncoding, TransformerDecoder,  # noqa: F401
                                 TransformerDecoderForAttentionViz)

__all__ = ["Encoder", "Attention", "DecoderWithAttention", "PositionalEncoding", "TransformerDecoder",
           "TransformerDecoderForAttentionViz"]
