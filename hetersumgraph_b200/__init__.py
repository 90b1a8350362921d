"""hetersumgraph_b200 - B200-native WSWGAT message-passing path of HeterSumGraph (HSG / HDSG).

Drop-in modules (same constructors / forward / state_dict as the reference's module/GAT.py,
module/GATStackLayer.py, module/GATLayer.py) whose forward and backward run in hand-written
sm_100a CUDA kernels behind the C ABI of include/hsg_b200.h.  No DGL, no Triton, no CPU fallback.
"""
from . import _lib, synthetic  # noqa: F401
from ._lib import get_gemm_mode, set_gemm_mode  # noqa: F401
from .graph import BuildPipeline, DeviceTokenBatch, HeteroBatch, csc_pair_from_edges  # noqa: F401
from .modules import (MultiHeadLayer, MultiHeadSGATLayer, PositionwiseFeedForward, SWGATLayer, WSGATLayer, WSWGAT,  # noqa: F401
                      WSWGATUpdateLoop)

from .encoder import EncoderPlan, SentenceEncoder  # noqa: F401,E402
from .model import HSumDocGraph, HSumGraph  # noqa: F401,E402

__all__ = ["EncoderPlan", "SentenceEncoder", "HSumGraph", "HSumDocGraph", "BuildPipeline", "DeviceTokenBatch", "HeteroBatch", "csc_pair_from_edges", "MultiHeadLayer", "MultiHeadSGATLayer", "PositionwiseFeedForward", "SWGATLayer",
           "WSGATLayer", "WSWGAT", "WSWGATUpdateLoop", "synthetic"]
