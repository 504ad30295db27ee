"""agenda_b200 — B200-native (sm_100a) implementation of AGenDA's heat-map data-generation hot path.

Only what the path needs: the CUDA kernels + C ABI (csrc/, include/agenda_b200.h), the ctypes binding (_lib),
tensor-level operators (ops), the drop-in AttnProcessor (processor), the daam-style trace shim (trace), the
post-processing / CLI mirrors (postprocess, data_generation, postprocess_heatmap) and seed sharding (sharding).
"""
from .processor import B200CrossAttnProcessor, UNetCrossAttentionHooker  # noqa: F401

__all__ = ["UNetCrossAttentionHooker", "B200CrossAttnProcessor"]
__version__ = "0.1.0"
