"""yagi_b200 -- B200-native drop-in for yagi's polyphase filterbank channelizer path.

The product is libyagi_b200.so (hand-written CUDA for sm_100a behind the C ABI in
include/yagi_b200.h).  This package is the thin host-side mirror of the reference's object
protocol; it fails loudly if the library cannot be loaded -- there is no CPU fallback.
"""
from ._buffers import PinnedArray
from ._lib import launch_count
from .error import (ConfigError, InternalError, ModeError, NoConvergenceError, RangeError, ValueError_, YagiError)
from .filter import FirFilt, fir_design_kaiser
from .multichannel import ANALYZER, SYNTHESIZER, FirPfbCh, FirPfbCh2, FirPfbChType
from .gather import all_gather_frames, channel_major
from .sharding import TimeShard, firpfbch2_time_shards, stream_shards
from ._numa import bind_to_gpu_numa_node, gpu_numa_node

__all__ = [
    "ANALYZER", "SYNTHESIZER", "FirPfbChType", "FirPfbCh2", "FirPfbCh", "FirFilt", "fir_design_kaiser",
    "PinnedArray", "TimeShard", "firpfbch2_time_shards", "stream_shards", "all_gather_frames", "channel_major",
    "bind_to_gpu_numa_node", "gpu_numa_node", "launch_count",
    "YagiError", "InternalError", "ConfigError", "ValueError_", "RangeError", "ModeError", "NoConvergenceError",
]
