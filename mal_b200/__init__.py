"""mal_b200: B200-native (sm_100a) implementation of the MAL photometric hot path.

The package mirrors the reference's Python call surface on top
of hand-written CUDA kernels reached through the C-ABI library `libmal_b200.so`
(declared in include/mal_b200.h).  There is no CPU fallback: importing the package
is cheap, but calling any operator without the CUDA library and a CUDA tensor raises.
"""
__version__ = "0.1.0"
