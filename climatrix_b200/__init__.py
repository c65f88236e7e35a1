"""climatrix_b200 -- B200-native IDW / ordinary-kriging reconstruction (one hot path of climatrix).

Hand-written sm_100a CUDA behind a C ABI (``include/climatrix_b200.h``), a Python
host layer that mirrors climatrix's reconstructor plug-in API.  No CPU fallback.
"""

__version__ = "0.1.0"
