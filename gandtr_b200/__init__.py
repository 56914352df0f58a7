"""gandtr_b200 -- B200-native (sm_100a) implementation of gandtr's data-parallel retrieval hot path:
CLAHE preprocessing, GeM + L2N + multi-scale + learned whitening, query x database scoring with top-k and mAP.

The compute lives in libgandtr_b200.so (hand-written CUDA behind the C ABI of include/gandtr_b200.h);
this package is the host-side mirror of the reference's plugin interface for that path.
"""
__version__ = "0.1.0"
