"""b200slam — host side of the B200-native ORB front-end hot path.

Python mirrors the reference's matcher / pose-estimator interfaces
(integration.feature_pipeline_bridge, integration.pose_bridge); all arithmetic of the
hot path runs in libb2s.so (hand-written sm_100a CUDA, C ABI in include/b2s.h).
Importing this package does not touch CUDA (fork/thread safety).
"""
from ._capi import B2SError, IDX_BITS, IDX_MASK, NONE_KEY, lib_path, load_library  # noqa: F401

__all__ = ["B2SError", "IDX_BITS", "IDX_MASK", "NONE_KEY", "lib_path", "load_library"]
