"""change3d_b200 — B200-native (sm_100a) execution engine for the Change3D X3D-L hot path.

Public surface mirrors the reference's modules:
    change3d_b200.model.x3d.create_x3d
    change3d_b200.model.change_decoder.ChangeDecoder
    change3d_b200.model.trainer.Trainer / Encoder
All compute runs in libchange3d_b200.so (C ABI: include/change3d_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
