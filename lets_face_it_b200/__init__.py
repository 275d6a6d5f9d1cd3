"""lets_face_it_b200 — B200-native conditional-Glow hot path of jonepatr/lets_face_it.

`lets_face_it_b200.glow` mirrors the reference's `glow_pytorch.glow` module API; every forward goes
through hand-written sm_100a kernels in `_lib/liblfi_b200.so` (C ABI: include/lfi_b200.h)."""
__version__ = "0.1.0"
