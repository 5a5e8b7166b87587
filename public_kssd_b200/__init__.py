"""kssd-b200: B200-native (sm_100a) sketch -> index -> compare hot path of Kssd.

The compute lives in libkssd_b200.so (CUDA, built in-tree from csrc/); `kssd` mirrors the reference's
stage functions on top of its C-ABI, `hostfmt` holds the on-disk formats, `synth` the deterministic
synthetic inputs.  There is no CPU implementation in this package.
"""
__all__ = ["capi", "kssd", "hostfmt", "synth"]
