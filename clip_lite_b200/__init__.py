"""B200-native JSD contrastive-loss hot path of CLIP-Lite (loss.py), behind the
reference's own JSDInfoMaxLoss interface.  CUDA kernels live in csrc/ and are
reached through the C ABI of include/jsd_b200.h."""
__version__ = "0.1.0"
