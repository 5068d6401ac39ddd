"""freud_b200 -- B200-native SAE training and feature search behind ksadov/FREUD's own Python API.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all hot-path arithmetic runs in the
hand-written sm_100a kernels of libfreud_b200.so, reached through the C ABI of include/freud_b200.h.
"""
__version__ = "0.1.0"
