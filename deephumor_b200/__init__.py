"""deephumor_b200: B200-native (sm_100a) batched caption generation with the DeepHumor model-class API.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all compute on the path is
hand-written CUDA behind the C ABI declared in ``include/deephumor_b200.h`` (``libdeephumor_sm100.so``).
There is no CPU fallback: constructing a model or calling a kernel without the library raises.
"""
__version__ = '0.1.0'
