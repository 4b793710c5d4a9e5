"""cvpytorch_b200 -- B200 (sm_100a) native implementation of CvPytorch's detector forward hot path.

Host code is Python + PyTorch tensors (plumbing only); all hot-path arithmetic runs in hand-written CUDA
behind the C ABI of libcvb200.so (include/cvb200.h).  There is no CPU fallback.
"""
__version__ = '0.1.0'
