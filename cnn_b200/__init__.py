"""cnn_b200 -- B200-native (sm_100a) backend for the hermosayhl/CNN train-step hot path.

The product is the C-ABI library `libcnn_b200.so` (include/cnn_b200.h, cnn_b200/csrc) plus
the C++17 host mirror of the reference layer API (cnn_b200/host).  This Python package is
the thin ctypes face of that library used by tests/ and bench.py; torch is only plumbing
(device memory, streams, torch.distributed).
"""
