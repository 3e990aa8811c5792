"""B200-native (sm_100a) implementation of the mSWE-GNN message-passing hot path.

Drop-in for the reference's ``models/models.py`` / ``models/gnn.py`` classes and the rollout /
training-step loops of ``training/train.py``; every numeric operation of the path runs in
hand-written CUDA kernels reached through the C ABI declared in ``include/swe_gnn_b200.h``.
There is no CPU fallback: constructing the native library wrapper without the built ``.so`` or
calling a model on a non-CUDA tensor raises.
"""
__version__ = "0.1.0"
