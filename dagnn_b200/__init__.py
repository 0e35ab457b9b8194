"""dagnn_b200 — B200-native (sm_100a) implementation of the DAGNN layer-wise forward hot path.

Public surface mirrors the reference's module API for this path:
  dagnn_b200.ogb.DAGNN / ASTNodeEncoder      <- ogbg-code/model/dagnn.py:16, ogbg-code/utils.py:7
  dagnn_b200.dvae.DAGNN / DAGNN_BN           <- dvae/dagnn.py:18, dvae/dagnn_bn.py:19
All arithmetic of the path runs in hand-written CUDA reached through the C-ABI library
`libdagnn_sm100.so` (include/dagnn_b200.h); there is no CPU or eager-PyTorch fallback.
"""
__version__ = "0.1.0"
