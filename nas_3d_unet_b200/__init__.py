"""nas_3d_unet_b200 - the supernet / searched-net compute path of woodywff/nas_3d_unet,
rebuilt for B200 (sm_100a): the reference's PyTorch module surface over hand-written CUDA
kernels reached through a C-ABI shared library (include/nas3d_b200.h)."""
__version__ = "0.1.0"
