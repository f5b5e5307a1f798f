"""Minimal stand-ins for the mmcv / mmdet / mmdet3d surface that haoy945/DeMF imports.

The reference is a plug-in of mmdetection3d (requirements.txt:2-4); none of those packages
exist here, and their CUDA extensions are exactly the hot path this repository re-implements.
This sub-package provides, under the upstream names and call signatures, only what the DeMF
VoteNet path touches: registries and the python-dict config loader, the point ops and
MSDeformAttn wrappers over libdemf_b200.so, and the nn.Modules that host them.
"""
