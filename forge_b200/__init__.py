"""forge_b200 -- B200-native (sm_100a) backend for the FORGE render / rotate hot path.

    forge_b200.models     drop-in mirrors of the reference's models/*.py (VolRender, Rotate_world, Encoder3D, ConvGRU_3D, FORGE)
    forge_b200.ops        torch-facing wrappers of the C ABI (include/forge_b200.h, libforge_b200.so)
    forge_b200.pipeline   StreamedRenderer (host-to-host, copies overlapped), GraphedVolRender (CUDA-graph replay)
    forge_b200.refine     prepare_for_pose_refinement (test-time pose optimisation loop)

CUDA is mandatory: there is no CPU path, and a missing / stale shared library is rebuilt or raises.
"""
__version__ = "0.1.0"
