"""metro-b200: B200-native drop-in for the MeTRo (isarandi/metro-pose3d) inference hot path.

Everything that computes lives in ``libmetro.so`` (hand-written CUDA for sm_100a, C-ABI in
``include/metro.h``); this package is the host-side mirror of the reference's ``inference.py``
contract plus the pure-Python layer plan, joint tables and synthetic-weight generator.
"""
from .spec import NetSpec, CONFIGS
from .joints import JointInfo, model_joint_info, exported_joint_info, export_permutation

__all__ = ['NetSpec', 'CONFIGS', 'JointInfo', 'model_joint_info', 'exported_joint_info',
           'export_permutation', 'MetroModel', 'estimate_pose', 'SoftArgmax']


def __getattr__(name):
    if name in ('MetroModel', 'estimate_pose', 'SoftArgmax', 'conv2d'):
        from . import inference
        return getattr(inference, name)
    raise AttributeError(name)
