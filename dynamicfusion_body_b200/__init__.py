"""dynamicfusion_body_b200 -- B200 (sm_100a) hot paths of nintendops/DynamicFusion_Body behind its own Python surface.

    from dynamicfusion_body_b200 import Fusion, FusionDM, FusionDM_GPU

Importing the classes requires the in-tree CUDA library (python -m dynamicfusion_body_b200.build); there is no
CPU or eager fallback.
"""

__all__ = ["Fusion", "FusionDM", "FusionDM_GPU"]


def __getattr__(name):
    if name in __all__:
        from . import fusion
        return getattr(fusion, name)
    raise AttributeError(name)
