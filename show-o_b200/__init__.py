"""show-o_b200: B200-native (sm_100a) engine for the Show-o hot path.

Importable as `showo_b200` (see ../showo_b200.py).  Mirrors the reference's `models/__init__.py` surface:
    from showo_b200 import Showo, MAGVITv2, get_mask_chedule, CLIPVisionTower
"""
from ._lib import ShowoError, load as load_library  # noqa: F401
from .schedules import cosine_schedule, get_mask_chedule, linear_schedule, step_schedule  # noqa: F401
from .showo_model import Showo  # noqa: F401
from .magvit_model import MAGVITv2  # noqa: F401
from . import masks  # noqa: F401
from . import editing  # noqa: F401  (inpainting / extrapolation flows of inference_t2i.py on top of t2i_generate)


def __getattr__(name):          # CLIPVisionTower pulls in `transformers`: import it only when asked for
    if name == "CLIPVisionTower":
        from .clip_tower import CLIPVisionTower
        return CLIPVisionTower
    raise AttributeError(name)
