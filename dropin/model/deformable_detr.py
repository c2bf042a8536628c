import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from egtr_b200.model.deformable_detr import *  # noqa: F401,F403,E402
from egtr_b200.model.deformable_detr import (  # noqa: F401,E402
    DeformableDetrConfig, DeformableDetrFeatureExtractor, DeformableDetrModel, MultiScaleDeformableAttentionFunction,
)
