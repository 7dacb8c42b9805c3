from forge_b200.models.fusion import *  # noqa: F401,F403
from forge_b200.models.fusion import ConvGRU_3D, ConvGRUCell_3D  # noqa: F401
