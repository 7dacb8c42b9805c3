from forge_b200.models.encoder import *  # noqa: F401,F403
from forge_b200.models.encoder import Encoder3D, get_resnet50  # noqa: F401
