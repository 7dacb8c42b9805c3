from forge_b200.models.rotate import *  # noqa: F401,F403
from forge_b200.models.rotate import Rotate_world  # noqa: F401
