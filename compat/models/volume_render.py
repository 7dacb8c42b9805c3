from forge_b200.models.volume_render import *  # noqa: F401,F403
from forge_b200.models.volume_render import VolRender  # noqa: F401
