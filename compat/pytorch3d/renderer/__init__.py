from forge_b200.geometry import look_at_view_transform  # noqa: F401
from .cameras import PerspectiveCameras  # noqa: F401
