class PerspectiveCameras:
    """Import-only placeholder (reference utils/train_utils.py:8 imports the name but the live path never builds one)."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("forge_b200.compat: PerspectiveCameras is not used by the forge_b200 backend")
