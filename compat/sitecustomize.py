"""Aliases the reference still uses but current numpy / Pillow dropped (models/model_utils.py:45, dataset/kubric.py:426)."""
try:
    import numpy as _np
    if not hasattr(_np, "float"):
        _np.float = float
except Exception:
    pass
try:
    from PIL import Image as _Image
    if not hasattr(_Image, "ANTIALIAS"):
        _Image.ANTIALIAS = _Image.LANCZOS
except Exception:
    pass
