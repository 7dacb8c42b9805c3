from .camera_conversions import cameras_from_opencv_projection  # noqa: F401
