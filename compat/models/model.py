"""models.model for the reference's scripts: forge_b200's FORGE wired to the reference's own pose networks."""
from forge_b200.models import model as _m
from forge_b200.models.model import sequence_from_distance, chose_selected  # noqa: F401


class FORGE(_m.FORGE):
    def __init__(self, config):
        from models.pose_estimator_3d import PoseEstimator3D      # reference checkout
        from models.pose_estimator_2d import PoseEstimator2D
        try:
            from utils.geo_utils import mat2quat
        except ImportError:
            mat2quat = None
        super().__init__(config, pose_estimator_3d_cls=PoseEstimator3D, pose_estimator_2d_cls=PoseEstimator2D,
                         mat2quat=mat2quat)
