"""models.model_single_pose_estimator for the reference's scripts (forge_b200 graph + reference PoseEstimator3D)."""
from forge_b200.models import model_single_pose_estimator as _m


class FORGE_poseEstimator3D(_m.FORGE_poseEstimator3D):
    def __init__(self, config):
        from models.pose_estimator_3d import PoseEstimator3D
        try:
            from utils.geo_utils import mat2quat
        except ImportError:
            mat2quat = None
        super().__init__(config, pose_estimator_3d_cls=PoseEstimator3D, mat2quat=mat2quat)
