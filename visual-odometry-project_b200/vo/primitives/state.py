"""Pipeline state: current / previous frame and pose (reference: src/vo/primitives/state.py)."""
import numpy as np

from vo.helpers import to_cartesian_coordinates, to_homogeneous_coordinates

__all__ = ["State"]


def _as_4x4(pose):
    if pose.shape == (3, 4):
        return np.concatenate((pose, np.array([[0, 0, 0, 1]])), axis=0)
    return pose


class State:
    def __init__(self, initial_frame, bearing_threshold: float = 0.0075) -> None:
        self.curr_pose = np.eye(4)
        self.curr_frame = initial_frame
        self.prev_pose = None
        self.prev_frame = None
        self._bearing_threshold = bearing_threshold

    def update_from_matches(self, matches) -> None:
        self.prev_frame, self.prev_pose = self.curr_frame, self.curr_pose
        self.curr_frame, self.curr_pose = matches.frame2, None

    def update_with_local_pose(self, pose: np.ndarray) -> None:
        """`pose` maps camera-1 coordinates into camera-2 coordinates (state.py:25-37)."""
        self.curr_pose = self.prev_pose @ np.linalg.inv(_as_4x4(pose))
        self.curr_frame.features.set_pose_for_new_tracks(self.curr_pose)

    def update_with_world_pose(self, pose: np.ndarray) -> None:
        """`pose` maps world coordinates into the camera (state.py:39-51)."""
        self.curr_pose = np.linalg.inv(_as_4x4(pose))
        self.curr_frame.features.set_pose_for_new_tracks(self.curr_pose)

    def update_with_local_landmarks(self, landmarks: np.ndarray, keypoints_mask: np.ndarray) -> None:
        """Landmarks given in the previous camera's frame (state.py:53-68)."""
        world = to_cartesian_coordinates(self.prev_pose @ to_homogeneous_coordinates(landmarks))
        self.update_with_world_landmarks(world, keypoints_mask)

    def update_with_world_landmarks(self, landmarks: np.ndarray, keypoints_mask: np.ndarray) -> None:
        feats = self.curr_frame.features
        assert np.sum(keypoints_mask) == len(landmarks), "Mismatch in length"
        assert np.all(feats.state[keypoints_mask] == 1), "Already triangulated point"
        feats.landmarks[keypoints_mask] = landmarks
        feats.state[keypoints_mask] = 2
        self._check_landmarks()
        assert not np.any(np.isnan(feats.landmarks[feats.state == 2])), "NaN in triangulated landmarks"

    def _check_landmarks(self) -> None:
        """Drop landmarks that lie behind the current or the previous camera (state.py:92-110)."""
        feats = self.curr_frame.features
        hom = to_homogeneous_coordinates(feats.landmarks)
        in_curr = to_cartesian_coordinates(np.linalg.inv(self.curr_pose) @ hom)
        in_prev = to_cartesian_coordinates(np.linalg.inv(self.prev_pose) @ hom)
        behind = (in_curr[:, 2].flatten() < 0) | (in_prev[:, 2].flatten() < 0)
        feats.landmarks[behind] = np.nan
        self.reset_outliers(behind)

    def get_frame(self):
        return self.curr_frame

    def get_pose(self) -> np.ndarray:
        return self.curr_pose

    def get_landmarks(self) -> np.ndarray:
        return self.curr_frame.features.landmarks

    def get_keypoints(self) -> np.ndarray:
        return self.curr_frame.features.keypoints

    def compute_candidates(self) -> None:
        """Matched-but-untriangulated tracks whose bearing angle is large enough (state.py:139-165)."""
        feats = self.curr_frame.features
        ends = feats.matched_candidate_inliers_keypoints
        starts = feats.matched_candidate_inliers_tracks
        pose_start = feats.matched_candidate_inliers_poses
        pose_end = np.stack([self.curr_pose] * pose_start.shape[0], axis=0)
        angles = self._calculate_bearing_angle(self.curr_frame.sensor, pose_start, pose_end, starts, ends)
        feats.candidate_mask[feats.matched_candidate_inliers] = angles >= self._bearing_threshold

    def reset_outliers(self, outliers: np.ndarray) -> None:
        """Outliers fall back to 'unmatched' and restart their track here (state.py:167-178)."""
        feats = self.curr_frame.features
        feats.state[outliers] = 0
        feats.tracks[outliers] = feats.keypoints[outliers]
        feats.poses[outliers] = self.curr_pose

    def _calculate_bearing_angle(self, camera, T1, T2, points1, points2) -> np.ndarray:
        """Angle between the viewing rays of a track's start and end, in the world frame (state.py:180-229)."""
        assert len(points1) == len(points2), "Points must have same length"
        assert points1.ndim == 3 and points2.ndim == 3, "Points must have three dimensions"
        assert not np.any(np.isnan(points1)), "Points1 contains invalid points"
        assert not np.any(np.isnan(points2)), "Points2 contains invalid points"
        assert not np.any(np.isnan(T1)), "Invalid start poses"
        assert not np.any(np.isnan(T2)), "Invalid end poses"
        ray1 = np.matmul(T1[:, :3, :3], camera.to_normalized_image_coordinates(points1)).reshape(-1, 3)
        ray2 = np.matmul(T2[:, :3, :3], camera.to_normalized_image_coordinates(points2)).reshape(-1, 3)
        cosang = np.sum(ray1 * ray2, axis=-1) / (np.linalg.norm(ray1, axis=-1) * np.linalg.norm(ray2, axis=-1))
        return np.arccos(cosang)
