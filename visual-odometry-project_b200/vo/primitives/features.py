"""Per-frame feature table (reference: src/vo/primitives/features.py).

Rows are keypoints; columns are the pixel location, optional descriptor, the landmark (NaN while
unknown), the life-cycle state (0 unmatched, 1 matched, 2 triangulated), a uid, and -- for matched
but not yet triangulated points -- where and at which camera pose the track started."""
import numpy as np

__all__ = ["Features"]


class Features:
    def __init__(self, keypoints: np.ndarray, landmarks: np.ndarray = None, uids: np.ndarray = None) -> None:
        assert keypoints.ndim == 3 and keypoints.shape[1:] == (2, 1), "Invalid shape for keypoints"
        n = keypoints.shape[0]
        self._keypoints = keypoints
        self.descriptors = None
        if landmarks is not None:
            assert landmarks.ndim == 3 and landmarks.shape[1:] == (3, 1), "Invalid shape for landmarks"
        self.landmarks = landmarks if landmarks is not None else np.full((n, 3, 1), np.nan)
        self.state = np.zeros(n)
        self._uids = uids
        self._tracks = keypoints.copy()                                   # a new track starts where it is seen
        self._poses = np.stack([np.eye(4)] * n) if n > 0 else np.empty((0, 4, 4))
        self._candidate_mask = np.zeros(n, dtype=bool)

    # ---- state masks (features.py:87-101) ------------------------------------------------
    @property
    def matched_candidate_inliers(self) -> np.ndarray:
        return self.state == 1

    @property
    def match_inliers(self) -> np.ndarray:
        return self.state >= 1

    @property
    def triangulate_inliers(self) -> np.ndarray:
        return self.state >= 2

    @property
    def p3p_inliers(self) -> np.ndarray:
        return self.state >= 2

    # ---- masked views (features.py:54-85) -------------------------------------------------
    @property
    def matched_candidate_inliers_tracks(self):
        return self._tracks[self.matched_candidate_inliers]

    @property
    def matched_candidate_inliers_poses(self):
        return self._poses[self.matched_candidate_inliers]

    @property
    def matched_candidate_inliers_keypoints(self):
        return self._keypoints[self.matched_candidate_inliers]

    @property
    def candidate_inliers_keypoints(self):
        return self._keypoints[self.candidate_mask]

    @property
    def matched_inliers_keypoints(self):
        return self._keypoints[self.match_inliers]

    @property
    def triangulated_inliers_keypoints(self):
        return self._keypoints[self.triangulate_inliers]

    @property
    def triangulated_inliers_landmarks(self):
        return self._landmarks[self.triangulate_inliers]

    @property
    def p3p_inliers_keypoints(self):
        return self._keypoints[self.p3p_inliers]

    # ---- columns with length checks (features.py:103-207) --------------------------------
    def _same_len(self, value, what, allow_none=True):
        if value is None and allow_none:
            return
        assert value.shape[0] == self._keypoints.shape[0], f"Unequal number of {what} and keypoints."

    @property
    def keypoints(self):
        return self._keypoints

    @keypoints.setter
    def keypoints(self, value):
        assert value.shape[0] == self._keypoints.shape[0], "Unequal number of keypoints."
        self._keypoints = value

    @property
    def state(self):
        return self._state

    @state.setter
    def state(self, value):
        self._same_len(value, "state", allow_none=False)
        self._state = value

    @property
    def descriptors(self):
        return self._descriptors

    @descriptors.setter
    def descriptors(self, value):
        self._same_len(value, "descriptors")
        self._descriptors = value

    @property
    def landmarks(self):
        return self._landmarks

    @landmarks.setter
    def landmarks(self, value):
        self._same_len(value, "landmarks", allow_none=False)
        self._landmarks = value

    @property
    def uids(self):
        return self._uids

    @uids.setter
    def uids(self, value):
        self._same_len(value, "uids")
        self._uids = value

    @property
    def tracks(self):
        return self._tracks

    @tracks.setter
    def tracks(self, value):
        self._same_len(value, "tracks")
        self._tracks = value

    @property
    def poses(self):
        return self._poses

    @poses.setter
    def poses(self, value):
        self._same_len(value, "poses")
        self._poses = value

    @property
    def candidate_mask(self):
        return self._candidate_mask

    @candidate_mask.setter
    def candidate_mask(self, value):
        self._same_len(value, "candidate_mask")
        self._candidate_mask = value

    @property
    def length(self) -> int:
        n = self._keypoints.shape[0]
        assert self._descriptors is None or self._descriptors.shape[0] == n, "Unequal number of descriptors and keypoints."
        assert self._landmarks is None or self._landmarks.shape[0] == n, "Unequal number of landmarks and keypoints."
        assert self._state.shape[0] == n, "Unequal number of state and keypoints."
        return n

    def set_pose_for_new_tracks(self, pose: np.ndarray) -> None:
        """Stamp the start pose of every track that begins in this frame (features.py:227-241)."""
        fresh = self.state == 0
        assert np.all(np.isnan(self.poses[fresh]))
        assert pose.shape == (4, 4) or (pose.ndim == 3 and pose.shape[1:] == (4, 4) and pose.shape[0] == np.sum(fresh)), \
            "Invlaid shape for pose"
        self._poses[fresh] = pose

    def mask(self, mask: np.ndarray) -> None:
        """Keep only the rows selected by a boolean mask (features.py:243-270)."""
        assert mask.shape == (self.length,), "Invalid mask shape"
        for name in ("_keypoints", "_state", "_landmarks", "_uids", "_tracks", "_poses", "_candidate_mask"):
            col = getattr(self, name)
            if col is not None:
                setattr(self, name, col[mask])
