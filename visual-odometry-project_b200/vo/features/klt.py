"""KLT tracker (reference: src/vo/features/klt.py).

track_features runs the pyramidal Lucas-Kanade tracker on the GPU (vo_klt_track_host: one warp per
keypoint) in place of cv2.calcOpticalFlowPyrLK, with the reference's parameters (klt.py:29-33) and
the same status / error filtering.  Corner seeding (cv2.goodFeaturesToTrack) and the feature-table
bookkeeping are host code, as in the reference."""
import sys

import numpy as np

from vo import _ops
from vo.primitives import Features, Frame, Matches

__all__ = ["KLTTracker"]


class KLTTracker:
    _feature_params = dict(maxCorners=500, qualityLevel=0.01, minDistance=8, blockSize=7)   # klt.py:24-26
    _lk_params = dict(winSize=(17, 17), maxLevel=2, criteria=(3, 10, 0.03))                  # klt.py:29-33 (EPS|COUNT = 3)
    _colors = np.random.randint(0, 255, (_feature_params["maxCorners"], 3))
    _error_threshold = 100

    def __init__(self, frame):
        self._min_inliers = 90
        self._num_features = None
        self._old_frame = frame
        self._frame = frame
        self._frame.features = Features(keypoints=self.find_corners(frame=self._frame))
        self._frame.features.uids = self._get_udis(self._frame.features.length)
        self._last_masks = []

    @property
    def frame(self) -> Frame:
        return self._frame

    @frame.setter
    def frame(self, frame: Frame) -> None:
        self._frame = frame
        self._frame.features = Features(self.find_corners(self._frame))

    @property
    def old_img_gray(self) -> np.ndarray:
        return self.to_gray(self._old_frame.image)

    @property
    def img_gray(self) -> np.ndarray:
        return self.to_gray(self._frame.image)

    def to_gray(self, img) -> np.ndarray:
        """klt.py:84-85 (cv2.cvtColor BGR2GRAY) on the device, bit-exact."""
        return _ops.bgr2gray(img)

    def _get_udis(self, length: int) -> np.ndarray:
        return np.random.randint(0, np.iinfo(np.int32).max, size=length, dtype=np.int32)

    def _fill_udis(self, array, target_length: int) -> np.ndarray:
        if array is None:
            return self._get_udis(target_length)
        return np.concatenate((array, self._get_udis(target_length - array.shape[0])))

    def find_corners(self, frame: Frame, mask=None, use_goodFeaturesToTrack=True) -> np.ndarray:
        """Shi-Tomasi corners (klt.py:87-115): vo_gftt_host on the GPU (a mask with holes, which the reference never
        builds, falls back to the host call)."""
        import cv2
        img = frame.image
        if img.ndim == 3:
            img = self.to_gray(img)
        if use_goodFeaturesToTrack and (mask is None or bool(np.all(mask))):
            # the reference's own call (mask all 255, klt.py:214-226) on the GPU: cv2's arithmetic, cv2's corner list
            fp = self._feature_params
            points = _ops.good_features_to_track(img, fp["maxCorners"], fp["qualityLevel"], fp["minDistance"], fp["blockSize"])
        elif use_goodFeaturesToTrack:
            points = cv2.goodFeaturesToTrack(img, mask=mask, **self._feature_params)
        else:
            dst = cv2.dilate(cv2.cornerHarris(img, 2, 3, 0.04), None)
            _, dst = cv2.threshold(dst, 0.01 * dst.max(), 255, 0)
            _, _, _, centroids = cv2.connectedComponentsWithStats(np.uint8(dst))
            crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 100, 0.001)
            points = cv2.cornerSubPix(img, np.float32(centroids), (5, 5), (-1, -1), crit)
        points = points.reshape((-1, 2, 1))
        self._num_features = points.shape[0]
        return points

    def update_features(self, new_keypoints: np.ndarray) -> Features:
        """Append freshly detected keypoints to the old frame's table (klt.py:117-189)."""
        old = self._old_frame.features
        m = new_keypoints.shape[0]
        keypoints = np.concatenate((old.keypoints, new_keypoints))
        landmarks = np.concatenate((old.landmarks if old.landmarks is not None else np.array([]), np.full((m, 3, 1), np.nan)))
        state = np.concatenate((old.state if old.state is not None else np.array([]), np.zeros(m)))
        uids = self._fill_udis(old.uids, m + old.length)
        tracks = np.concatenate((old.tracks, new_keypoints))
        poses = np.concatenate((old.poses if old.poses is not None else np.array([]), np.stack([np.eye(4)] * m)))
        cand = np.concatenate((old.candidate_mask if old.candidate_mask is not None else np.array([]),
                               np.zeros(keypoints.shape[0] - old.candidate_mask.shape[0]).astype(bool)))
        assert keypoints.shape[0] == landmarks.shape[0] == state.shape[0] == uids.shape[0] == tracks.shape[0] \
            == poses.shape[0] == cand.shape[0], \
            "The number of keypoints, landmarks, state, uids, tracks, poses and candidate_mask must be the same."
        feats = Features(keypoints=keypoints, landmarks=landmarks)
        feats.state, feats.uids, feats.tracks, feats.poses, feats.candidate_mask = state, uids, tracks, poses, cand
        return feats

    def track_features(self, curr_frame: Frame, new_frame: Frame) -> Matches:
        """Track the current frame's keypoints into the new frame (klt.py:191-280)."""
        import cv2
        self._old_frame, self._frame = curr_frame, new_frame
        if self._old_frame.features is None or self._old_frame.features.length < self._num_features * 0.8:
            if sys.gettrace() is not None:
                print("Adding new features")
            mask = np.ones_like(self.img_gray) * 255
            if self._old_frame.features is None:           # klt.py:217-222 (kept as written there)
                for x, y in self._old_frame.features.keypoints.reshape(-1, 2).astype(int):
                    cv2.circle(mask, (x, y), 5, 0, -1)
            fresh = self.find_corners(frame=self._old_frame, mask=mask)
            self._old_frame.features = self.update_features(new_keypoints=fresh)
            self._old_frame.features.uids = self._fill_udis(self._old_frame.features.uids, self._old_frame.features.length)

        win = self._lk_params["winSize"][0]
        _, max_count, eps = self._lk_params["criteria"]
        prev_pts = np.asarray(self._old_frame.features.keypoints, dtype=np.float32).reshape(-1, 2)
        # BGR frames go to the device as they are (the grayscale conversion of klt.py:57-62 is fused into the upload),
        # and the pyramid of a frame is built once: the previous call's `next` is this call's `prev`
        old_img, new_img = self._old_frame.image, self._frame.image
        next_pts, status, error = _ops.klt_track(old_img, new_img, prev_pts, win=win,
                                                 max_level=self._lk_params["maxLevel"], max_iters=max_count, epsilon=eps)
        next_pts = next_pts.reshape((-1, 2, 1))
        keep = np.logical_and(status.astype(bool), error < self._error_threshold)       # klt.py:244-249
        if sys.gettrace() is not None:
            print(f"{np.sum(keep)/keep.shape[0]*100:.2f}% inliers")
        self.frame.features = Features(keypoints=next_pts)
        self.frame.features.uids = self._fill_udis(self._old_frame.features.uids, next_pts.shape[0])
        self.frame.features.mask(keep)
        self._old_frame.features.mask(keep)
        idx = np.arange(0, self.frame.features.keypoints.shape[0]).reshape(-1, 1)
        self._matches = Matches(self._old_frame, self.frame, np.hstack((idx, idx)))
        return self._matches

    def draw_tracks(self):
        """Overlay of the last tracks (klt.py:282-333); drawing only."""
        import cv2
        img = self.frame.image.copy()
        mask = np.zeros_like(img)
        feats = self.frame.features
        for i, (new, old) in enumerate(zip(feats.keypoints, self._old_frame.features.keypoints)):
            a, b = new.ravel()
            c, d = old.ravel()
            color = self._colors[feats.uids[i] % len(self._colors)].tolist()
            mask = cv2.line(mask, (int(a), int(b)), (int(c), int(d)), color, 2)
            img = cv2.circle(img, (int(a), int(b)), 5, color, -1)
        self._last_masks = (self._last_masks + [mask])[-10:]
        overlay = np.zeros_like(img)
        for m in self._last_masks:
            overlay = cv2.add(overlay, m)
        return cv2.addWeighted(img, 1, overlay, 0.5, 0), mask
