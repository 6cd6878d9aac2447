"""SIFT detector / matcher (reference: src/vo/features/sift.py).  OpenCV on the host, exactly as the
reference; not part of the accelerated path."""
import numpy as np

from vo.primitives import Features, Frame, Matches

__all__ = ["SIFTDetector"]


class SIFTDetector:
    def __init__(self, frame: Frame):
        import cv2
        self.sift = cv2.SIFT_create()
        kp, desc = self.detect_and_compute(frame=frame)
        frame.features = Features(keypoints=kp)
        frame.features.descriptors = desc

    def detect_and_compute(self, frame: Frame):
        raw, desc = self.sift.detectAndCompute(frame.image, None)
        return np.array([k.pt for k in raw]).reshape(-1, 2, 1), np.array(desc)

    def get_sift_matches(self, curr_frame: Frame, new_frame: Frame) -> Matches:
        """Detect in the new frame and ratio-test match against the current one (sift.py:23-57)."""
        import cv2
        kp2, desc2 = self.detect_and_compute(new_frame)
        new_frame.features = Features(kp2)
        new_frame.features.descriptors = desc2
        used = np.zeros(len(desc2))
        good = []
        for m, n in cv2.BFMatcher().knnMatch(curr_frame.features.descriptors, desc2, k=2):
            if m.distance < 0.8 * n.distance and used[m.trainIdx] == 0:
                good.append([m.queryIdx, m.trainIdx])
                used[m.trainIdx] = 1
        return Matches(curr_frame, new_frame, matches=np.array(good))
