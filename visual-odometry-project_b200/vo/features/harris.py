"""Harris corner detector + patch descriptors (reference: src/vo/features/harris.py).

extractKeypoints (Sobel -> structure tensor -> 9x9 box sums -> response -> greedy NMS) and
extractDescriptors run on the GPU through vo_harris_detect_host; the float64 response map and the
keypoint order are bit-identical to the reference's.  matchDescriptor keeps the reference's
OpenCV brute-force ratio-test matcher (host; not part of the accelerated path yet)."""
import numpy as np

from vo import _ops
from vo.primitives import Features, Frame, Matches

__all__ = ["HarrisCornerDetector"]


def _gray(image):
    if image.ndim == 2:
        return image
    import cv2
    return cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)


class HarrisCornerDetector:
    def __init__(self, frame: Frame = None, patch_size: int = 9, kappa: float = 0.09, num_keypoints: int = 1000,
                 nonmaximum_supression_radius: int = 5, descriptor_radius: int = 9, match_lambda: float = 4.0):
        self._frame1 = frame
        self._frame2 = frame
        self._patch_size = patch_size
        self._kappa = kappa
        self._num_keypoints = num_keypoints
        self._nonmaximum_supression_radius = nonmaximum_supression_radius
        self._descriptor_radius = descriptor_radius
        self._match_lambda = match_lambda

    @property
    def img1_gray(self) -> np.ndarray:
        return _gray(self._frame1.image)

    @property
    def img2_gray(self) -> np.ndarray:
        return _gray(self._frame2.image)

    def featureMatcher(self, curr_frame: Frame, new_frame: Frame) -> Matches:
        """Detect + describe in the new frame (and in the current one if it has no features yet),
        then match (harris.py:50-84)."""
        self._frame1, self._frame2 = curr_frame, new_frame
        self._frame1.image = self.img1_gray
        self._frame2.image = self.img2_gray
        if self._frame1.features is None:
            self._frame1 = self.extractDescriptors(self.extractKeypoints(self._frame1))
        self._frame2 = self.extractDescriptors(self.extractKeypoints(self._frame2))
        return self.matchDescriptor(self._frame1, self._frame2)

    def extractKeypoints(self, frame: Frame) -> Frame:
        """The num_keypoints strongest Harris corners under greedy box suppression (harris.py:86-158)."""
        kp, _, _ = _ops.harris_detect(_as_u8(frame.image), self._num_keypoints, self._patch_size, self._kappa,
                                      self._nonmaximum_supression_radius)
        assert frame.features is None, "Frame already has features"
        frame.features = Features(kp.astype(np.float64).reshape(-1, 2, 1))
        return frame

    def extractDescriptors(self, frame: Frame) -> Frame:
        """(2r+1)^2 raw-pixel patch per keypoint from the zero-padded image (harris.py:160-194)."""
        kp = frame.features.keypoints.reshape(-1, 2).astype(np.int32)
        desc = _ops.harris_descriptors(_as_u8(frame.image), kp, self._descriptor_radius)
        frame.features.descriptors = desc.astype(np.float64).reshape(kp.shape[0], -1, 1)
        return frame

    def matchDescriptor(self, frame1: Frame, frame2: Frame) -> Matches:
        """2-NN brute-force matching with a 0.85 ratio test, one match per train descriptor
        (harris.py:196-264).  8-bit descriptors (what extractDescriptors produces) are matched on the GPU
        with exact integer distances; anything else takes the reference's OpenCV path on the host."""
        d1, d2 = frame1.features.descriptors, frame2.features.descriptors
        if _is_u8_valued(d1) and _is_u8_valued(d2) and len(d2) >= 2:
            pairs = _ops.match_descriptors(d1.reshape(len(d1), -1), d2.reshape(len(d2), -1), 0.85)
            if len(pairs) == 0:
                pairs = np.empty(shape=(0, 2), dtype=int)
            return Matches(frame1, frame2, pairs)
        import cv2
        f1, f2 = d1.astype(np.float32), d2.astype(np.float32)
        used = np.zeros(len(f1))
        good = []
        for m, n in cv2.BFMatcher().knnMatch(f1, f2, k=2):
            if m.distance < 0.85 * n.distance and used[m.trainIdx] == 0:
                good.append([m.queryIdx, m.trainIdx])
                used[m.trainIdx] = 1
        pairs = np.stack(good) if len(good) > 0 else np.empty(shape=(0, 2), dtype=int)
        return Matches(frame1, frame2, pairs)


def _is_u8_valued(d):
    d = np.asarray(d)
    if d.dtype == np.uint8:
        return True
    return bool(np.all((d >= 0) & (d <= 255) & (d == np.floor(d))))


def _as_u8(image):
    img = np.asarray(image)
    if img.ndim == 3 and img.shape[-1] == 1:
        img = img[..., 0]
    if img.dtype != np.uint8:
        raise TypeError("HarrisCornerDetector: the CUDA path takes 8-bit grayscale images")
    return img
