"""One image of a sequence plus what is known about it (reference: src/vo/primitives/frame.py)."""
import numpy as np

__all__ = ["Frame"]


class Frame:
    def __init__(self, image: np.ndarray, features=None, sensor=None, intrinsics: np.ndarray = None):
        self.image = image
        self.frame_id = None
        self.features = features
        self.intrinsics = intrinsics
        self.sensor = sensor

    def get_frame_id(self) -> int:
        return self.frame_id

    def get_intrinsics(self) -> np.ndarray:
        return self.intrinsics

    def show(self, other=None) -> None:
        """Display the image, optionally beside another frame (frame.py:41-60); needs a GUI OpenCV."""
        import cv2 as cv
        if other is None:
            cv.imshow("Frame {}".format(self.frame_id), self.image)
        else:
            pad = dict(top=20, bottom=20, left=20, right=20, borderType=cv.BORDER_CONSTANT, value=(255, 255, 255))
            both = np.hstack((cv.copyMakeBorder(other.image, **pad), cv.copyMakeBorder(self.image, **pad)))
            cv.imshow("Frame {} and {}".format(other.frame_id, self.frame_id), both)
        if cv.waitKey(30) & 0xFF == 27:
            cv.destroyAllWindows()

    def __repr__(self) -> str:
        return "Frame id: {}".format(self.frame_id)
