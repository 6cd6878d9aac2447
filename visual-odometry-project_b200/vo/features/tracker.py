"""Tracker front (reference: src/vo/features/tracker.py): picks KLT, Harris or SIFT by name."""
from vo.features.harris import HarrisCornerDetector
from vo.features.klt import KLTTracker
from vo.features.sift import SIFTDetector
from vo.primitives import Frame, Matches

__all__ = ["Tracker"]


class Tracker:
    _KINDS = {"klt": (KLTTracker, "track_features"), "harris": (HarrisCornerDetector, "featureMatcher"),
              "sift": (SIFTDetector, "get_sift_matches")}

    def __init__(self, frame, mode="klt"):
        self._init_frame = frame
        self._mode = mode
        self._tracker = None
        self.initTracker(frame)

    def initTracker(self, frame: Frame) -> None:
        if self._mode not in self._KINDS:
            raise Exception("Tracker Name not valid")
        self._tracker = self._KINDS[self._mode][0](frame)

    def trackFeatures(self, curr_frame: Frame, new_frame: Frame) -> Matches:
        if self._mode not in self._KINDS:
            raise Exception("Tracker Name not valid")
        return getattr(self._tracker, self._KINDS[self._mode][1])(curr_frame, new_frame)
