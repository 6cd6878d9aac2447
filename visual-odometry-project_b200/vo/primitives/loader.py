"""Dataset iterator (reference: src/vo/primitives/loader.py): KITTI 05, Malaga 07, parking."""
import configparser
import glob
import os

import numpy as np

from vo.primitives.frame import Frame
from vo.sensors import Camera

__all__ = ["Sequence"]


class Sequence:
    """Iterates the frames of one dataset under `path` (relative to the project root unless absolute)."""

    project_name = "visual-odometry-project"

    def __init__(self, dataset: str, path: str = "./data", camera: int = 0, increment: int = 1,
                 rectified: bool = False, use_lowres: bool = False):
        self.dataset = dataset
        self._rel_data_path = path
        self.data_dir = self.get_data_dir()
        self.camera = camera
        self.intrinsics = None
        self.idx = 0
        self.increment = increment
        self._rectified = rectified
        self._use_lowres = use_lowres
        self.images = self._load()
        self.sensor = Camera(intrinsic_matrix=self.intrinsics)

    def get_data_dir(self):
        if os.path.isabs(self._rel_data_path):
            return self._rel_data_path
        here = os.path.dirname(os.path.realpath(__file__))
        parts = here.split(self.project_name)
        if len(parts) < 2:                       # not inside a checkout of that name: relative to the cwd
            return os.path.abspath(self._rel_data_path)
        root = parts[0] if len(parts) == 2 else os.path.join(parts[0], self.project_name)
        return os.path.join(root, self.project_name, self._rel_data_path)

    def _load(self):
        loaders = {"kitti": self._load_kitti, "malaga": self._load_malaga, "parking": self._load_parking}
        if self.dataset not in loaders:
            raise Exception("Invalid dataset")
        return loaders[self.dataset]()

    def _load_kitti(self):
        base = os.path.join(self.data_dir, "kitti", "05")
        paths = sorted(glob.glob(os.path.join(base, f"image_{self.camera}") + "/*.png"))
        with open(os.path.join(base, "calib.txt"), "r") as fh:
            line = fh.readlines()[2 * self.camera + 1]
        vals = [np.float32(v) for v in line.split(" ")[1:]]
        self.intrinsics = np.array(vals).reshape(3, 4)[:, :3]
        return paths

    def _load_malaga(self):
        base = os.path.join(self.data_dir, "malaga-urban-dataset-extract-07")
        side = "left" if self.camera == 0 else "right"
        paths = sorted(glob.glob(os.path.join(base, "Images") + f"/*{side}.jpg"))
        if not self._rectified:
            cfg_file, section = "camera_params_raw_1024x768.txt", f"CAMERA_PARAMS_{side.upper()}"
        else:
            res = "800x600" if self._use_lowres else "1024x768"
            cfg_file, section = f"camera_params_rectified_a=0_{res}.txt", f"CAMERA_{side.upper()}"
        cfg = configparser.ConfigParser()
        cfg.read(os.path.join(base, cfg_file))
        val = lambda k: np.float32(cfg[section][k].split("//")[0])  # noqa: E731
        self.intrinsics = np.array([[val("fx"), 0, val("cx")], [0, val("fy"), val("cy")], [0, 0, 1]])
        return paths

    def _load_parking(self):
        base = os.path.join(self.data_dir, "parking")
        paths = sorted(glob.glob(os.path.join(base, "images") + "/*.png"))
        with open(os.path.join(base, "K.txt"), "r") as fh:
            txt = fh.read().replace(" ", "").replace("\n", "")
        self.intrinsics = np.asarray(txt.split(",")).astype(np.float32).reshape(3, 3)
        return paths

    def get_frame(self, idx: int) -> Frame:
        import cv2
        frame = Frame(cv2.imread(self.images[idx]))
        frame.frame_id = idx
        frame.intrinsics = self.intrinsics
        frame.sensor = Camera(self.intrinsics)
        return frame

    def get_intrinsics(self) -> np.ndarray:
        return self.intrinsics

    def get_camera(self) -> Camera:
        return self.sensor

    def __len__(self) -> int:
        return len(self.images)

    def __next__(self) -> Frame:
        if self.idx >= len(self.images):
            raise StopIteration
        frame = self.get_frame(self.idx)
        self.idx += self.increment
        return frame

    def __iter__(self):
        return self

    def __repr__(self) -> str:
        return "Sequence(dataset={}, path={}, camera={}, increment={})".format(
            self.dataset, self._rel_data_path, self.camera, self.increment)
