"""3D point-cloud viewer (reference: src/vo/visualization/point_cloud.py).  Presentation only; needs
matplotlib + pytransform3d, which are imported lazily so that `import vo` works without them."""
import numpy as np

__all__ = ["PointCloudVisualizer"]


class PointCloudVisualizer:
    def __init__(self) -> None:
        import matplotlib.pyplot as plt
        import pytransform3d.plot_utils as pu
        self._plt = plt
        plt.ion()
        self.fig = plt.figure(figsize=(5, 5))
        self.ax = pu.make_3d_axis(ax_s=1, unit="m")
        self.x, self.y, self.z = np.zeros((1, 1)), np.zeros((1, 1)), np.zeros((1, 1))
        plt.show()

    def _draw(self, percentiles=[5, 95]) -> None:
        self.ax.view_init(elev=-70, azim=-80, roll=0)
        self.ax.set_xlim(np.percentile(self.x, percentiles))
        self.ax.set_ylim(np.percentile(self.y, percentiles))
        self.ax.set_zlim(np.percentile(self.z, percentiles))
        self._plt.draw()

    def visualize_points(self, points: np.ndarray, color="b") -> None:
        xs, ys, zs = points[:, 0].flatten(), points[:, 1].flatten(), points[:, 2].flatten()
        self.ax.scatter(xs, ys, zs, color=color, alpha=0.1)
        self.x, self.y, self.z = np.append(self.x, xs), np.append(self.y, ys), np.append(self.z, zs)
        self._draw()

    def visualize_camera(self, camera) -> None:
        import pytransform3d.camera as pc
        import pytransform3d.transformations as pt
        pose = np.linalg.inv(camera.c_T_w)
        pt.plot_transform(self.ax, pose, s=0.1)
        pc.plot_camera(self.ax, cam2world=pose, M=camera.intrinsic_matrix, virtual_image_distance=0.1,
                       sensor_size=(480, 320))
        self._draw()
