"""Image overlays used by src/main.py (reference: src/vo/visualization/overlays.py).  Presentation
only -- outside the accelerated path; thin OpenCV drawing with the reference's signatures."""
import time

import numpy as np

__all__ = ["display_fps", "display_keypoints_info", "draw_keypoints", "draw_lines", "plot_matches", "plot_keypoints"]


def _text(image, text, org):
    import cv2
    cv2.putText(image, text, org, cv2.FONT_HERSHEY_SIMPLEX, 0.5, (255, 255, 255), 1, cv2.LINE_AA)
    return image


def display_fps(image, start_time: float, fps_queue):
    fps_queue.append(1 / (time.time() - start_time))
    return _text(image, f"FPS: {sum(fps_queue) / len(fps_queue):.2f}", (10, 20)), fps_queue


def display_keypoints_info(image, features):
    info = (f"Keypoints: {features.length}, Matched: {np.sum(features.state == 1)}, "
            f"Triangulated: {np.sum(features.state == 2)}")
    return _text(image, info, (10, 30))


def draw_keypoints(img, keypoints, colors):
    import cv2
    if isinstance(colors, tuple):
        colors = [colors] * len(keypoints)
    for p, c in zip(keypoints.reshape(-1, 2), colors):
        img = cv2.circle(img, p.astype(int), radius=2, color=c, thickness=2)
    return img


def draw_lines(img, start_points, end_points, colors):
    import cv2
    if isinstance(colors, tuple):
        colors = [colors] * len(start_points)
    for a, b, c in zip(start_points.reshape(-1, 2), end_points.reshape(-1, 2), colors):
        img = cv2.line(img, a.astype(int), b.astype(int), color=c, thickness=2)
    return img


def _rgb(img):
    import cv2
    return cv2.cvtColor(img, cv2.COLOR_GRAY2RGB) if (img.ndim == 2 or img.shape[-1] == 1) else img.copy()


def plot_matches(matches):
    import cv2
    w = matches.frame1.image.shape[1]
    kp1 = matches.frame1.features.matched_inliers_keypoints[:25]
    kp2 = matches.frame2.features.matched_inliers_keypoints[:25] + np.array([w, 0]).reshape(1, 2, 1)
    img = cv2.hconcat([_rgb(matches.frame1.image), _rgb(matches.frame2.image)])
    colors = list(map(tuple, (np.random.rand(len(kp1), 3) * 255).astype(int).tolist()))
    return draw_lines(draw_keypoints(draw_keypoints(img, kp1, colors), kp2, colors), kp1, kp2, colors)


def plot_keypoints(image, features, show_tracks=True):
    img = _rgb(image)
    img = draw_keypoints(img, features.keypoints[features.state == 0], (255, 0, 0))
    img = draw_keypoints(img, features.keypoints[features.state == 1], (0, 255, 255))
    img = draw_keypoints(img, features.keypoints[features.state == 2], (0, 255, 0))
    if show_tracks:
        sel = features.state == 1
        img = draw_lines(img, features.tracks[sel], features.keypoints[sel], (0, 255, 255))
    return img
