"""Keypoint matches between two frames (reference: src/vo/primitives/matches.py).

Construction re-orders the feature tables of both frames into four blocks
    [triangulated | already matched | newly matched | unmatched]
so that row i of frame1 and row i of frame2 are the same physical point for every matched row, and
carries landmarks, track starts and track poses from frame1 over to frame2."""
import numpy as np

__all__ = ["Matches"]


def _blocks(arr, idx_blocks):
    return np.concatenate([arr[i] for i in idx_blocks], axis=0)


class Matches:
    def __init__(self, frame1, frame2, matches: np.ndarray):
        self.frame1, self.frame2 = frame1, frame2
        self.newly_matched_idx = None
        self._threshold = 0.1
        f1, f2 = frame1.features, frame2.features
        i1, i2 = matches[:, 0], matches[:, 1]

        st = f1.state[i1]
        is_tri, is_old, is_new = st == 2, st == 1, st == 0
        tri1, old1, new1 = i1[is_tri], i1[is_old], i1[is_new]
        rest1 = np.delete(np.arange(len(f1.keypoints)), np.concatenate([tri1, old1, new1]))
        order1 = [tri1, old1, new1, rest1]
        n_tri, n_old, n_new = len(tri1), len(old1), len(new1)

        assert not np.any(np.isnan(f1.tracks[old1])), "NaN in matched tracks"
        kp1 = _blocks(f1.keypoints, order1)
        desc1 = None if f1.descriptors is None else _blocks(f1.descriptors, order1)
        land1 = _blocks(f1.landmarks, order1)
        state1 = np.concatenate((2 * np.ones_like(tri1), np.ones_like(old1), np.ones_like(new1), 0 * np.ones_like(rest1)))
        tracks1 = np.concatenate((np.full((n_tri, 2, 1), np.nan), f1.tracks[old1], f1.keypoints[new1],
                                  np.full((len(rest1), 2, 1), np.nan)))
        poses1 = np.concatenate((np.full((n_tri, 4, 4), np.nan), f1.poses[old1], f1.poses[new1],
                                 np.full((len(rest1), 4, 4), np.nan)))
        assert not np.any(np.isnan(land1[state1 == 2])), "NaN in triangulated landmarks"
        f1.keypoints, f1.state, f1.descriptors = kp1, state1, desc1
        f1.landmarks, f1.tracks, f1.poses = land1, tracks1, poses1

        tri2, old2, new2 = i2[is_tri], i2[is_old], i2[is_new]
        rest2 = np.delete(np.arange(len(f2.keypoints)), np.concatenate([tri2, old2, new2]))
        order2 = [tri2, old2, new2, rest2]
        kp2 = _blocks(f2.keypoints, order2)
        desc2 = None if f1.descriptors is None else _blocks(f2.descriptors, order2)
        # triangulated landmarks are inherited from frame1 (already in block order there)
        land2 = np.concatenate([f1.landmarks[:n_tri], f2.landmarks[old2], f2.landmarks[new2], f2.landmarks[rest2]], axis=0)
        state2 = np.concatenate((2 * np.ones_like(tri2), np.ones_like(old2), np.ones_like(new2), np.zeros_like(rest2)))
        a, b, c = n_tri, n_tri + n_old, n_tri + n_old + n_new
        tracks2 = np.concatenate((np.full((n_tri, 2, 1), np.nan), f1.tracks[a:b], f1.tracks[b:c], kp2[c:]))
        poses2 = np.concatenate((np.full((n_tri, 4, 4), np.nan), f1.poses[a:b], f1.poses[b:c],
                                 np.full((len(rest2), 4, 4), np.nan)))
        assert len(tracks2) == len(kp2), "Length of tracks and keypoints do not match"
        assert not np.any(np.isnan(land2[state2 == 2])), "NaN in triangulated landmarks"
        f2.keypoints, f2.state, f2.descriptors = kp2, state2, desc2
        f2.landmarks, f2.tracks, f2.poses = land2, tracks2, poses2
