"""`vo` -- drop-in for the reference package of the same name (saegsali/visual-odometry-project,
src/vo), with the data-parallel front end running as hand-written CUDA on a B200 through
libvo_b200.so (include/vo_b200.h).  GPU-backed: HarrisCornerDetector.extractKeypoints /
extractDescriptors, KLTTracker.track_features (pyramidal LK), P3PPoseEstimator.estimate_pose
(P3P + RANSAC, use_opencv=False semantics), LandmarksTriangulator._linear_triangulation /
triangulate_candidates.  Everything else is host bookkeeping, as in the reference."""
