import cProfile, pstats, sys, os, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-odometry-project_b200")); sys.path.insert(0, ROOT)
import bench
wd = bench.World(1234, 1, 8)
bench.class_api_loop(wd, 6)
pr = cProfile.Profile(); pr.enable()
fps, err = bench.class_api_loop(wd, 30)
pr.disable()
print("fps", fps, "err", err)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
