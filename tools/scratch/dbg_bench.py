import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/visual-odometry-project_b200")
import bench
from vo import _native as nat
from vo.pipeline import Pipeline, DETECTOR_HARRIS
S, P = 3, 8
wd = bench.World(1234, S, P)
pool = wd.pool()
pl = Pipeline(S, bench.H, bench.W, bench.K_INTR, detector=DETECTOR_HARRIS, **bench.pipeline_kwargs())
pl.prime(pool[0], init_tables=True)
bench.start_tables(pl, S)
for i in range(1, 20):
    t = wd.time_index(i)
    summ = pl.step(pool[t])
    print(i, t, "true", bench.World.true_position(t)[:2].round(4), "est", summ["pose"][:, :, 3].round(4).tolist(), "N", summ["p3p_N"], "inl", summ["n_inliers"], "cand", summ["n_candidates"], "tri", summ["n_tri"], "rows", summ["n_rows"], "draws", summ["draws"], "flags", summ["flags"], "gn", summ["gn_iters"])
