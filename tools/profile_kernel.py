"""Run one stage of the front end a few times on synthetic frames (for ncu captures).
usage: python tools/profile_kernel.py {harris|nms|klt|p3p|step} [S]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-odometry-project_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vo import _native as nat  # noqa: E402
from vo.frontend import Frontend  # noqa: E402

if os.environ.get("VO_LIB"):          # development: time another build of the library side by side
    nat.LIB_PATH = os.path.abspath(os.environ["VO_LIB"])
what = sys.argv[1] if len(sys.argv) > 1 else "harris"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 148
H, W = bench.H, bench.W
dev = torch.device("cuda", 0)
ctx = nat.Context(0)
L = nat.lib()
tex = bench.make_texture(1234)
pitch = (W + 15) & ~15
P = 3
pool = torch.zeros((P, S, H, pitch), dtype=torch.uint8, device=dev)
for t in range(P):
    fr = np.stack([bench.frame_from_texture(tex, s, t) for s in range(S)])
    pool[t, :, :, :W] = torch.from_numpy(fr).to(dev)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
stream = st.cuda_stream
resp = torch.empty((S, H, W), dtype=torch.float64, device=dev)
kp = torch.empty((S, 1000, 2), dtype=torch.int32, device=dev)
if what == "harris_time":
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e0.record(st)
        for i in range(20):
            fr = pool[i % P]
            nat.check(L.vo_harris_response_dev(ctx.handle, fr.data_ptr(), S, H, W, pitch, H * pitch, 9, C.c_double(0.09),
                                               resp.data_ptr(), stream), "resp")
        e1.record(st)
        torch.cuda.synchronize()
    print("harris_response ms/launch: %.4f" % (e0.elapsed_time(e1) / 20))
elif what == "nms_time":
    resps = []
    for t in range(P):
        rr = torch.empty((S, H, W), dtype=torch.float64, device=dev)
        nat.check(L.vo_harris_response_dev(ctx.handle, pool[t].data_ptr(), S, H, W, pitch, H * pitch, 9, C.c_double(0.09),
                                           rr.data_ptr(), stream), "resp")
        resps.append(rr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e0.record(st)
        for i in range(12):
            nat.check(L.vo_harris_nms_dev(ctx.handle, resps[i % P].data_ptr(), S, H, W, 5, 1000, kp.data_ptr(), None, stream), "nms")
        e1.record(st)
        torch.cuda.synchronize()
    print("harris_nms ms/call: %.4f" % (e0.elapsed_time(e1) / 12))
elif what in ("harris", "nms"):
    for i in range(5):
        fr = pool[i % P]
        nat.check(L.vo_harris_response_dev(ctx.handle, fr.data_ptr(), S, H, W, pitch, H * pitch, 9, C.c_double(0.09),
                                           resp.data_ptr(), stream), "resp")
        if what == "nms":
            nat.check(L.vo_harris_nms_dev(ctx.handle, resp.data_ptr(), S, H, W, 5, 1000, kp.data_ptr(), None, stream), "nms")
else:
    N = T = 1000
    Hn = 512
    geo = bench.make_geometry(99, S, N, T, Hn)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in geo.items()}
    table = torch.from_numpy(bench.iters_table(N, bench.P3P_CONF, bench.P3P_MAX_ITER)).to(dev)
    K9 = np.ascontiguousarray(bench.K_INTR.reshape(9))
    fe = Frontend(S, H, W, n_corr=N, n_hyp=Hn, p3p_threshold=bench.P3P_THR, n_tri=T, ctx=ctx)
    for i in range(5):
        fr = pool[i % P]
        fe.step_dev(fr.data_ptr(), pitch, H * pitch, d["landmarks"].data_ptr(), d["kp2d"].data_ptr(), K9,
                    d["samples"].data_ptr(), table.data_ptr(), bench.initial_iters(), d["tri_p1"].data_ptr(),
                    d["tri_p2"].data_ptr(), d["tri_proj1"].data_ptr(), d["tri_proj2"].data_ptr(), stream)
torch.cuda.synchronize()
print("done", what, S)
