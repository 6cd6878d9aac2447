import sys, numpy as np, torch, ctypes as C
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/visual-odometry-project_b200")
import bench, cv2, os
from vo import _native as nat
ctx = nat.default_context(0); L = nat.lib()
wd = bench.World(1234, 8, 2)
frames = [wd.frame(s, 0) for s in range(8)]
k = cv2.imread("/root/repo/tests/golden/kitti05/000000.png", cv2.IMREAD_GRAYSCALE)
def run(img, K=1000):
    H, W = img.shape
    pitch = (W + 15) & ~15
    d = torch.zeros((H, pitch), dtype=torch.uint8, device="cuda"); d[:, :W] = torch.from_numpy(img).cuda()
    resp = torch.empty((H, W), dtype=torch.float64, device="cuda")
    kp = torch.empty((K, 2), dtype=torch.int32, device="cuda"); st = torch.zeros(4, dtype=torch.int32, device="cuda")
    nat.check(L.vo_harris_response_dev(ctx.handle, d.data_ptr(), 1, H, W, pitch, H*pitch, 9, C.c_double(0.09), resp.data_ptr(), None))
    nat.check(L.vo_harris_nms_dev(ctx.handle, resp.data_ptr(), 1, H, W, 5, K, kp.data_ptr(), st.data_ptr(), None))
    ctx.synchronize(); torch.cuda.synchronize()
    r = resp.cpu().numpy()
    return st.cpu().numpy(), (r > 0).mean()
for f in frames[:4]: print("synthetic", run(f))
print("kitti", run(k))
