"""The other BASELINE.json configs as bench workloads (same contract line as bench.py, one line each):

    python bench.py --workload p3p_sweep   # configs[2]: 3000 correspondences x 65536 hypotheses per frame
    python bench.py --workload stress      # configs[3]: 4096x2160 frames, 10k tracked keypoints, 4-level pyramid

They are stage-level measurements (inputs resident in HBM, CUDA events on the launching stream, inputs larger than L2 or
cycled); parity for both shapes is covered by tests/test_p3p_gpu.py::test_full_size_sweep_3000x65536 and
tests/test_stress_gpu.py."""
import ctypes as C
import json
import os

import numpy as np

import bench

FLOP_PER_POINT = 33        # reproj_err2 (p3p.py:81-108): 3x3 transform 9 mul + 9 add, 1 div, 4 mul + 2 add (projection), 2 sub, 2 mul + 1 add, + the compare


def _timed(fn, tstream, steps, warmup):
    import torch
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    for i in range(steps):
        fn(warmup + i)
    e1.record(tstream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_p3p_sweep(args, ctx, dev, rank, world, sampler, peaks):
    import torch
    from vo import _native as nat
    L = nat.lib()
    F, N, Hn = 8, 3000, 65536                                  # frames per launch
    rng = np.random.default_rng(5 + rank)
    K = bench.K_INTR
    land = rng.uniform(-12, 12, (F, N, 3)); land[..., 2] = rng.uniform(4, 60, (F, N))
    uv = np.einsum("ij,fnj->fni", K, land); uv = uv[..., :2] / uv[..., 2:] + rng.normal(0, 0.4, (F, N, 2))
    for f in range(F):
        out = rng.choice(N, int(0.4 * N), replace=False)
        uv[f, out] += rng.uniform(-100, 100, (len(out), 2))
    S = (np.argsort(rng.random((F, Hn, 16)), axis=2)[..., :4] + rng.integers(0, N - 16, (F, Hn, 1))).astype(np.int32)
    d_l, d_p, d_s = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (land, uv, S))
    d_m = torch.empty((F, Hn, 12), dtype=torch.float64, device=dev)
    d_v = torch.empty((F, Hn), dtype=torch.uint8, device=dev)
    d_c = torch.empty((F, Hn), dtype=torch.int32, device=dev)
    K9 = np.ascontiguousarray(K.reshape(9))
    tstream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(tstream)
    st = tstream.cuda_stream

    def step(i):
        nat.check(L.vo_p3p_ransac_score_dev(ctx.handle, d_l.data_ptr(), d_p.data_ptr(), F, N, nat.ptr(K9), d_s.data_ptr(), Hn,
                                            C.c_double(bench.P3P_THR), 0, d_m.data_ptr(), d_v.data_ptr(), d_c.data_ptr(), st), "p3p score")
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ms = _timed(step, tstream, args.steps, max(args.warmup, 3))
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    hyp_s = world * F * Hn / (ms / 1e3)
    g = C.c_double()
    nat.check(L.vo_test_dfma_peak(ctx.handle, C.byref(g)), "dfma peak")
    valid = float(d_v.float().mean().item())
    achieved = F * Hn * valid * N * FLOP_PER_POINT / (ms / 1e3) / 1e9
    return {"metric": "P3P-RANSAC hypotheses scored/s @3000 correspondences (solve + reprojection inlier count)", "value": hyp_s,
            "unit": "hypotheses/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[2]: {F} frames x {N} correspondences x {Hn} hypotheses per launch (40 % outliers)",
                       "l2_policy": f"models + samples + counts {F * Hn * 116 / 1e6:.0f} MB per launch, rewritten every launch",
                       "valid_hypotheses": valid},
            "gpu_launches": int(launches) * world,
            "roofline": {"kernel": "p3p_count_kernel (+ p3p_solve_kernel)", "bound": "fp64", "achieved": achieved, "peak": g.value,
                         "unit": "GFLOP/s", "frac": achieved / g.value, "traffic": None,
                         "peak_source": "DFMA rate measured in this run (vo_test_dfma_peak: 8 independent chains per thread); the kernel "
                                        "issues unfused DMUL / DADD for bit-exact parity with numpy, so its own ceiling is half of it",
                         "flop_per_hypothesis_point": FLOP_PER_POINT},
            "clocks": clocks}


def run_stress(args, ctx, dev, rank, world, sampler, peaks):
    import torch
    from vo import _native as nat
    L = nat.lib()
    Hs, Ws, Kp, F, P = 2160, 4096, 10000, 4, 3
    tex = bench.make_texture(77 + rank, 2304, 4352)
    pitch = (Ws + 15) & ~15
    pool = torch.zeros((P, F, Hs, pitch), dtype=torch.uint8, device=dev)
    for t in range(P):
        for f in range(F):
            pool[t, f, :, :Ws] = torch.from_numpy(np.ascontiguousarray(tex[8 * f + t:8 * f + t + Hs, 16 * f + 3 * t:16 * f + 3 * t + Ws])).to(dev)
    nl = C.c_int(); fb = C.c_size_t()
    lh, lw = (C.c_int * 8)(), (C.c_int * 8)(); lp, lo = (C.c_size_t * 8)(), (C.c_size_t * 8)()
    win, max_level = 21, 3
    nat.check(L.vo_klt_pyramid_layout(Hs, Ws, max_level, win, C.byref(nl), lh, lw, lp, lo, C.byref(fb)), "layout")
    pyr = [torch.zeros(F * fb.value, dtype=torch.uint8, device=dev) for _ in range(2)]
    resp = torch.empty((F, Hs, Ws), dtype=torch.float64, device=dev)
    kp = torch.zeros((F, Kp, 2), dtype=torch.int32, device=dev)
    pts = torch.zeros((F, Kp, 2), dtype=torch.float32, device=dev)
    nxt = torch.zeros((F, Kp, 2), dtype=torch.float32, device=dev)
    stt = torch.zeros((F, Kp), dtype=torch.uint8, device=dev)
    err = torch.zeros((F, Kp), dtype=torch.float32, device=dev)
    tstream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(tstream)
    st = tstream.cuda_stream
    state = {"cur": 0}

    def step(i):
        fr = pool[i % P]
        cur, nx = state["cur"], 1 - state["cur"]
        nat.check(L.vo_klt_build_pyramid_dev(ctx.handle, fr.data_ptr(), F, Hs, Ws, pitch, Hs * pitch, max_level, win, pyr[nx].data_ptr(), st), "pyramid")
        if i > 0:
            nat.check(L.vo_klt_track_dev(ctx.handle, pyr[cur].data_ptr(), pyr[nx].data_ptr(), F, Hs, Ws, max_level, win, 30, C.c_double(0.01),
                                         C.c_double(1e-4), pts.data_ptr(), Kp, nxt.data_ptr(), stt.data_ptr(), err.data_ptr(), st), "track")
        nat.check(L.vo_harris_detect_dev(ctx.handle, fr.data_ptr(), F, Hs, Ws, pitch, Hs * pitch, 9, C.c_double(0.09), 5, Kp,
                                         resp.data_ptr(), kp.data_ptr(), st), "harris")
        with torch.cuda.stream(tstream):
            pts.copy_(kp.to(torch.float32))
        state["cur"] = nx
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ms = _timed(step, tstream, args.steps, max(args.warmup, 3))
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    tracked = float(stt.float().mean().item())

    def harris_only(i):
        fr = pool[i % P]
        nat.check(L.vo_harris_response_dev(ctx.handle, fr.data_ptr(), F, Hs, Ws, pitch, Hs * pitch, 9, C.c_double(0.09), resp.data_ptr(), st), "harris")
    hms = _timed(harris_only, tstream, 10, 3)
    algo = F * Hs * Ws * 9
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = algo / (hms / 1e3) / 1e9
    return {"metric": "large-frame Harris + KLT frames/s @4096x2160 (10k keypoints, 4-level pyramid, win 21)", "value": world * F / (ms / 1e3),
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[3]: {F} frames of {Ws}x{Hs} per step: pyramid (4 levels), KLT of {Kp} points (win 21, 30 it., eps 0.01), "
                                   f"Harris response + NMS of {Kp} keypoints", "l2_policy": f"inputs larger than L2: {P} frame sets of {F * Hs * pitch / 1e6:.0f} MB "
                                   f"+ {F * Hs * Ws * 8 / 1e6:.0f} MB score maps per step", "tracked_fraction": tracked},
            "gpu_launches": int(launches) * world,
            "roofline": {"kernel": "harris_response", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "algorithmic_bytes_per_launch": algo, "launch_ms": hms,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s"},
            "clocks": clocks}


def run(args, ctx, dev, rank, world, sampler, peaks):
    return (run_p3p_sweep if args.workload == "p3p_sweep" else run_stress)(args, ctx, dev, rank, world, sampler, peaks)
