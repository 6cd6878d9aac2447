#!/bin/bash
# Run on the GPU box (gpurun): everything profiles/ holds for a round, into gpurun_out/<tag>_*.
# usage: bash tools/refresh_all.sh [tag]
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_gputests.log 2>&1; tail -2 gpurun_out/${tag}_gputests.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || { tail -5 gpurun_out/${tag}_bench.err; exit 1; }
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python bench.py --workload p3p_sweep > gpurun_out/${tag}_bench_p3p_sweep.json 2> gpurun_out/${tag}_bench_p3p.err
python bench.py --workload stress > gpurun_out/${tag}_bench_stress.json 2> gpurun_out/${tag}_bench_stress.err
# launch lists (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-sequence --no-extras --e2e-steps 1 > gpurun_out/${tag}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_s1_launches.csv \
    python bench.py --steps 2 --warmup 3 --seqs 1 --no-cpu-baseline --no-single-sequence --no-extras --e2e-steps 1 > gpurun_out/${tag}_s1_ncu.log 2>&1
# full captures of the top kernels (one launch each)
ncu --set full --clock-control none --import-source on -k regex:harris_response_fast -c 1 -f -o gpurun_out/${tag}_response \
    python tools/profile_kernel.py harris 148 > gpurun_out/${tag}_response.log 2>&1
for k in klt_track_packed harris_nms_bands harris_localmax harris_nms_scan pipe_pose_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/${tag}_$k \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-sequence --no-extras --e2e-steps 1 > gpurun_out/${tag}_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:bootstrap_kernel -c 1 -f -o gpurun_out/${tag}_bootstrap_kernel \
    python -m pytest tests/test_bootstrap_gpu.py -q -m gpu -k batched > gpurun_out/${tag}_bootstrap.log 2>&1
ls -la gpurun_out/${tag}_*
