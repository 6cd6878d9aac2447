#!/bin/bash
# Run on the GPU box (gpurun): the trimmed refresh used for the last state of round 2 (the response / local-maximum /
# scan / pose / bootstrap kernels did not change after tools/refresh_all.sh ran; the tracker and the band kernel did).
# usage: bash tools/refresh_final.sh [tag]
tag=${1:-r02f}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_gputests.log 2>&1; tail -2 gpurun_out/${tag}_gputests.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || { tail -5 gpurun_out/${tag}_bench.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-sequence --no-extras --e2e-steps 1 > gpurun_out/${tag}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt
for k in harris_nms_bands klt_track_packed; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/${tag}_$k \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-sequence --no-extras --e2e-steps 1 > gpurun_out/${tag}_$k.log 2>&1
done
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python bench.py --workload p3p_sweep > gpurun_out/${tag}_bench_p3p_sweep.json 2> gpurun_out/${tag}_bench_p3p.err
python bench.py --workload stress > gpurun_out/${tag}_bench_stress.json 2> gpurun_out/${tag}_bench_stress.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_s1_launches.csv \
    python bench.py --steps 2 --warmup 3 --seqs 1 --no-cpu-baseline --no-single-sequence --no-extras --e2e-steps 1 > gpurun_out/${tag}_s1_ncu.log 2>&1
ls -la gpurun_out/${tag}_*
