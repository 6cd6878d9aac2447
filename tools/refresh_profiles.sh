#!/bin/bash
# Run on the GPU box (gpurun): the bench line and the ncu launch list of the same command, into gpurun_out/.
# usage: bash tools/refresh_profiles.sh [tag]     (copy gpurun_out/<tag>_* into profiles/ afterwards)
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 60 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || exit 1
tail -c 600 gpurun_out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-sequence --e2e-steps 1 > gpurun_out/${tag}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt
cat gpurun_out/${tag}_launches_summary.txt
