#!/bin/bash
# Run on the GPU box (gpurun): one `ncu --set full` capture of the roofline kernel, raw metrics as csv.
tag=${1:-r01}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:harris_response_fast -c 1 -f -o gpurun_out/${tag}_response \
    python tools/profile_kernel.py harris 148 > gpurun_out/${tag}_response.log 2>&1
tail -2 gpurun_out/${tag}_response.log
