#!/bin/bash
# Everything profiles/ needs from one GPU box, in one gpurun call:
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh r2f'      then, back here:   bash tools/refresh_profiles.sh r2f
# (numbers come from the plain runs; the ncu passes only produce the launch list and the per-kernel captures)
tag=${1:-r2f}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
HYORB_LANES=1 HYORB_SIDE_BLUR=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${tag}_launches.log 2>&1
bash tools/gpu_prof.sh $tag k_fast k_level:8 k_describe k_quadtree k_stereo_search
tail -1 gpurun_out/${tag}_bench.json | cut -c1-600
