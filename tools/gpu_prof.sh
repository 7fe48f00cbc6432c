# usage: bash tools/gpu_prof.sh <tag> <kernel-regex> ...   -> gpurun_out/<tag>_<kernel>.ncu-rep (one launch each, 32 pairs, 1 lane, no side stream)
tag=$1; shift
for k in "$@"; do
HYORB_LANES=1 HYORB_SIDE_BLUR=0 ncu --set full --clock-control none --import-source on -k regex:"^$k" --launch-skip 4 --launch-count 1 -o gpurun_out/${tag}_$k -f python bench.py --steps 3 --warmup 3 --no-cpu --pairs 32 > gpurun_out/prof_$k.log 2>&1
done
