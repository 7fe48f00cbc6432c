# usage: bash tools/gpu_prof.sh <tag> <kernel-regex>[:launches] ...  -> gpurun_out/<tag>_<kernel>.ncu-rep
# one `ncu --set full` capture per kernel of the bench's own step (128 pairs = 256 images per launch, 1 lane, no side stream);
# "k_level:8" captures the eight per-level launches of one step
tag=$1; shift
for spec in "$@"; do
k=${spec%%:*}; n=1; [[ "$spec" == *:* ]] && n=${spec##*:}
HYORB_LANES=1 HYORB_SIDE_BLUR=0 ncu --set full --clock-control none --import-source on -k regex:"^$k" --launch-skip $((4 * n)) --launch-count $n -o gpurun_out/${tag}_$k -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/prof_$k.log 2>&1
done
