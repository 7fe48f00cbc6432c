#!/bin/bash
# compute-sanitizer over a small extraction + stereo + matching run (both level paths): gpurun -- 'bash tools/san_round.sh'
mkdir -p gpurun_out
T='tests/test_gpu_extract.py::test_extract_stages_match_oracle tests/test_gpu_match.py::test_process_stereo_batch_matches_oracle'
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $T -x -q -k "240 or 280 or stereo_batch" > gpurun_out/r2_san_$tool.log 2>&1
  echo "== $tool: $(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2_san_$tool.log | tail -1) | $(tail -1 gpurun_out/r2_san_$tool.log)"
done
