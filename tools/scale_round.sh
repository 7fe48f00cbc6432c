#!/bin/bash
# One multi-GPU box visit: gpurun --gpus N -- 'bash tools/scale_round.sh N tag'  ->  gpurun_out/<tag>_{pcie,bench,c5}_nN.json
N=${1:-8}; tag=${2:-r2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
mkdir -p gpurun_out
$TR tools/pcie_bw.py 2>/dev/null | tail -1 > gpurun_out/${tag}_pcie_n$N.json
$TR bench.py --gpus $N --no-cpu 2> gpurun_out/${tag}_bench_n$N.err | tail -1 > gpurun_out/${tag}_bench_n$N.json
$TR tools/run_sequence.py --frames 10000 2>/dev/null | tail -1 > gpurun_out/${tag}_c5_n$N.json
nproc > gpurun_out/${tag}_nproc_n$N.txt; free -g | head -2 >> gpurun_out/${tag}_nproc_n$N.txt
cut -c1-400 gpurun_out/${tag}_pcie_n$N.json; cut -c1-300 gpurun_out/${tag}_c5_n$N.json
