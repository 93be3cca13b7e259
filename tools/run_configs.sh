#!/bin/bash
# BASELINE.json configs as reproducible commands (run under gpurun; outputs in gpurun_out/, copy what is kept to profiles/):
#   config 2  aggregator k=18, 1 x B200        bench.py --k 18
#   config 3  aggregator k=20, 1 x B200        bench.py --k 20
#   config 4  aggregator k=22, N x B200        bench.py --gpus N            (the driver's own SCALE run)
#   config 5  MSM + NTT sweep 2^16..2^26, N x B200 vs the CPU port        tools/sweep.py
# usage: tools/run_configs.sh <tag> <n_gpus> [what: bench|sweep|all]
set -x
TAG=${1:-r02}
N=${2:-1}
WHAT=${3:-all}
mkdir -p gpurun_out
RUN="python"
if [ "$N" != "1" ]; then
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530"
fi
if [ "$WHAT" = "bench" ] || [ "$WHAT" = "all" ]; then
  for K in 18 20; do
    $RUN bench.py --gpus $N --k $K --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_k${K}_n${N}.json 2> gpurun_out/${TAG}_bench_k${K}_n${N}.err
  done
fi
if [ "$WHAT" = "sweep" ] || [ "$WHAT" = "all" ]; then
  $RUN tools/sweep.py --kmin 16 --kmax 26 > gpurun_out/${TAG}_sweep_n${N}.jsonl 2> gpurun_out/${TAG}_sweep_n${N}.err
fi
tail -c 600 gpurun_out/${TAG}_*_n${N}.err
