#!/bin/bash
# bench.py at N GPUs of one box (N = $1), launched exactly as the driver does.
N=${1:-2}
cd "$(dirname "$0")/.."
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err
tail -c 600 gpurun_out/bench_c4_n$N.err
