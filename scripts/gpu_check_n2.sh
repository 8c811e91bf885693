set -x
python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --config c4 --no-cpu > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err; tail -c 1500 gpurun_out/bench_c4_n2.err
python bench.py --gpus 1 --steps 3 --config c4 > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err; tail -c 800 gpurun_out/bench_c4_n1.err
