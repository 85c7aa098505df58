cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_multi_device.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s15_tests_2gpu.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s15_bench_2gpu.json 2> gpurun_out/s15_bench_2gpu.err; echo "bench rc=$?"; tail -1 gpurun_out/s15_bench_2gpu.json | cut -c1-400
