cd /root/repo
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2n_bench_8gpu.json 2> gpurun_out/r2n_bench_8gpu.err; echo rc=$?; cut -c1-250 gpurun_out/r2n_bench_8gpu.json; python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_8gpu.json')); print(json.dumps(d['next_rows'].get('dp256'))[:900])"
timeout 300 python tools/stream_bench.py large-v3 300 5 8 0 8 > gpurun_out/r2n_stream8_beam5_8gpu.json 2> gpurun_out/r2n_stream.err; cat gpurun_out/r2n_stream8_beam5_8gpu.json | cut -c1-500
timeout 300 python tools/stream_bench.py large-v3 300 0 8 0 8 > gpurun_out/r2n_stream8_greedy_8gpu.json 2>> gpurun_out/r2n_stream.err; cat gpurun_out/r2n_stream8_greedy_8gpu.json | cut -c1-500
timeout 200 python -m pytest tests/test_gpu_multi_device.py -x -q -m gpu 2>&1 | tail -2
tail -3 gpurun_out/r2n_stream.err
