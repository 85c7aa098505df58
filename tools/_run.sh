cd /root/repo
timeout 400 python -m pytest tests/test_gpu_transcribe.py tests/test_gpu_stages.py -x -q -m gpu 2>&1 | tail -3
timeout 300 bash tools/ab.sh 64 2 2>&1 | tee gpurun_out/r2p_ab.log
