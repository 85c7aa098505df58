cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/s6_tests.txt
SS_NO_PROF=1 python tools/mega_prof.py large-v3 64 2>&1 | grep -E "ms/step|hash" | tee gpurun_out/s6_mega.txt
for B in 4 8 16 32; do timeout 100 python tools/batch_bench.py large-v3 $B 2 1 2>&1 | tee -a gpurun_out/s6_batch.json | cut -c1-400; done
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:bd_ -c 300 --csv --log-file gpurun_out/s6_batch_launches.csv python tools/batch_bench.py large-v3-l2 32 0 1 > gpurun_out/s6_ncu_batch.log 2>&1; tail -1 gpurun_out/s6_ncu_batch.log | cut -c1-200
