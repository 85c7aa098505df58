cd "$(dirname "$0")/.."
bash tools/ab.sh 64 2 > gpurun_out/s12_ab.txt 2>&1; cat gpurun_out/s12_ab.txt
