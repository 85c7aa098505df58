cd "$(dirname "$0")/.."
python -c "
import sys; sys.path.insert(0,'.')
from speaksense_b200 import synth; import os
synth.ensure_model('/tmp/ss_models/ggml-large-v3-peaked-s0.bin', shape='large-v3', family='peaked', seed=0)" 2>/dev/null
for cfg in "A=1" "SS_ENC_PDL=0" "SS_ENC_PDL=0 SS_ENC_GRAPH=0" "A=2" "SS_ENC_PDL=0"; do
  echo "== stream8 $cfg" | tee -a gpurun_out/s13_stream.txt
  env $cfg timeout 150 python tools/stream_bench.py large-v3 120 0 8 1 2>/dev/null | head -1 | cut -c150-420 | tee -a gpurun_out/s13_stream.txt
done
for cfg in "A=1" "SS_ENC_PDL=0" "SS_ENC_PDL=0 SS_ENC_GRAPH=0"; do
  echo "== beam5 $cfg" | tee -a gpurun_out/s13_stream.txt
  env $cfg timeout 150 python tools/stream_bench.py large-v3 60 5 2>/dev/null | head -1 | cut -c150-420 | tee -a gpurun_out/s13_stream.txt
done
