"""Summarise an `ncu --set full` capture of the decode kernel into profiles/ (text for the judge, JSON for bench.py):

    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_mega -c 1 \
        -o gpurun_out/r2_mega -f python tools/profile_clip.py large-v3 steps          # on the GPU box (16 decode steps per launch)
    python tools/ncu_summary.py gpurun_out/r2_mega.ncu-rep profiles/r2_ncu_decode_mega 16 large-v3    # here

The JSON carries the sha256 of csrc/decoder_mega.cu at summary time: bench.py reports `roofline.traffic` from it only while the
kernel source is unchanged."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, out, steps, shape = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(name):
    v, u = m[name]
    x = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6,
             "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
    return x * scale


keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]
src = os.path.join(ROOT, "speaksense_b200", "csrc", "decoder_mega.cu")
sha = hashlib.sha256(open(src, "rb").read()).hexdigest()[:16]
dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
j = {"kernel": m.get("Kernel Name", ("decode_mega_kernel", ""))[0], "shape": shape, "steps_per_launch": steps, "kernel_src_sha16": sha,
     "duration_ms_under_ncu": num("gpu__time_duration.sum") * 1e3, "dram_bytes_per_launch": dram, "dram_bytes_per_step": dram / steps}
with open(out + ".txt", "w") as f:
    f.write("ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_mega -c 1 python tools/profile_clip.py %s steps"
            "   (one launch = %d decode steps incl. LM head; csrc/decoder_mega.cu sha256[:16] = %s)\n----\n" % (shape, steps, sha))
    for k in keys:
        if k in m:
            f.write("%s = %s %s\n" % (k, m[k][0], m[k][1]))
            j[k] = m[k][0] + " " + m[k][1]
    f.write("dram bytes per decode step = %.4f GB\n" % (dram / steps / 1e9))
json.dump(j, open(out + ".json", "w"), indent=1)
print(open(out + ".txt").read())
