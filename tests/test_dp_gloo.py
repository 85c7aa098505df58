"""N>1 host logic on CPU: world_size-2 gloo group (the GPU path uses the same code over NCCL).  Units are
sharded contiguously with no data-path collective; the ncclUniqueId is ferried from rank 0; results are
gathered in clip order."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from speaksense_b200 import dp
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
clips = ["clip%%03d" %% i for i in range(257)]
mine = dp.shard(clips, r, w)
lo, hi = dp.shard_bounds(len(clips), r, w)
assert mine == clips[lo:hi]
token = dp.broadcast_bytes(bytes(range(128)) if r == 0 else None, 0)
assert token == bytes(range(128))
res = dp.gather_results([c.upper() for c in mine], 0)
if r == 0:
    assert res == [c.upper() for c in clips], "gather order"
    print("OK", w, len(res))
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_cover_everything():
    from speaksense_b200 import dp
    for n in (0, 1, 7, 8, 256, 257):
        for w in (1, 2, 3, 8):
            spans = [dp.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert dp.shard_bounds(256, 3, 8) == (96, 128)      # BASELINE config 4: 32 clips per GPU


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    port = _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK 2 257" in out.stdout
