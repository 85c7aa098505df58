"""Critical path of one decode step from per-CTA timestamps (needs a -DSS_MEGA_TRACE=1 build of decoder_mega.cu:
`tools/build_variant.sh t_trace speaksense_b200/csrc/decoder_mega.cu -DSS_MEGA_TRACE=1`, copied over the library).
Thread 0 of every CTA stamps %globaltimer when a phase's input is complete ("in") and when its outputs are published ("out").
Per phase: hop = last "in" of this phase - last "out" of the previous phase (exchange latency seen by the slowest reader),
work = last "out" - last "in".   usage: python tools/mega_trace.py [shape] [out.txt]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
raw = os.environ.setdefault("SS_MEGA_TRACE", "/tmp/ss_mega_trace.bin")
from speaksense_b200 import AsrParams, WhisperAsr, synth  # noqa: E402
shape = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
path = os.path.join(os.environ.get("SS_MODEL_DIR", "/tmp/ss_models"), "ggml-%s-peaked-s0.bin" % shape)
synth.ensure_model(path, shape=shape, family="peaked", seed=0)
eng = WhisperAsr(path)
st = eng.create_state()
eng.upload_pcm(st, synth.synth_audio(seed=1234))
eng.transcribe_resident(st, AsrParams(language=None if shape.endswith(".en") else "en", stream_mode=True))
ms = eng.bench_decode_steps(st, 64, 0)
t = np.fromfile(raw, dtype=np.int64).reshape(-1, 1024)
grid = t.shape[0]
names = ["QKV", "self", "O", "CQ", "cross", "fold", "CO", "FC1", "FC2"]
NP = len(names)
L = 0
while L < 56 and (t[:, (L * NP) * 2] != 0).any():
    L += 1
ev = t[:, :L * NP * 2].reshape(grid, L, NP, 2).astype(np.float64)
ev[ev == 0] = np.nan
out = []
out.append("ms/step (trace build) %.4f, grid %d, layers %d; times in ns (globaltimer); resolution: %s ns" % (
    ms, grid, L, np.unique(np.diff(np.unique(t[t != 0])))[:4]))
hop = np.zeros((L, NP)); work = np.zeros((L, NP)); first_in = np.zeros((L, NP)); spread_out = np.zeros((L, NP)); n_part = np.zeros(NP)
prev_out = None
for l in range(L):
    for p in range(NP):
        tin = np.nanmax(ev[:, l, p, 0]); tout = np.nanmax(ev[:, l, p, 1])
        n_part[p] = np.count_nonzero(~np.isnan(ev[:, l, p, 0]))
        if prev_out is not None:
            hop[l, p] = tin - prev_out
            first_in[l, p] = np.nanmin(ev[:, l, p, 0]) - prev_out
        work[l, p] = tout - tin
        spread_out[l, p] = tout - np.nanmin(ev[:, l, p, 1])
        prev_out = tout
sl = slice(1, L)      # layer 0's first hop has no predecessor
out.append("%-6s %5s %9s %9s %9s %9s" % ("phase", "CTAs", "hop", "first-in", "work", "out-spread"))
tot = 0.0
for p in range(NP):
    h, f, w, so = hop[sl, p].mean(), first_in[sl, p].mean(), work[sl, p].mean(), spread_out[sl, p].mean()
    tot += h + w
    out.append("%-6s %5d %9.0f %9.0f %9.0f %9.0f" % (names[p], n_part[p], h, f, w, so))
out.append("sum per layer %.0f ns  (x %d layers = %.3f ms; hops %.0f ns, work %.0f ns)" % (tot, L, tot * L * 1e-6, hop[sl].sum(1).mean(), work[sl].sum(1).mean()))
sub = t[:, 600:700].astype(np.float64)
labels = {}
for kd, nm in enumerate(["QKV", "O", "CQ", "XK", "XV", "CO", "FC1", "FC2", "LM"]):
    for k, st_ in enumerate(["B frags + ldmatrix + mma", "slice reductions", "epilogue + stores"]):
        labels[kd * 8 + k] = "%s: %s" % (nm, st_)
labels.update({80: "cross: scores", 81: "cross: max", 82: "cross: P.V", 83: "cross: fold + stores",
               90: "self: scores + group soft-max", 91: "self: merges + own key", 92: "self: barrier", 93: "self: final merge + store"})
out.append("sub-stages of the traced step, CYCLES per layer (thread 0; mean / max over the CTAs that ran the stage)")
for k in sorted(labels):
    col = sub[:, k]
    if (col > 0).any():
        out.append("  %-34s %7.0f %7.0f" % (labels[k], col[col > 0].mean() / L, col.max() / L))
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
