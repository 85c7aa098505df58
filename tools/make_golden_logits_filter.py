"""Golden masks for the decoder's logits filter from an INDEPENDENT implementation: HuggingFace transformers'
WhisperTimeStampLogitsProcessor (OpenAI's timestamp rules).  Run here (transformers is in the image); writes
tests/golden/logits_filter.npz, which tests/test_oracle_golden.py holds oracle/whisper_oracle.c's process_logits to.

Only states where whisper.cpp's rules and OpenAI's coincide are compared (see the test for the documented differences):
the sampled history is non-empty, and the "timestamps do not decrease" rule is exercised only in the
(last = timestamp, penultimate = text) state, where both implementations mask [timestamp_begin, last timestamp)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speaksense_b200 import synth  # noqa: E402
from transformers import GenerationConfig  # noqa: E402
from transformers.generation.logits_process import WhisperTimeStampLogitsProcessor  # noqa: E402

N_VOCAB = 51866
st = synth.special_tokens(N_VOCAB)
BEG, EOT, NOT = st["beg"], st["eot"], st["no_timestamps"]

# (name, history of sampled ids, logits seed, boost) ; boost: added to every timestamp logit: -8 keeps the total timestamp
# probability below the best text token (probability-sum rule off), +6 switches it on
T = lambda k: BEG + k      # noqa: E731
CASES = [
    ("text_text", [1000, 2000], 1, -8.0),
    ("ts_then_text", [T(0), 1000], 2, -8.0),
    ("text_then_ts", [T(0), 1000, 2000, T(50)], 3, -16.0),          # closing timestamp: next must be a timestamp >= T(50) or EOT
    ("ts_ts", [T(0), 1000, T(50), T(50)], 4, -8.0),                # pair complete: next cannot be a timestamp
    ("single_ts", [T(10)], 5, -8.0),                               # one sampled token, a timestamp: penultimate counts as timestamp
    ("sum_rule_text", [1000, 2000, 3000], 6, 6.0),                # timestamp mass beats every text token: text masked
    ("sum_rule_off", [1000, 2000, 3000], 7, -6.0),
]


def main():
    cfg = GenerationConfig(eos_token_id=EOT, no_timestamps_token_id=NOT, max_initial_timestamp_index=50)
    out = {}
    for name, ids, seed, boost in CASES:
        rng = np.random.default_rng(seed)
        raw = rng.standard_normal(N_VOCAB).astype(np.float32) * 2.0
        raw[BEG:] += np.float32(boost)
        proc = WhisperTimeStampLogitsProcessor(cfg, begin_index=0)
        scores = proc(torch.tensor([ids]), torch.from_numpy(raw)[None].clone())[0].numpy()
        out[name + "_ids"] = np.array(ids, np.int32)
        out[name + "_seed"] = np.int64(seed)
        out[name + "_boost"] = np.float32(boost)
        out[name + "_masked"] = np.packbits(np.isneginf(scores))
        print(name, "masked", int(np.isneginf(scores).sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "logits_filter.npz"), names=np.array([c[0] for c in CASES]),
                        n_vocab=np.int64(N_VOCAB), beg=np.int64(BEG), eot=np.int64(EOT), **out)


if __name__ == "__main__":
    main()
