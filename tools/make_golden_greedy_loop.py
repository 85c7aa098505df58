"""Whole-loop pin of the CPU oracle: a FREE-RUNNING greedy decode done entirely with independent code - HuggingFace
transformers' Whisper (fp32, the same synthetic tensors) for log-mel / encoder / decoder and HF's
WhisperTimeStampLogitsProcessor (OpenAI's timestamp grammar) for the logits filter - writes its token sequence to
tests/golden/greedy_loop.npz.  tests/test_oracle_golden.py requires oracle.full() to reproduce it token for token up to
the point where whisper.cpp's own completion rule stops the window.  The only glue that is ours: argmax, the list of
task / language tokens whisper.cpp suppresses, and suppress_blank at the first step.

Run:  python tools/make_golden_greedy_loop.py      (CPU, ~10 s)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden as mg  # noqa: E402
from speaksense_b200 import synth  # noqa: E402
from transformers import GenerationConfig  # noqa: E402
from transformers.generation.logits_process import WhisperTimeStampLogitsProcessor  # noqa: E402

CASES = [("tiny_en_peaked", "tiny.en", 0, None), ("micro_v3_peaked", "micro-v3", 1, "zh")]      # = the tests' fixtures
N_MAX = 160


def main():
    torch.set_grad_enabled(False)
    out = {}
    for name, shape, seed, lang in CASES:
        path = "/tmp/ss_golden_loop_%s.bin" % name
        synth.write_model(path, shape, "peaked", seed)
        model = synth.read_model(path)
        hp = model["hparams"]
        st = synth.special_tokens(hp.n_vocab)
        pcm = synth.synth_audio(seed=1234)
        hf = mg.hf_model(model)
        enc = hf.model.encoder(torch.from_numpy(mg.hf_log_mel(pcm, model["filters"])[None].astype(np.float32))).last_hidden_state
        prompt = mg.prompt_tokens(hp, lang)
        cfg = GenerationConfig(eos_token_id=st["eot"], no_timestamps_token_id=st["no_timestamps"], max_initial_timestamp_index=50)
        proc = WhisperTimeStampLogitsProcessor(cfg, begin_index=len(prompt))
        supp = [st["sot"] + i for i in range(101)] + [st["translate"], st["transcribe"], st["solm"], st["prev"], st["nosp"]]
        ids, toks = list(prompt), []
        for step in range(N_MAX):
            h = hf.model.decoder(input_ids=torch.tensor([ids]), encoder_hidden_states=enc).last_hidden_state[:, -1]
            logits = hf.proj_out(h).float()
            logits[:, supp] = -float("inf")
            if step == 0:
                logits[:, st["eot"]] = -float("inf")
            tok = int(proc(torch.tensor([ids]), logits).argmax(-1))
            toks.append(tok); ids.append(tok)
            if tok == st["eot"]:
                break
        out[name + "_tokens"] = np.array(toks, np.int32)
        out[name + "_lang"] = lang or ""
        print(name, len(toks), toks[:6])
        os.remove(path)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "greedy_loop.npz"), **out)


if __name__ == "__main__":
    main()
