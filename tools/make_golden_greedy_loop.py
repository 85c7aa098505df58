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


def greedy(hf, enc, prompt, st, n_steps, stop=None):
    """free-running greedy loop: HF decoder + HF's timestamp processor; `stop(tokens)` ends it early"""
    cfg = GenerationConfig(eos_token_id=st["eot"], no_timestamps_token_id=st["no_timestamps"], max_initial_timestamp_index=50)
    proc = WhisperTimeStampLogitsProcessor(cfg, begin_index=len(prompt))
    supp = [st["sot"] + i for i in range(101)] + [st["translate"], st["transcribe"], st["solm"], st["prev"], st["nosp"]]
    supp = [t for t in supp if t < st["beg"] and t != st["no_timestamps"]]
    ids, toks = list(prompt), []
    for step in range(n_steps):
        h = hf.model.decoder(input_ids=torch.tensor([ids]), encoder_hidden_states=enc).last_hidden_state[:, -1]
        logits = hf.proj_out(h).float()
        logits[:, supp] = -float("inf")
        if step == 0:
            logits[:, st["eot"]] = -float("inf")
        tok = int(proc(torch.tensor([ids]), logits).argmax(-1))
        toks.append(tok); ids.append(tok)
        if tok == st["eot"] or (stop is not None and stop(toks)):
            break
    return toks


def two_window_case():
    """A 45 s clip, not stream mode (context on), on a model scripted with 3.5 s per 24-token segment (seg_ticks = 175): window 1 runs
    into the token cap (n_text_ctx / 2 - 4 = 220 sampled tokens: 8 segments = 28 s, never the last second of the 45 s and never the
    <|30.00|> clamp, after which OpenAI's grammar and whisper.cpp's differ), window 2 starts at the last timestamp of window 1 and is
    prompted with [prev] + the tokens window 1 kept.  The model arithmetic is HF's; the window bookkeeping restated here is the documented
    whisper.cpp rule set: result_len / seek_delta follow the last timestamp token, the context is the last
    n_text_ctx / 2 kept tokens, a window completes on a timestamp within 1 s of the end of the audio."""
    name, shape, seed = "tiny_en_peaked", "tiny.en", 0
    path = "/tmp/ss_golden_loop2_%s.bin" % name
    synth.write_model(path, shape, "peaked", seed, seg_ticks=175)
    model = synth.read_model(path)
    hp = model["hparams"]
    st = synth.special_tokens(hp.n_vocab)
    pcm = synth.synth_audio(45 * 16000, seed=1234)
    hf = mg.hf_model(model)
    n_max = hp.n_text_ctx // 2 - 4
    beg = st["beg"]
    # window 1: frames [0, 3000)
    enc1 = hf.model.encoder(torch.from_numpy(mg.hf_log_mel(pcm[:480000], model["filters"])[None].astype(np.float32))).last_hidden_state
    w1 = greedy(hf, enc1, [st["sot"]], st, n_max)
    assert len(w1) == n_max and st["eot"] not in w1, "window 1 is meant to end on the token cap"
    ts_pos = [i for i, t in enumerate(w1) if t > beg]
    result_len = ts_pos[-1] + 1
    seek_delta = 2 * (w1[ts_pos[-1]] - beg)
    assert seek_delta >= 1500          # the cap rule does not fail the window (its last timestamp lies in the second half)
    kept = w1[:result_len]
    # window 2: frames [seek, seek + 3000), audio up to frame 4500
    seek = seek_delta
    pcm2 = pcm[seek * 160:]
    enc2 = hf.model.encoder(torch.from_numpy(mg.hf_log_mel(pcm2, model["filters"])[None].astype(np.float32))).last_hidden_state
    ctx = kept[-(hp.n_text_ctx // 2):]
    prompt2 = [st["prev"]] + ctx + [st["sot"]]
    seek_end = 4500 - 1      # n_len_org of a 45 s clip is 4499 frames

    def done(toks):
        ts = [t for t in toks if t > beg]
        return bool(ts) and toks[-1] > beg and seek + 2 * (toks[-1] - beg) + 100 >= seek_end
    w2 = greedy(hf, enc2, prompt2, st, n_max, stop=done)
    os.remove(path)
    print("two windows: w1", len(w1), "result_len", result_len, "seek_delta", seek_delta, "prompt2", len(prompt2), "w2", len(w2), w2[:6])
    return {name + "_ctx_w1_tokens": np.array(w1, np.int32), name + "_ctx_w1_result_len": np.int32(result_len),
            name + "_ctx_w1_seek_delta": np.int32(seek_delta), name + "_ctx_w2_prompt": np.array(prompt2, np.int32),
            name + "_ctx_w2_tokens": np.array(w2, np.int32)}


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
    out.update(two_window_case())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "greedy_loop.npz"), **out)


if __name__ == "__main__":
    main()
