"""Hand-derived cases for the candidate assignment of one beam-search step (whisper.cpp whisper_full_with_state, BEAM_SEARCH branch
behind "update each decoder"): the candidates of all live decoders are sorted by sum_logprobs_all, descending, and handed to the live
decoders in order; from the second sampled token on (i > 0) a decoder skips the candidates that follow its own and carry the same
token sequence; the candidate cursor wraps to 0 when it runs past the end.  Derived by hand from that rule - not from the oracle and
not from the engine - and applied to both: tests/test_oracle_rules.py (oracle.beam_assign) and tests/test_abi.py
(ss_debug_beam_assign, the function the engine's beam decoders run).

A case: (name, candidates [(token ids, sum_logprobs_all, decoder index)] in production order, live flags, i, expected candidate index
per decoder (-1: not live))."""

CASES = [
    # two decoders, two candidates each (history + new token): sums -1.0, -3.0 (decoder 0), -2.0, -2.5 (decoder 1)
    ("best first, two decoders", [([10, 20], -1.0, 0), ([10, 21], -3.0, 0), ([11, 30], -2.0, 1), ([11, 31], -2.5, 1)], [True, True], 1, [0, 2]),
    ("best first, four decoders", [([10, 20], -1.0, 0), ([10, 21], -3.0, 0), ([11, 30], -2.0, 1), ([11, 31], -2.5, 1)], [True] * 4, 1, [0, 2, 3, 1]),
    # both decoders carry the same history and propose the same best token: the second decoder must not continue with the copy
    ("duplicate skipped", [([10, 20], -1.0, 0), ([10, 21], -2.0, 0), ([10, 20], -1.0, 1), ([10, 22], -3.0, 1)], [True, True], 1, [0, 1]),
    # ... but at the first sampled token (i == 0) whisper.cpp does not look for duplicates: every decoder starts from the same logits,
    # the sorted list begins with one copy of the best token per decoder, and all beams take it
    ("no duplicate check at the first token", [([20], -1.0, 0), ([21], -2.0, 0), ([20], -1.0, 1), ([21], -2.0, 1)], [True, True], 0, [0, 2]),
    # a run of three copies is skipped as a whole
    ("run of three copies", [([10, 20], -1.0, 0), ([10, 20], -1.0, 1), ([10, 20], -1.0, 2), ([10, 23], -4.0, 2), ([10, 21], -2.0, 0)], [True] * 3, 2, [0, 4, 3]),
    # equal score, same length, different tokens: not duplicates (stable order)
    ("same score, different tokens", [([10, 20], -1.0, 0), ([10, 25], -1.0, 1), ([10, 21], -2.0, 0)], [True, True], 3, [0, 1]),
    # decoder 1 has completed: it gets nothing and does not consume a candidate; decoder 2 takes the second best
    ("finished decoder keeps out", [([10, 20], -1.0, 0), ([10, 25], -1.0, 1), ([10, 21], -2.0, 0)], [True, False, True], 3, [0, -1, 1]),
    # three live decoders, but after duplicate skipping only two distinct candidates: the cursor wraps to the best one
    ("cursor wraps", [([10, 20], -1.0, 0), ([10, 20], -1.0, 1), ([10, 21], -2.0, 2)], [True] * 3, 1, [0, 2, 0]),
    # a shorter sequence with the same prefix is not a duplicate
    ("prefix is not a duplicate", [([10, 20], -1.0, 0), ([10], -1.5, 1), ([10, 20], -2.0, 1)], [True, True, True], 4, [0, 1, 2]),
]
