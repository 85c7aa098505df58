"""Hand-derived cases for the parts of whisper_full that the HF-built goldens do not reach: the per-token decoder bookkeeping
(seek_delta / result_len update, completion rule, the two failure rules) and the statistics of the fallback gates (average
log-probability over result_len tokens, entropy of the last 32 tokens; thresholds whisper.rs:159-162).  Expected values are
worked out from the rules as SURVEY.md Appendix A.5 states them, not produced by the oracle: the oracle's own functions
(`token_bookkeeping`, `sequence_score` - the ones `wo_full` calls) are driven through test probes."""
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def st(oracle_mod, tiny_en_peaked):
    m = oracle_mod.OracleModel(tiny_en_peaked)
    s = m.new_state()
    yield s, m
    s.close(); m.close()


def _tok(m):
    from speaksense_b200 import synth
    t = synth.special_tokens(m.hparams["n_vocab"])
    return t["beg"], t["eot"]


def test_timestamp_pair_then_eot_completes(st):
    s, m = st
    beg, eot = _tok(m)
    ids = [beg, 1000, 1001, 1002, beg + 50, beg + 50, 1003, eot]
    r = s.bookkeeping(ids)
    # <|1.00|> at i = 4 sets seek_delta = 2 * 50 = 100 (10 ms ticks) and result_len = 5; its repeat at i = 5 moves result_len to 6;
    # EOT at i = 7 completes with what was recorded
    assert r == dict(failed=False, completed=True, result_len=6, seek_delta=100, has_ts=True, steps=8)


def test_timestamps_going_backwards_fail(st):
    s, m = st
    beg, _ = _tok(m)
    ids = [beg, 1000, beg + 100, beg + 100, 1001, beg + 50]
    r = s.bookkeeping(ids)
    # at i = 5: seek_delta (200) > 2 * 50 and result_len (4) < 5 -> failed; the state keeps the last good values
    assert r["failed"] and not r["completed"] and r["result_len"] == 4 and r["seek_delta"] == 200 and r["steps"] == 6


def test_repeating_the_open_timestamp_is_not_a_failure(st):
    s, m = st
    beg, _ = _tok(m)
    # the same timestamp twice in a row: sd_new == seek_delta, so the "backwards" rule (strictly greater) does not fire
    r = s.bookkeeping([beg, 1000, beg + 100, beg + 100, 1001, 1002])
    assert not r["failed"] and not r["completed"] and r["result_len"] == 4 and r["seek_delta"] == 200


def test_eot_without_any_timestamp(st):
    s, m = st
    beg, eot = _tok(m)
    # no timestamp > beg was seen: seek_delta is still the whole window (3000 ticks), so seek + seek_delta + 100 >= seek_end holds and
    # the window is accepted as a whole: result_len = i + 1
    r = s.bookkeeping([beg, 1000, 1001, eot])
    assert r == dict(failed=False, completed=True, result_len=4, seek_delta=3000, has_ts=False, steps=4)


def test_completion_rule_last_second_of_the_audio(st):
    s, m = st
    beg, _ = _tok(m)
    # 30 s of audio (seek_end = 3000): a timestamp at 29.00 s (tick 2900) is within 1 s of the end -> completes at that token
    r = s.bookkeeping([beg, 1000, 1001, beg + 1450, 1002])
    assert r == dict(failed=False, completed=True, result_len=4, seek_delta=2900, has_ts=True, steps=4)
    # ... but at 28.98 s it does not
    r = s.bookkeeping([beg, 1000, 1001, beg + 1449, 1002])
    assert not r["completed"] and not r["failed"] and r["seek_delta"] == 2898
    # a 12 s clip (seek_end = 1200): 11.00 s is enough
    r = s.bookkeeping([beg, 1000, beg + 550], seek_end=1200)
    assert r["completed"] and r["seek_delta"] == 1100 and r["result_len"] == 3
    # second window of a longer clip: seek = 2900, seek_end = 4500 -> needs seek_delta >= 1500
    r = s.bookkeeping([beg, 1000, beg + 749], seek=2900, seek_end=4500)
    assert not r["completed"]
    r = s.bookkeeping([beg, 1000, beg + 750], seek=2900, seek_end=4500)
    assert r["completed"] and r["seek_delta"] == 1500


def test_token_cap(st):
    s, m = st
    beg, _ = _tok(m)
    n_max = 220
    text = list(range(1000, 1000 + n_max))
    # the cap (i == n_max - 1) with a last timestamp in the first half of the window (seek_delta < 1500): failed
    ids = [beg, 1000, beg + 700] + text
    r = s.bookkeeping(ids[:n_max], n_max=n_max)
    assert r["failed"] and r["steps"] == n_max and r["seek_delta"] == 1400
    # ... in the second half: neither failed nor completed - the window ends by the cap with result_len at that timestamp
    ids = [beg, 1000, beg + 750] + text
    r = s.bookkeeping(ids[:n_max], n_max=n_max)
    assert not r["failed"] and not r["completed"] and r["result_len"] == 3 and r["seek_delta"] == 1500 and r["steps"] == n_max
    # ... with no timestamp at all: failed (result_len == 0)
    r = s.bookkeeping(([beg] + text)[:n_max], n_max=n_max)
    assert r["failed"] and r["result_len"] == 0


def test_entropy_and_logprob_statistics(st):
    s, _ = st
    # 32 times the same token: entropy 0 (fails entropy_thold = 2.4: the repetition detector)
    r = s.score([7] * 40, [-0.5] * 40, 40)
    assert r["entropy"] == pytest.approx(0.0, abs=1e-12)
    assert r["avg_logprobs"] == pytest.approx(-0.5, rel=1e-6) and r["sum_logprobs"] == pytest.approx(-20.0, rel=1e-6)
    # 32 distinct tokens: ln 32
    r = s.score(list(range(100, 132)), [-1.0] * 32, 32)
    assert r["entropy"] == pytest.approx(math.log(32.0), rel=1e-12)
    # only the LAST 32 tokens of the first result_len count: 8 early repeats are out of the window
    ids = [5] * 8 + list(range(100, 132))
    r = s.score(ids, [-0.25] * len(ids), len(ids))
    assert r["entropy"] == pytest.approx(math.log(32.0), rel=1e-12)
    # half of the window is one token: -(1/2) ln(1/2) - 16 (1/32) ln(1/32) = (1/2) ln 2 + (1/2) ln 32 = 3 ln 2 = 2.0794 < 2.4
    ids = [9] * 16 + list(range(200, 216))
    r = s.score(ids, [-0.1] * 32, 32)
    assert r["entropy"] == pytest.approx(3.0 * math.log(2.0), rel=1e-12) and r["entropy"] < 2.4
    # result_len shorter than the sequence: statistics over the first result_len tokens only; length_penalty <= 0 -> score = sum / len
    plogs = np.array([-0.1, -0.2, -0.3, -5.0, -5.0], np.float32)
    r = s.score([1, 2, 3, 4, 5], plogs, 3)
    assert r["sum_logprobs"] == pytest.approx(-0.6, rel=1e-6) and r["avg_logprobs"] == pytest.approx(-0.2, rel=1e-6)
    assert r["score"] == pytest.approx(-0.2, rel=1e-6)
    assert r["entropy"] == pytest.approx(math.log(3.0), rel=1e-12)


# ---- beam search: candidate assignment: the hand-derived cases of tests/beam_cases.py against the oracle's beam_assign
def test_beam_candidate_assignment_hand_derived_cases():
    from oracle import oracle
    from tests.beam_cases import CASES
    for name, cands, live, i, want in CASES:
        assert oracle.beam_assign(cands, live, i) == want, name
