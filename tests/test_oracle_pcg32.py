"""The oracle's pcg32 restatement against the reference's own known-answer files
(pcg-cpp/test-high/expected/check-pcg32.out, check-pcg32_oneseq.out) -- CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

KAT = json.load(open(os.path.join(GOLDEN, "pcg32_kat.json")))


def test_full_known_answer_text_two_arg(O):
    # pcg32 rng(42u, 54u): outputs, backstep(6) ("Again"), bounded coins/dice, distance, shuffle
    assert O.pcg32_kat_text(True, 5) == KAT["check-pcg32"]


def test_full_known_answer_text_one_arg(O):
    # pcg32{42}: the seeding form df.cpp:334 uses; identical to pcg32_oneseq(42)
    assert O.pcg32_kat_text(False, 5) == KAT["check-pcg32_oneseq"]


def test_headline_vectors(O):
    assert [hex(x) for x in O.pcg32_draw(42, 54, 0, 6)] == ["0xa15c02b7", "0x7b47f409", "0xba1d3330", "0x83d2f293", "0xbfa4784b", "0xcbed606e"]
    assert [hex(x) for x in O.pcg32_draw(42, 0, 0, 6, False)] == ["0xc2f57bd6", "0x6b07c4a9", "0x72b7b29b", "0x44215383", "0xf5af5ead", "0x68beb632"]
    assert O.pcg32_state(42, 54) == (1753877967969059832, 109)
    assert O.pcg32_state(42, 0, False) == (10915315373440060052, 1442695040888963407)


@pytest.mark.parametrize("rec", KAT["jump_ahead"], ids=lambda r: "delta=" + r["delta"])
def test_jump_ahead_matches_vendored_header(O, rec):
    d = int(rec["delta"])
    assert [int(x) for x in O.pcg32_draw(42, 54, d, 4)] == rec["two_arg"]
    assert [int(x) for x in O.pcg32_draw(42, 0, d, 4, False)] == rec["one_arg"]


def test_jump_ahead_equals_stepping(O):
    seq = O.pcg32_draw(9, 3, 0, 5000)
    for d in (1, 2, 63, 64, 1000, 4999):
        assert O.pcg32_draw(9, 3, d, 1)[0] == seq[d]


def test_live_reference_header_if_built(O):
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(0)
    for _ in range(20):
        seed, stream, delta = (int(x) for x in rng.integers(0, 2 ** 62, 3))
        assert np.array_equal(O.pcg32_draw(seed, stream, delta, 8), O.ref_pcg32_draw(seed, stream, delta, 8))
