"""The C oracle must reproduce, byte for byte, the dbs that the unmodified
reference binary wrote for the committed golden inputs (tests/golden/, made by
tests/golden/make_golden.py).  Runs without the reference and without a GPU."""
import os

import pytest

import util


@pytest.mark.parametrize("name", util.golden_cases())
def test_oracle_reproduces_golden(tmp_path, name):
    inp, gold, words, _ = util.load_golden(name, str(tmp_path))
    out = str(tmp_path / "ora.meryl")
    util.run_oracle(words + [inp, "output", out])
    util.assert_same_db(gold, out)
