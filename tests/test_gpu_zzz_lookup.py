"""GPU tests of the entry points written after the round's last GPU call (run in the emulator only so far): `print`, mb_existence /
mb_position_lookup and the meryl-lookup tool.  The file sorts last on purpose: a surprise here must not hide the tests in front of it.
The same cases in the emulator: tests/test_cusim_union.py."""
import pytest

import util
import union_cases as uc

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not on this box")
@pytest.mark.parametrize("k", [21, 40])
def test_cli_print_vs_reference(tmp_path, k):
    uc.case_cli_print(str(tmp_path), k, n=30_000)


@pytest.mark.skipif(not util.have_ref_lookup(), reason="reference meryl-lookup not on this box")
@pytest.mark.parametrize("k", [21, 40, 9])
def test_existence_vs_reference_lookup(tmp_path, monkeypatch, k):
    uc.case_existence(str(tmp_path), k, n=20_000, monkeypatch=monkeypatch)


@pytest.mark.skipif(not util.have_ref_lookup(), reason="reference meryl-lookup not on this box")
@pytest.mark.parametrize("k", [21, 40])
def test_cli_lookup_vs_reference(tmp_path, k):
    uc.case_cli_lookup(str(tmp_path), k, n=20_000)
