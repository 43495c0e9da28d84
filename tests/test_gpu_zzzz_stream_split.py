"""The stream writer's block limit at its real size, against the reference binary (tests/union_cases.py::
case_stream_block_real_limit; ::case_stream_block_split covers the logic at lowered limits in tests/test_gpu_union.py).
This file sorts last because it is the heaviest."""
import pytest

import util
import union_cases as uc

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not on this box")
def test_prefix_of_more_than_16_mi_kmers_vs_reference(tmp_path):
    uc.case_stream_block_real_limit(tmp_path)
