"""`meryl union-sum` through the product's CUDA sources run in the cusim host emulator (tests/cusim): block decode,
merge-path union with summed counts, stream-mode encoder -- byte-compared with the reference's golden union dbs and
with the oracle.  (The GPU versions of the same cases: tests/test_gpu_union.py.)"""
import os
import subprocess

import pytest

import util
import meryl_b200 as mb
import union_cases as uc

SIM_DIR = os.path.join(util.ROOT, "tests", "cusim")
SIM_LIB = os.path.join(SIM_DIR, "_build", "libmeryl_b200_sim.so")


@pytest.fixture(scope="module", autouse=True)
def sim_library():
    subprocess.run(["make", "-s", "-j8", "-C", SIM_DIR], check=True)
    saved = (mb.LIB_PATH, mb._lib)
    mb.LIB_PATH, mb._lib = SIM_LIB, None
    yield
    mb.LIB_PATH, mb._lib = saved


@pytest.mark.parametrize("name", util.union_golden_cases())
def test_sim_union_golden(tmp_path, name):
    uc.case_golden(str(tmp_path), name)


@pytest.mark.parametrize("k", [7, 8, 21, 31, 32, 33, 51, 64])
def test_sim_union_k_sweep(tmp_path, k):
    uc.case_k_sweep(str(tmp_path), k, n=8_000)


def test_sim_union_of_unions_empty_single(tmp_path):
    uc.case_union_of_unions(str(tmp_path), n=6_000)


def test_sim_union_low_complexity(tmp_path):
    uc.case_low_complexity(str(tmp_path))


def test_sim_union_groups(tmp_path, monkeypatch):
    uc.case_groups(str(tmp_path), monkeypatch)


@pytest.mark.parametrize("op", mb.SET_OPERATIONS)
def test_sim_set_operations(tmp_path, op):
    uc.case_set_operations(str(tmp_path), op, n=6_000)


def test_sim_value_operations(tmp_path):
    uc.case_value_operations(str(tmp_path), n=5_000)


def test_sim_union_small_ring(tmp_path, monkeypatch):
    uc.case_small_ring(str(tmp_path), monkeypatch)


@pytest.mark.parametrize("k,decoder", [(21, "warp"), (21, "thread"), (51, "warp"), (64, "thread"), (32, "warp"), (31, "auto")])
def test_sim_union_long_blocks(tmp_path, monkeypatch, k, decoder):
    uc.case_long_blocks(str(tmp_path), monkeypatch, k, decoder)


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not built")
def test_sim_threshold_words_vs_reference(tmp_path):
    uc.case_threshold_words(str(tmp_path))


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not built")
def test_sim_count_suffix_vs_reference(tmp_path):
    uc.case_count_suffix(str(tmp_path))


@pytest.mark.skipif(not util.have_ref(), reason="reference binary not built")
@pytest.mark.parametrize("k", [21, 40, 9])
def test_sim_lookup_vs_reference_print(tmp_path, monkeypatch, k):
    uc.case_lookup(str(tmp_path), k, monkeypatch=monkeypatch)
