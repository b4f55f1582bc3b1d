"""CPU functional cases: golden-value comparison and checkpoint-resume equivalence through the real ``pretrain_gpt.py`` entry point."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "functional_tests", "python_test_utils"))
CASES = os.path.join(HERE, "functional_tests", "test_cases")


@pytest.mark.parametrize("case", ["gpt/gpt_tiny_tp1_cpu", "moe/mixtral_tiny_ep1_cpu", "gpt/gpt_tiny_tp2_reshard_pp2_cpu"])
def test_functional_case(case):
    from run_case import run_case

    run_case(os.path.join(CASES, case))
