"""The restated oracle against golden vectors generated from the compiled
reference header (tests/golden/gen_golden.py).  Runs anywhere (no reference tree)."""
import json
import os

import pytest

import orc as O
import scenarios

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.json")


@pytest.fixture(scope="module")
def results(orc):
    orc.set_rules(O.RULES_REFERENCE)
    return {r["name"]: r for r in scenarios.all_scenarios(orc)}


def _gold():
    with open(GOLD) as f:
        return json.load(f)["scenarios"]


@pytest.mark.parametrize("gold", _gold(), ids=lambda g: g["name"])
def test_golden(results, gold):
    assert results[gold["name"]] == gold
