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


# ---- golden vectors produced by the RUNNING reference (tests/golden/gen_refstack_golden.py) -----------------------
import hashlib

import refstack as R

REFGOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refstack_golden.json")


def _refgold():
    with open(REFGOLD) as f:
        return json.load(f)["scenarios"]


@pytest.mark.parametrize("gold", _refgold(), ids=lambda g: g["name"])
def test_oracle_reproduces_reference_run(orc, gold):
    """The oracle's cluster restatement, given the leader index and term the reference's election produced,
    reproduces the logs the reference's own replicas held (SHA-256; followers under the H5 reply mask)."""
    orc.set_rules(O.RULES_REFERENCE)
    n = gold["n"]
    c = O.Cluster(orc, n, leader=gold["leader"], term=gold["term"], length=O.LOG_SIZE)
    c.prologue()
    for typ, clt, rid, payload in R.expected_stream(gold["leader"], gold["nconn"], gold["nreq"], gold["plen"]):
        assert c.submit(typ, clt, rid, O.cmd_image(payload))
    c.round(); c.round()
    end = gold["end"]
    ents = O.walk_entries(c.image(gold["leader"], 0, end), 0, end, O.LOG_SIZE)
    assert len(ents) == gold["entries"]
    for i in range(n):
        o = c.offsets(i)
        assert {k: o[k] for k in ("head", "apply", "commit", "end")} == gold["offsets"][i]
        img = c.image(i, 0, end)
        img = img if i == gold["leader"] else O.mask_replies(img, ents)
        assert hashlib.sha256(img.tobytes()).hexdigest() == gold["sha256"][i], f"replica {i}"
    c.close()
