"""The two documented divergences of the CUDA engine from the reference,
as implemented by the oracle's ORC_RULES_ENGINE mode (the GPU parity tests compare the
engine against the oracle in this mode)."""
import numpy as np

import orc as O


def test_E1_exact_fit_wraps_to_zero_and_keeps_counting(orc, ref):
    L = 8192
    # reference: the entry that ends exactly at len makes the log look empty, idx restarts at 1
    log = O.Log(ref, L)
    for i in range(64):
        assert log.append(1, i + 1, 7, O.SEND, O.cmd_image(b"x" * 64)) == i + 1
        if i == 40:
            o = log.offsets(); log.set_offsets(head=o["tail"], apply=o["tail"], commit=o["tail"])
    assert log.offsets()["end"] == L                      # == "empty" sentinel (dare_log.h:158-162)
    assert log.append(1, 65, 7, O.SEND, O.cmd_image(b"y" * 64)) == 1      # the reference's bug (H11 iv)
    log.close()
    # engine rules: end = 0, the next entry is idx 65 at offset 0
    orc.set_rules(O.RULES_ENGINE)
    log = O.Log(orc, L)
    for i in range(64):
        assert log.append(1, i + 1, 7, O.SEND, O.cmd_image(b"x" * 64)) == i + 1
        if i == 40:
            o = log.offsets(); log.set_offsets(head=o["tail"], apply=o["tail"], commit=o["tail"])
    assert log.offsets()["end"] == 0
    assert log.append(1, 65, 7, O.SEND, O.cmd_image(b"y" * 64)) == 65
    o = log.offsets()
    assert (o["tail"], o["end"]) == (0, 128)
    log.close()
    orc.set_rules(O.RULES_REFERENCE)


def test_E2_full_ring_refuses_without_touching_state(orc):
    orc.set_rules(O.RULES_ENGINE)
    L = 4096
    log = O.Log(orc, L)
    n = 0
    while log.append(1, n + 1, 3, O.SEND, O.cmd_image(b"z" * 70)):
        n += 1
        assert n < 100
    before, img = log.offsets(), log.image()
    assert before["end"] != before["head"]                     # strictly before head
    for _ in range(3):
        assert log.append(1, 999, 3, O.SEND, O.cmd_image(b"z" * 70)) == 0
    assert log.offsets() == before and np.array_equal(log.image(), img)
    # space appears once the head moves (pruning): the append goes through and wraps with a ghost header
    log.set_offsets(head=before["tail"], apply=before["tail"], commit=before["tail"])
    assert log.append(1, 1000, 3, O.SEND, O.cmd_image(b"w" * 70)) == n + 1
    assert log.offsets()["tail"] == 0
    log.close()
    orc.set_rules(O.RULES_REFERENCE)


def test_engine_rules_equal_reference_rules_without_wrap(orc):
    """Away from the two bug cases the modes are the same function."""
    import streams as S
    stream = S.ragged_stream(800, 400, seed=9, close_every=90)
    imgs = []
    for rules in (O.RULES_REFERENCE, O.RULES_ENGINE):
        orc.set_rules(rules)
        log = O.Log(orc, 1 << 20)
        rets = [log.append(1, rid, clt, typ, O.cmd_image(p)) for typ, clt, rid, p in stream]
        imgs.append((rets, log.offsets(), log.image()))
        log.close()
    orc.set_rules(O.RULES_REFERENCE)
    assert imgs[0][0] == imgs[1][0] and imgs[0][1] == imgs[1][1] and np.array_equal(imgs[0][2], imgs[1][2])
