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


def _lockstep_until_they_part(ref, orc, seed, steer, L=8192):
    """The reference's own dare_log.h (compiled unmodified) and the oracle in ENGINE rules get the same appends and the same
    pruning, over several laps of a small ring.  Returns (step, kind) of the first step at which ANY observable differs
    (return value, offsets, bytes) -- kind is "E1" or "E2" -- having asserted that everything was identical before it and
    that the difference is exactly one of the two documented divergences."""
    rng = np.random.default_rng(seed)
    orc.set_rules(O.RULES_ENGINE)
    a, b = O.Log(ref, L), O.Log(orc, L)
    try:
        for step in range(1, 4000):
            n = int(rng.integers(1, 200))
            o = b.offsets()
            left = L - o["end"] if o["end"] != L else 0
            if steer == "E1" and step > 40 and 65 <= left <= 264:
                n = left - 64                                      # header + data == what is left: the entry ends exactly at len
            data = O.cmd_image(bytes(rng.integers(0, 256, size=n, dtype=np.uint8)))
            ra, rb = a.append(1, step, 5, O.SEND, data), b.append(1, step, 5, O.SEND, data)
            oa, ob = a.offsets(), b.offsets()
            if ra == rb and oa == ob and np.array_equal(a.image(), b.image()):
                if step % 7 == 0 and not (steer == "E2" and step > 40):   # prune: all but the newest entry is applied
                    a.set_offsets(head=oa["tail"], apply=oa["tail"], commit=oa["tail"])
                    b.set_offsets(head=ob["tail"], apply=ob["tail"], commit=ob["tail"])
                continue
            # they part here, and only in one of the two documented ways
            if ra == rb and ra != 0 and oa["end"] == L and ob["end"] == 0:
                # E1: the entry ended exactly at len.  Same entry, same bytes; the reference stores end = len (its "log is
                # empty" sentinel, dare_log.h:158-162: the next append restarts at idx 1), the engine stores end = 0
                assert {k: v for k, v in oa.items() if k != "end"} == {k: v for k, v in ob.items() if k != "end"}
                assert np.array_equal(a.image(), b.image())
                return step, "E1"
            if rb == 0 and ra != 0:
                # E2: the ring is full.  The engine refuses and changes nothing; the reference appends anyway
                # (dare_log.h:487-505 compares against head without keeping a byte free) and ends up with end == head --
                # indistinguishable from an empty log -- or past it
                used = (ob["end"] - ob["head"]) % L if ob["end"] != L else 0
                assert used + 64 + n >= L - 64, (used, n)            # it really was (nearly) full
                return step, "E2"
            raise AssertionError(f"seed {seed} step {step}: reference {ra} {oa} vs engine rules {rb} {ob}")
        raise AssertionError("no divergence trigger reached")
    finally:
        a.close(); b.close()
        orc.set_rules(O.RULES_REFERENCE)


def test_reference_and_engine_rules_part_only_at_E1_or_E2(ref, orc):
    """say plainly where wrap-lap parity stops being parity with the reference.  From an empty log,
    through wraps and pruning, the reference and the engine rules are the same function up to the first exact-fit entry
    (E1) or the first append into a full ring (E2) -- and they part nowhere else."""
    kinds = {}
    for seed in range(40):
        # after 40 steps (a few laps with pruning) the stream is steered: even seeds craft an entry that ends exactly at
        # len as soon as the random traffic leaves room for one, odd seeds stop pruning so that the ring fills up
        step, kind = _lockstep_until_they_part(ref, orc, seed, "E1" if seed % 2 == 0 else "E2")
        kinds.setdefault(kind, []).append(step)
    assert set(kinds) == {"E1", "E2"}, kinds
    assert min(min(v) for v in kinds.values()) > 40          # laps of byte-identical behaviour first (8 KiB ring)
