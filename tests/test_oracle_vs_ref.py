"""The restated oracle (oracle/orc_log.h, cluster_sim.inc) against the compiled
reference header (oracle/_ref, built from /root/reference/src/include/dare/dare_log.h).

These tests only run where the reference tree exists (this container); on the GPU
box the same restatement is pinned by tests/golden/ instead (test_golden.py).
"""
import numpy as np
import pytest

import orc as O
import streams as S


def play(oracle, stream, length, term=1, pre=None):
    log = O.Log(oracle, length)
    if pre:
        pre(log)
    rets = []
    for typ, clt, rid, payload in stream:
        rets.append(log.append(term, rid, clt, typ, O.cmd_image(payload)))
    return log, rets


def same_log(a: O.Log, b: O.Log):
    oa, ob = a.offsets(), b.offsets()
    assert oa == ob
    ia, ib = a.image(), b.image()
    assert np.array_equal(ia, ib), f"first diff at {int(np.argmax(ia != ib))}"


def test_layout_constants(ref):
    assert ref.sizeof_entry() == 64
    assert ref.lib.ref_sizeof_log() == 319656
    assert ref.lib.ref_log_size() == O.LOG_SIZE


def test_kat_from_survey(ref, orc):
    """SURVEY.md s8c known-answer vector, re-generated from the compiled reference."""
    orc.set_rules(O.RULES_REFERENCE)
    lens = [64, 64, 64, 100, 4096]
    expect = [(1, 0, 128), (2, 128, 256), (3, 256, 384), (4, 384, 548), (5, 548, 4708),
              (6, 4708, 4772), (7, 4772, 4836)]
    for oracle in (ref, orc):
        log = O.Log(oracle)
        off = log.offsets()
        assert (off["head"], off["apply"], off["commit"]) == (0, 0, 0)
        assert off["end"] == off["tail"] == off["len"] == 67108864
        got = []
        for i, ln in enumerate(lens):
            idx = log.append(1, i + 1, 0x0100, O.SEND, O.cmd_image(S.payload_kat(i, ln)))
            o = log.offsets()
            got.append((idx, o["tail"], o["end"]))
        idx = log.append(1, 6, 0x0100, O.CONNECT, O.cmd_image(b""))
        o = log.offsets(); got.append((idx, o["tail"], o["end"]))
        idx = log.append(1, 7, 0x0100, O.NOOP, b"")
        o = log.offsets(); got.append((idx, o["tail"], o["end"]))
        assert got == expect
        # (idx, tail, end) above are the triples quoted in SURVEY.md s8c.  The FNV-1a-64
        # quoted there came from a probe harness that was never committed and cannot be
        # reproduced byte for byte; this checksum is regenerated from the compiled
        # reference header (oracle/_ref) and pinned in tests/golden/ as well.
        assert O.fnv1a(log.image(0, 4836)) == 0x6EF37439CDE8F856
        log.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("length", [4096, 1 << 16])
def test_random_streams_with_wraps(ref, orc, seed, length):
    """Small rings force every wrap rule of log_append_entry (ghost header, header
    does not fit, exact fit, full) -- reference rules must match bit for bit."""
    orc.set_rules(O.RULES_REFERENCE)
    stream = S.ragged_stream(600, 300, conns=3, seed=seed, close_every=50)
    rng = np.random.default_rng(seed)

    def run(oracle):
        log = O.Log(oracle, length)
        rets = []
        r = np.random.default_rng(seed + 99)
        for k, (typ, clt, rid, payload) in enumerate(stream):
            rets.append(log.append(1 + k // 200, rid, clt, typ, O.cmd_image(payload)))
            # advance head now and then so the ring keeps accepting entries
            if k % 7 == 0:
                o = log.offsets()
                if o["end"] != o["len"]:
                    log.set_offsets(head=o["tail"], apply=o["tail"], commit=o["tail"])
            if r.random() < 0.05:
                rets.append(log.append(1, 0, 0, O.NOOP, b""))
            if r.random() < 0.03:
                rets.append(log.append(1, 0, 0, O.CONFIG, O.cid_image(3)))
            if r.random() < 0.03:
                rets.append(log.append(1, 0, 0, O.HEAD, (12345).to_bytes(8, "little")))
        return log, rets

    la, ra = run(ref)
    lb, rb = run(orc)
    assert ra == rb
    same_log(la, lb)
    la.close(); lb.close()


@pytest.mark.parametrize("left", [0, 1, 40, 63, 64, 65, 100, 127, 128, 129, 200])
@pytest.mark.parametrize("typ", [O.SEND, O.NOOP, O.HEAD, O.CONFIG])
def test_wrap_edges(ref, orc, left, typ):
    """Place `end` exactly `left` bytes before len and append one entry (64 B payload)."""
    orc.set_rules(O.RULES_REFERENCE)
    length = 8192
    res = []
    for oracle in (ref, orc):
        log = O.Log(oracle, length)
        # fill with 64 B NOOP-size strides up to the desired position, head moved away
        log.append(1, 1, 7, O.SEND, O.cmd_image(b"x" * 64))
        pos = length - left
        log.poke(0, bytes(range(256)) * (length // 256))      # stale bytes everywhere
        # fabricate a log whose last entry ends at pos: tail entry must parse
        tail = pos - 64
        hdr = (41).to_bytes(8, "little") + (1).to_bytes(8, "little") + bytes(8) + bytes([7, 0, O.NOOP, 0]) + bytes(36)
        log.poke(tail, hdr)
        log.set_offsets(head=256, apply=256, commit=256, end=pos, tail=tail, old_end=pos)
        data = {O.SEND: O.cmd_image(bytes(range(64))), O.NOOP: b"", O.HEAD: (77).to_bytes(8, "little"),
                O.CONFIG: O.cid_image(5)}[typ]
        idx = log.append(2, 9, 0x0203, typ, data)
        res.append((idx, log.offsets(), log.image()))
        log.close()
    assert res[0][0] == res[1][0]
    assert res[0][1] == res[1][1]
    assert np.array_equal(res[0][2], res[1][2])


def test_wrap_into_head_zero_is_full(ref, orc):
    """H11(iii): wrapping while head == 0 reports full and leaves end = 0."""
    orc.set_rules(O.RULES_REFERENCE)
    out = []
    for oracle in (ref, orc):
        log = O.Log(oracle, 4096)
        rets = [log.append(1, i + 1, 1, O.SEND, O.cmd_image(b"a" * 70)) for i in range(40)]
        out.append((rets, log.offsets(), log.image()))
        log.close()
    assert out[0][0] == out[1][0]
    assert out[0][1] == out[1][1]
    assert np.array_equal(out[0][2], out[1][2])
    assert 0 in out[0][0]


def cluster_script(oracle, n, stream, length, schedule_seed):
    """Drive the cluster with a randomised (but seeded) interleaving of follower
    replication / ack steps; return everything observable."""
    c = O.Cluster(oracle, n, leader=0, term=1, length=length)
    rng = np.random.default_rng(schedule_seed)
    c.prologue()
    commits = []
    for k, (typ, clt, rid, payload) in enumerate(stream):
        c.submit(typ, clt, rid, O.cmd_image(payload))
        if rng.random() < 0.6:
            c.leader_persist()
            order = rng.permutation(n)
            for i in order:
                if rng.random() < 0.8:
                    c.replicate(int(i))
                if rng.random() < 0.7:
                    c.follower_persist(int(i))
            if c.commit_scan():
                commits.append(c.offsets(0)["commit"])
            for i in range(n):
                if rng.random() < 0.5:
                    c.push_commit(i)
                    c.apply(i)
    for _ in range(3):
        c.round()
    commits.append(c.offsets(0)["commit"])
    obs = dict(
        offsets=[c.offsets(i) for i in range(n)],
        images=[c.image(i) for i in range(n)],
        applied=[c.applied(i) for i in range(n)],
        commits=commits,
        store=[c.store_cmd_calls(i) for i in range(n)],
        update_state=c.update_state_calls(),
        bytes_rep=c.bytes_replicated(),
    )
    c.close()
    return obs


@pytest.mark.parametrize("n", [1, 3, 5, 7])
def test_cluster_steps_match(ref, orc, n):
    orc.set_rules(O.RULES_REFERENCE)
    stream = S.ragged_stream(400, 200, conns=4, seed=n, close_every=40)
    a = cluster_script(ref, n, stream, 1 << 20, 5)
    b = cluster_script(orc, n, stream, 1 << 20, 5)
    assert a["offsets"] == b["offsets"]
    assert a["applied"] == b["applied"]
    assert a["commits"] == b["commits"]
    assert a["store"] == b["store"] and a["update_state"] == b["update_state"]
    assert a["bytes_rep"] == b["bytes_rep"]
    for x, y in zip(a["images"], b["images"]):
        assert np.array_equal(x, y)
    # invariants of SURVEY.md s8a: commit == end on the leader at quiescence,
    # apply order == log order, every follower applied the same sequence
    assert a["offsets"][0]["commit"] == a["offsets"][0]["end"]
    idxs = [t[0] for t in a["applied"][0]]
    assert idxs == sorted(idxs)
    for i in range(1, n):
        assert a["applied"][i] == a["applied"][0]


def test_cluster_wrap_and_prune_match(ref, orc):
    """Ring of 16 KiB, pruning by HEAD entries between rounds, several laps."""
    orc.set_rules(O.RULES_REFERENCE)
    stream = S.ragged_stream(900, 150, conns=2, seed=77)

    def run(oracle):
        c = O.Cluster(oracle, 3, leader=0, term=1, length=16384)
        c.prologue()
        heads = []
        cido = [(O.u64)(0) for _ in range(3)]
        for k, (typ, clt, rid, payload) in enumerate(stream):
            c.submit(typ, clt, rid, O.cmd_image(payload))
            if k % 5 == 4:
                c.round()
                heads.append(c.prune())
                c.round()
                for i in (1, 2):
                    c.poll_head(i, cido[i])
        c.round()
        obs = ([c.offsets(i) for i in range(3)], [c.image(i) for i in range(3)],
               [c.applied(i) for i in range(3)], heads)
        c.close()
        return obs

    a, b = run(ref), run(orc)
    assert a[0] == b[0] and a[2] == b[2] and a[3] == b[3]
    for x, y in zip(a[1], b[1]):
        assert np.array_equal(x, y)
    assert any(h for h in a[3])
