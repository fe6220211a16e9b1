"""Pins the oracle's CLUSTER restatement (oracle/cluster_sim.inc: submit, replicate, ack, commit scan, apply,
HEAD-entry pruning) against the reference ITSELF: the reference's unmodified election / replication / commit code
(src/dare/dare_server.c, dare_ibv_rc.c, dare_ibv_ud.c, dare_ibv.c) and proxy.c run here as N processes on the
verbs shim (oracle/verbs_shim; build recipe oracle/build_refapp.sh), an application driver issues a deterministic
request stream through proxy_on_accept/read/close, and the log the reference leaves behind is compared with the
log the oracle computes for the same stream -- every byte, reply bytes included."""
import hashlib

import numpy as np
import pytest

import orc as O
import refstack as R
import streams as S

pytestmark = [pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref_stack.so absent (needs /root/reference)"),
              pytest.mark.timeout(300)]


def oracle_for(orc, rr, n, nconn, nreq, plen):
    orc.set_rules(O.RULES_REFERENCE)
    c = O.Cluster(orc, n, leader=rr["leader"], term=rr["term"], length=O.LOG_SIZE)
    c.prologue()
    for typ, clt, rid, payload in R.expected_stream(rr["leader"], nconn, nreq, plen):
        assert c.submit(typ, clt, rid, O.cmd_image(payload))
    for _ in range(2):
        c.round()
    return c


def check_replay(rr, nconn, nreq, plen):
    """followers replayed every connection's byte stream, in order, into their local application"""
    want = []
    for c in range(nconn):
        h = hashlib.sha256()
        for typ, clt, rid, payload in R.expected_stream(rr["leader"], nconn, nreq, plen):
            if typ == S.SEND and (clt & 0xFF) == c:
                h.update(payload)
        want.append(h.hexdigest())
    for r in rr["results"]:
        if not r["leader"]:
            assert r["replay"]["conns"] == nconn
            assert sorted(r["replay"]["sha"]) == sorted(want)


@pytest.mark.parametrize("n,nconn,nreq,plen", [(3, 2, 300, 64), (5, 3, 200, 128), (3, 1, 120, -3000), (7, 4, 400, 64)])
def test_reference_log_equals_oracle_log(orc, n, nconn, nreq, plen):
    """log_pruning_period is set out of reach, so the log holds exactly CONFIG + the stream."""
    _log_equals_oracle(orc, n, nconn, nreq, plen)


@pytest.mark.parametrize("n,nconn,nreq,plen", [(5, 3, 200, 128), (3, 1, 120, -3000)])
def test_reference_log_equals_oracle_log_shm_transport(orc, n, nconn, nreq, plen):
    """The shim's shm transport (the log re-backed by a shared mapping, RDMA WRITE = memcpy: what bench.py's reference arm
    also times) leaves the same logs as the process_vm transport."""
    _log_equals_oracle(orc, n, nconn, nreq, plen, transport="shm")


def _log_equals_oracle(orc, n, nconn, nreq, plen, transport=None):
    rr = R.run(n, nconn, nreq, plen, prune=1000.0, transport=transport)
    c = oracle_for(orc, rr, n, nconn, nreq, plen)
    lead = rr["leader"]
    oo = c.offsets(lead)
    for i in range(n):
        ro = rr["results"][i]["offsets"]
        assert ro["end"] == oo["end"] and ro["head"] == 0 and ro["len"] == O.LOG_SIZE, (i, ro, oo, rr["term"], lead)
        assert ro["commit"] == ro["end"] == ro["apply"], ro
        if i == lead:
            assert ro["tail"] == oo["tail"]
        img, want = rr["images"][i], c.image(i, 0, oo["end"])
        assert len(img) == oo["end"]
        ents = O.walk_entries(want, 0, oo["end"], O.LOG_SIZE)
        if i != lead:
            # A follower's copy of an entry carries whatever reply bytes the LEADER's copy held at the instant it was
            # replicated (acks of faster followers; timing, the H5 mask of SURVEY.md s8c) plus its own ack (I7).
            for off, _ in ents:
                rep = img[off + 28: off + 41]
                assert rep[i] == 1 and rep[lead] == 0 and not rep[n:].any()
            img, want = O.mask_replies(img, ents), O.mask_replies(want, ents)
        if not np.array_equal(img, want):                 # the leader's copy: every byte, reply bytes included
            dd = np.nonzero(img != want)[0]
            raise AssertionError(f"replica {i} (leader {lead}, term {rr['term']}): {len(dd)} bytes differ from the oracle, "
                                 f"first at {int(dd[0])}: reference {img[dd[0]]} oracle {want[dd[0]]}")
    if n > 1:
        check_replay(rr, nconn, nreq, plen)
    # what proxy.c / db-interface.c counted (the callbacks the engine entry must reproduce):
    # store_cmd once per entry on every replica -- 4 B records for CONNECT/CLOSE, 24 B for SEND whatever its payload;
    # update_state once per committed request on the leader
    for r in rr["results"]:
        assert r["proxy"]["records_len"] == 8 * nconn + 24 * nreq, r["proxy"]
        if r["leader"]:
            assert r["proxy"]["highest_rec"] == r["proxy"]["cur_rec"] == 2 * nconn + nreq
    c.close()


def test_reference_pruning_matches_oracle_rules(orc):
    """With the stock log_pruning_period (0.05 s) the reference leader interleaves HEAD entries at timer-dependent
    places.  Their PLACEMENT is timing; their content and consequences are rules the oracle restates:
    a HEAD entry carries the new head, which is an earlier entry boundary, larger than the previous head, never two
    HEAD entries in a row (dare_server.c:1996-2067, dare_log.h:472-478); followers adopt it (dare_server.c:2163-2186).
    Rebuilding the log with the oracle's append -- the stream plus HEAD entries where the reference put them --
    must reproduce the reference's bytes."""
    n, nconn, nreq, plen = 3, 2, 6000, 64
    for attempt in range(3):
        rr = R.run(n, nconn, nreq, plen, prune=0.005)
        lead = rr["leader"]
        img = rr["images"][lead]
        end = rr["results"][lead]["offsets"]["end"]
        ents = O.walk_entries(img, 0, end, O.LOG_SIZE)
        # whether the timer finds something to prune is timing (it needs every follower's apply offset to have moved since
        # the last HEAD entry): a run without two HEAD entries says nothing about the rules, take another one
        if sum(1 for off, _ in ents if int(img[off + 26]) == O.HEAD) >= 2:
            break
    bounds = {off for off, _ in ents}
    stream = iter(R.expected_stream(lead, nconn, nreq, plen))
    orc.set_rules(O.RULES_REFERENCE)
    log = O.Log(orc, O.LOG_SIZE)
    heads, prev_head, prev_was_head = [], 0, False
    for k, (off, stride) in enumerate(ents):
        typ = int(img[off + 26])
        if k == 0:
            assert typ == O.CONFIG
            assert log.append(rr["term"], 0, 0, O.CONFIG, O.cid_image(n)) == 1
        elif typ == O.HEAD:
            h = int(np.frombuffer(img[off + 48: off + 56].tobytes(), dtype="<u8")[0])
            assert h in bounds and h < off, "the new head is an earlier entry boundary"
            assert h > prev_head and not prev_was_head
            heads.append(h)
            prev_head = h
            assert log.append(rr["term"], 0, 0, O.HEAD, h.to_bytes(8, "little")) == k + 1
        else:
            t, clt, rid, payload = next(stream)
            assert typ == t
            assert log.append(rr["term"], rid, clt, t, O.cmd_image(payload)) == k + 1
        prev_was_head = typ == O.HEAD
    assert next(stream, None) is None, "every request is in the log"
    assert len(heads) >= 2, "the run was long enough to prune"
    want = log.image(0, end)
    got = O.mask_replies(img, ents)
    for off, _ in ents:                       # the single-log oracle has no followers and no leader stamp
        want[off + 27] = lead
    assert np.array_equal(got, O.mask_replies(want, ents))
    for i in range(n):
        ro = rr["results"][i]["offsets"]
        assert ro["end"] == end and ro["commit"] == end
        assert ro["head"] in heads[-2:], "every replica adopted one of the last heads"
        assert np.array_equal(O.mask_replies(rr["images"][i], ents), got)
    log.close()
    check_replay(rr, nconn, nreq, plen)


def test_reference_reply_bytes_invariant_I7():
    """At quiescence the leader's copy of an entry holds reply[f] == 1 for every follower f and follower f's own copy
    holds reply[f] == 1 (dare_ibv_rc.c:1838-1839) -- the bytes the GPU engine also deposits."""
    n = 3
    rr = R.run(n, 1, 100, 64, prune=1000.0)
    lead = rr["leader"]
    end = rr["results"][lead]["offsets"]["end"]
    ents = O.walk_entries(rr["images"][lead], 0, end, O.LOG_SIZE)
    for i in range(n):
        img = rr["images"][i]
        for off, _ in ents:
            rep = img[off + 28: off + 41]
            if i == lead:
                assert all(rep[f] == 1 for f in range(n) if f != lead) and rep[lead] == 0 and not rep[n:].any()
            else:
                assert rep[i] == 1 and rep[lead] == 0
            assert img[off + 27] == lead          # sender stamped before replication (dare_server.c:1803)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_reference_log_equals_oracle_log_random_shapes(orc, seed):
    """Group size, connection count, request count and payload regime drawn from a seeded generator."""
    rng = np.random.default_rng(0xA5A5 + seed)
    n = int(rng.choice([3, 5]))
    nconn = int(rng.integers(1, 7))
    nreq = int(rng.integers(50, 500))
    plen = int(rng.choice([1, 17, 64, 255, 1024, -200, -1500, -9000]))
    rr = R.run(n, nconn, nreq, plen, prune=1000.0)
    c = oracle_for(orc, rr, n, nconn, nreq, plen)
    lead = rr["leader"]
    end = c.offsets(lead)["end"]
    ents = O.walk_entries(c.image(lead, 0, end), 0, end, O.LOG_SIZE)
    assert len(ents) == 1 + 2 * nconn + nreq
    for i in range(n):
        assert rr["results"][i]["offsets"]["end"] == end, (n, nconn, nreq, plen)
        img, want = rr["images"][i], c.image(i, 0, end)
        if i != lead:
            img, want = O.mask_replies(img, ents), O.mask_replies(want, ents)
        assert np.array_equal(img, want), (i, lead, n, nconn, nreq, plen)
    check_replay(rr, nconn, nreq, plen)
    c.close()


def test_reference_log_with_concurrent_application_threads(orc):
    """Four application threads (memcached-style) race into proxy.c's admission lock, so the global order is whatever
    the lock produced.  The order is read back from the reference's log, the oracle replays exactly that order and
    must produce the same bytes; per connection the requests are all there, in order, with consecutive req_ids
    (proxy.c:121-133), and followers replay each connection's stream in order."""
    n, nconn, nreq, plen, threads = 3, 8, 800, 96, 4
    rr = R.run(n, nconn, nreq, plen, threads=threads, prune=1000.0)
    lead = rr["leader"]
    img = rr["images"][lead]
    end = rr["results"][lead]["offsets"]["end"]
    ents = O.walk_entries(img, 0, end, O.LOG_SIZE)
    assert len(ents) == 1 + 2 * nconn + nreq
    orc.set_rules(O.RULES_REFERENCE)
    c = O.Cluster(orc, n, leader=lead, term=rr["term"], length=O.LOG_SIZE)
    c.prologue()
    seen = {}
    for off, stride in ents[1:]:
        typ = int(img[off + 26])
        clt = int(img[off + 24]) | (int(img[off + 25]) << 8)
        rid = int(np.frombuffer(img[off + 16: off + 24].tobytes(), dtype="<u8")[0])
        ln = int(img[off + 48]) | (int(img[off + 49]) << 8)
        payload = img[off + 50: off + 50 + ln].tobytes()
        assert (clt >> 8) == lead
        assert rid == seen.get(clt, 0) + 1, "req_ids of one connection are consecutive, in log order"
        seen[clt] = rid
        if typ == S.SEND:
            assert ln == plen and payload[1] == (payload[0] + 1) & 0xFF
        assert c.submit(typ, clt, rid, O.cmd_image(payload))
    c.round(); c.round()
    assert len(seen) == nconn and sum(seen.values()) == 2 * nconn + nreq
    for i in range(n):
        got, want = rr["images"][i], c.image(i, 0, end)
        if i != lead:
            got, want = O.mask_replies(got, ents), O.mask_replies(want, ents)
        assert np.array_equal(got, want), i
    # followers: every connection's byte stream arrived in order
    want_sha = {}
    for off, stride in ents[1:]:
        if int(img[off + 26]) == S.SEND:
            clt = int(img[off + 24]) | (int(img[off + 25]) << 8)
            ln = int(img[off + 48]) | (int(img[off + 49]) << 8)
            want_sha.setdefault(clt, hashlib.sha256()).update(img[off + 50: off + 50 + ln].tobytes())
    for r in rr["results"]:
        if not r["leader"]:
            assert sorted(r["replay"]["sha"]) == sorted(h.hexdigest() for h in want_sha.values())
    c.close()
