"""The host pumps of libapus_dare.so on a box without a GPU: tests/hostlogic/pump_harness.c includes the product source
(apus_b200/csrc/dare_entry.c) and stands a test double behind it.

follower_pump (apply_committed_entries, follower branch, dare_server.c:1815-1967): the replica's ring is fed in stages
produced by the ORACLE's cluster (a 3-replica group, a follower's copy of the ring and its commit offset every few
requests, twenty laps of a 16 KiB ring with pruning -- HEAD entries, ghost headers, headers that do not fit before the
ring's end); the pump must call store_cmd + do_action once per request entry, in log order, with the request's bytes,
skip CONFIG / HEAD entries, report exactly the committed offsets as applied (never beyond), and cope with a read buffer
smaller than the committed range.

leader_pump (get_tailq_message + persist_new_entries + the leader branch of apply_committed_entries): application threads
enqueue under tailq_lock like proxy.c:108-161 and spin until update_state has counted their request; the engine must be
handed every request exactly once, per connection in the order it was enqueued, store_cmd once per entry with the
reference's record image, update_state once per committed request."""
import os
import subprocess

import numpy as np
import pytest

import orc as O
import streams as S

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("pumps") / "pump_harness")
    subprocess.run(["gcc", "-O1", "-g", "-std=gnu99", "-Wall", "-Wno-unused-function", f"-I{ROOT}/include", "-o", exe,
                    os.path.join(HERE, "hostlogic", "pump_harness.c"), "-lpthread"], check=True)
    return exe


def fnv(b):
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.parametrize("read_cap", [0, 1500])
def test_follower_pump_replays_the_oracles_log(orc, harness, tmp_path, read_cap):
    L, n = 16384, 3
    stream = S.ragged_stream(1500, 400, seed=11, close_every=70)
    orc.set_rules(O.RULES_ENGINE)
    c = O.Cluster(orc, n, leader=0, term=1, length=L)
    c.prologue()
    stages, cido = [], (O.u64)(0)
    try:
        for k, (typ, clt, rid, payload) in enumerate(stream):
            assert c.submit(typ, clt, rid, O.cmd_image(payload)), k
            if k % 6 == 5 or k == len(stream) - 1:
                c.round(); c.round()
                stages.append((c.image(1), c.offsets(1)["commit"]))
                if k % 12 == 11:
                    c.prune(); c.round(); c.round()
                    c.poll_head(1, cido); c.poll_head(2, cido)
                    stages.append((c.image(1), c.offsets(1)["commit"]))
        laps_bytes = sum(64 + 2 + len(p) for _, _, _, p in stream)
        assert laps_bytes > 15 * L                                    # the ring really went round many times
    finally:
        c.close()
        orc.set_rules(O.RULES_REFERENCE)
    for k, (img, commit) in enumerate(stages):
        img.tofile(tmp_path / f"stage{k}.bin")
        (tmp_path / f"stage{k}.commit").write_text(f"{commit if commit != L else 0}\n")
    out = subprocess.run([harness, "follower", str(tmp_path), str(L), str(len(stages)), str(read_cap)], capture_output=True,
                         text=True, timeout=120)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "FATAL" not in out.stdout, out.stdout[-1500:]
    calls = (tmp_path / "calls.txt").read_text().split("\n")
    acts = [l.split() for l in calls if l.startswith("A ")]
    stores = [l.split() for l in calls if l.startswith("S ")]
    applied = [int(l.split()[1]) for l in calls if l.startswith("P ")]
    end = [l for l in calls if l.startswith("END")][0]
    want = [(clt, typ, payload) for typ, clt, rid, payload in stream]
    assert len(acts) == len(want) == len(stores)
    for a, s, (clt, typ, payload) in zip(acts, stores, want):
        assert (int(a[1]), int(a[2]), int(a[3])) == (clt, typ, len(payload))
        assert int(a[4], 16) == fnv(payload)
        assert (int(s[1]), int(s[2]), int(s[3])) == (clt, typ, len(payload))     # record image: clt_id, type, cmd.len
    # the offsets reported as applied: every stage's commit offset is reached, in order, and nothing beyond it is reported
    commits = [cm if cm != L else 0 for _, cm in stages]
    it = iter(applied)
    assert all(any(a == cm for a in it) for cm in dict.fromkeys(commits)), "a committed offset was never reported applied"
    assert applied[-1] == commits[-1]
    assert "rc=0" in end
    if read_cap:
        assert "capped=0" not in end, end                             # the small buffer really cut ranges
    assert "two_piece=0" not in end, end                              # ranges that wrap were read in two pieces


@pytest.mark.parametrize("threads,nreq,plen", [(1, 300, 64), (6, 400, 100)])
def test_leader_pump_order_callbacks(harness, tmp_path, threads, nreq, plen):
    out = subprocess.run([harness, "leader", str(tmp_path), str(threads), str(nreq), str(plen)], capture_output=True, text=True,
                         timeout=120)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    calls = (tmp_path / "calls.txt").read_text().split("\n")
    subs = [l.split() for l in calls if l.startswith("T ")]
    stores = [l.split() for l in calls if l.startswith("S ")]
    end = [l for l in calls if l.startswith("END")][0]
    total = threads * (nreq + 2)
    assert f"tickets={total} " in end and f"update_state={total}" in end   # every CONNECT / SEND / CLOSE, once
    assert [int(s[1]) for s in subs] == list(range(1, total + 1))
    per = {}
    for _, _tk, typ, conn, req, ln, h in subs:
        per.setdefault(int(conn), []).append((int(typ), int(req), int(ln), int(h, 16)))
    assert sorted(per) == list(range(threads))
    for conn, seq in per.items():
        want = [(S.CONNECT, 1, 0, fnv(b""))]
        want += [(S.SEND, 1 + i, plen, fnv(bytes((conn * 131 + i * 31 + k) & 0xFF for k in range(plen)))) for i in range(1, nreq + 1)]
        want += [(S.CLOSE, nreq + 2, 0, fnv(b""))]
        assert seq == want, conn                                      # per connection: enqueue order, consecutive req_ids
    # store_cmd: one record per entry, the reference's image (clt_id, type, cmd.len) -- dare_server.c:1802
    assert len(stores) == total
    assert sorted((int(s[1]), int(s[2]), int(s[3])) for s in stores) == sorted(
        (c, t, ln) for c, seq in per.items() for t, _, ln, _ in seq)
