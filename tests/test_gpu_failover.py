"""BASELINE config 5 on the GPU engine: kill the leader process under load; the survivors' failure detector (heartbeat
words written by the leader kernel), election and log adjustment (control plane on NVLink words) must leave what the
reference's own stack leaves in this scenario (tools/refstack_failover.py, profiles/r1_refstack_failover_buildbox.txt):
on every survivor the same log -- the prefix the old leader had replicated, then the blank CONFIG entry every election
winner appends and the CONFIG entry that removes the dead server, both of the new term and stamped with the new leader's
index -- and followers that replay every connection's bytes in order, nothing lost that a client saw committed."""
import os
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tools"))
REFPROXY = os.path.join(ROOT, "oracle", "_ref", "libref_proxy.so")


@pytest.mark.parametrize("n", [3, 5])
def test_leader_failover(n):
    import __graft_entry__ as g
    g.build()
    if not os.path.exists(REFPROXY):
        pytest.skip("oracle/_ref/libref_proxy.so absent (built only where /root/reference exists)")
    import apus_b200
    import failover_drill as FD
    nd = max(1, apus_b200.lib().apus_device_count())
    nconn, nreq2, plen = 3, 600, 64
    r = FD.run(n=n, nconn=nconn, nreq2=nreq2, plen=plen, kill_after_s=0.5, spread=False, ndev=nd)
    res, lead = r["results"], r["new_leader"]
    T = r["term"]
    assert lead in res and T >= 2              # (a split first round costs a term, as in the reference's own runs: term 4 there)
    assert "] LEADER" in r["logs"][lead]                       # the line reconf_bench.sh greps for
    # every survivor that is part of the new configuration holds the same entries (reply bytes masked), committed up to
    # the same end; a survivor the winner removed (it answered too late: process contexts on one GPU are time-sliced)
    # stays behind -- but the new leader and its followers are a majority of the group
    ents = {i: [(e["idx"], e["term"], e["type"], e["sender"], e["sha"]) for e in res[i]["entries"]] for i in res}
    ref = ents[lead]
    removed_by_leader = {i for i in range(n) if f"REMOVE SERVER p{i}" in r["logs"][lead]}
    members = [i for i in res if i == lead or i not in removed_by_leader]
    assert len(members) >= n // 2 + 1, (members, removed_by_leader, r.get("missing"))
    for i in members:
        assert ents[i] == ref, f"survivor {i} differs from the new leader"
    res = {i: res[i] for i in members}
    idx = [e[0] for e in ref]
    assert idx == list(range(1, len(idx) + 1))
    # structure: term-1 entries stamped by p0, then two CONFIG entries of term 2 stamped by the new leader, then its requests
    first2 = next(k for k, e in enumerate(ref) if e[1] == T)
    assert all(e[1] == 1 and e[3] == 0 for e in ref[:first2])
    assert all(e[1] == T and e[3] == lead for e in ref[first2:])
    c1, c2 = res[lead]["entries"][first2], res[lead]["entries"][first2 + 1]
    assert c1["type"] == 2 and c2["type"] == 2
    full = (1 << n) - 1
    assert int.from_bytes(bytes.fromhex(c1["data"])[12:16], "little") == full
    gone = 1
    for i in removed_by_leader:
        gone |= 1 << i
    assert int.from_bytes(bytes.fromhex(c2["data"])[12:16], "little") == full & ~gone      # p0 (and late survivors) removed
    tail = res[lead]["entries"][first2 + 2:]
    assert [e["type"] for e in tail] == [4] * nconn + [5] * nreq2 + [6] * nconn
    # nothing a client saw committed was lost: phase-1 SENDs in the log >= the progress the old leader reported before the kill
    p1_sends = sum(1 for e in res[lead]["entries"][:first2] if e["type"] == 5)
    assert p1_sends >= r["requests_before_kill"] - 64         # (progress is written every 64 requests)
    # followers replayed phase 2 completely, per connection, in order
    import hashlib
    expect = sorted(hashlib.sha256(b"".join(bytes(((i * 31 + k) & 0xFF) for k in range(plen)) for i in range(nreq2) if i % nconn == c)).hexdigest()
                    for c in range(nconn))
    for i in res:
        if i == lead:
            continue
        shas = sorted(x["sha"] for x in res[i]["replay"] if x["bytes"] == (nreq2 // nconn + (1 if nreq2 % nconn else 0)) * plen
                      or x["bytes"] == (nreq2 // nconn) * plen)
        assert all(s in shas for s in expect), f"follower {i} did not replay phase 2 exactly"
    print(f"failover: p{lead} leads term {r['term']}; kill -> LEADER line {r['recovery_ms_kill_to_leader_line']} ms, "
          f"kill -> first commit {r['recovery_ms_kill_to_first_commit']} ms")
