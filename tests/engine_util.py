"""Helpers for the GPU parity tests: run one request stream through the CUDA
engine (via the C ABI) and through the CPU oracle, then compare everything
observable (SURVEY.md s8c "Masking rule for parity")."""
import numpy as np

import orc as O


def oracle_cluster(orc, n, length, stream, rules=O.RULES_ENGINE, prologue=True, leader=0, term=1):
    orc.set_rules(rules)
    c = O.Cluster(orc, n, leader=leader, term=term, length=length)
    if prologue:
        c.prologue()
    for typ, clt, rid, payload in stream:
        idx = c.submit(typ, clt, rid, O.cmd_image(payload))
        assert idx != 0, "oracle refused an append (ring full): shorten the stream or prune"
    for _ in range(2):
        c.round()
    return c


def compare_group_to_oracle(group, c, exact=True):
    """group: apus_b200.Group at quiescence; c: orc.Cluster at quiescence."""
    n, L = group.n, group.replicas[0].log_len
    lead = group.leader_idx
    oo = [c.offsets(i) for i in range(n)]
    eo = [r.offsets() for r in group.replicas]
    for i in range(n):
        assert eo[i]["len"] == oo[i]["len"] == L
        assert eo[i]["end"] == oo[i]["end"], f"replica {i} end {eo[i]} vs {oo[i]}"
        assert eo[i]["commit"] == oo[i]["commit"], f"replica {i} commit {eo[i]} vs {oo[i]}"
        assert eo[i]["apply"] == oo[i]["apply"], f"replica {i} apply {eo[i]} vs {oo[i]}"
    assert eo[lead]["tail"] == oo[lead]["tail"]
    assert eo[lead]["commit"] == eo[lead]["end"], "leader: commit == end at quiescence"
    for i in range(n):
        ei = group.replicas[i].image()
        oi = c.image(i)
        if exact:
            if not np.array_equal(ei, oi):
                d = np.nonzero(ei != oi)[0]
                raise AssertionError(f"replica {i}: {len(d)} bytes differ, first at {int(d[0])} "
                                     f"(engine {ei[d[0]]} oracle {oi[d[0]]}); offsets {eo[i]}")
        else:
            ents = O.walk_entries(oi, 0, oo[i]["end"], L) if oo[i]["end"] != L else []
            assert np.array_equal(O.mask_replies(ei, ents), O.mask_replies(oi, ents))
    # invariant I7: at quiescence the leader holds reply[i]==1 for every follower,
    # follower i holds at least its own byte
    limg = group.replicas[lead].image()
    end = eo[lead]["end"]
    if end != L and eo[lead]["head"] == 0 and end > 0:
        ents = O.walk_entries(limg, 0, end, L)
        for off, _ in ents[-64:]:
            for i in range(n):
                if i != lead:
                    assert limg[off + 28 + i] == 1
    return eo, oo
