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


def test_failure_detector_removal_then_join(harness, tmp_path):
    """leader_check_followers + leader_serve_join + join_group, two processes sharing a rendezvous directory: p2 of a
    group of three stops beating after 60 ms; a replacement asks to join the emptied slot."""
    import time
    lead = subprocess.Popen([harness, "membership", str(tmp_path), "60", "0", "0"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    t0 = time.time()
    while not (tmp_path / "removed").exists():
        assert time.time() - t0 < 20 and lead.poll() is None, "the failure detector never removed p2"
        time.sleep(0.005)
    removed_after_ms = int((tmp_path / "removed").read_text()) / 1e3
    # not before death + hb_timeout (10 periods, floored at 20 ms; the last beat was SEEN up to one 5 ms scan earlier); soon after
    assert 60 + 20 - 10 <= removed_after_ms < 5000, removed_after_ms
    join = subprocess.run([harness, "joiner", str(tmp_path), "0", "0", "0"], capture_output=True, text=True, timeout=60)
    lout = lead.communicate(timeout=60)[0]
    assert join.returncode == 0 and lead.returncode == 0, join.stdout + join.stderr + lout
    calls = (tmp_path / "calls.txt").read_text().split("\n")
    # the election winner's prologue (all three), the removal of p2, the re-admission of p2 -- in this order
    configs = [l for l in calls if l.startswith("CONFIG")]
    assert [c.split()[2:] for c in configs] == [["size=3", "mask=7"], ["size=3", "mask=3"], ["size=3", "mask=7"]], configs
    order = ["CONFIG" if l.startswith("CONFIG") else " ".join(l.split()[:2]) for l in calls if l.split()[:1] and l.split()[0] in ("D", "C", "J", "CONFIG")]
    assert order == ["CONFIG", "D 2", "CONFIG", "C 2", "J 2", "CONFIG"], order
    sid = int([l for l in calls if l.startswith("J ")][0].split()[2])
    assert sid == (5 << 9) | (1 << 8) | 0                                    # the joiner is adjusted under [term 5 | L | p0]
    end = [l for l in calls if l.startswith("END")][0]
    assert "live_mask=7" in end and "removed_mask=0" in end
    # the kernel was stopped and relaunched around each change of the peer set (removal, join)
    assert "stops=2" in end and "launches=2" in end, end
    assert "REMOVE SERVER p2" in lout and "JOIN request from p2" in lout and "p2 joined" in lout
    # the joiner: the leader's answer, the snapshot through apply_db_snapshot, where to follow from
    j = (tmp_path / "joiner_calls.txt").read_text().split("\n")
    snap = bytes((i * 7 + 3) & 0xFF for i in range(1000))
    assert [l for l in j if l.startswith("SNAP")] == [f"SNAP 1000 {fnv(snap):016x}"]
    jend = [l for l in j if l.startswith("END")][0]
    assert "rc=0 leader=0 term=5 apply=2624 next_idx=42 live_mask=7" in jend, jend
    # the handle the leader mapped is the one the joiner published
    assert [l for l in calls if l.startswith("C 2")][0].split()[2] == jend.split("handle=")[1]


@pytest.mark.parametrize("text,want", [
    # target/nodes.local.cfg: one setting per line, commented-out alternatives above them
    ('db_name = "node_test";\nport = 8888;\ndare_global_config = {\n    #hb_period = 0.001;\n    #elec_timeout_low = 10000;\n'
     '    hb_period = 0.01;\n    elec_timeout_low = 100000;\n    elec_timeout_high = 300000;\n    rc_info_period = 0.05;\n};\n',
     (0.01, 100000, 300000)),
    # benchmarks/run_gpu.sh: several settings on one line, the group spread over two
    ('port = 8890;\ndare_global_config = { hb_period = 0.001; elec_timeout_low = 10000; elec_timeout_high = 30000;\n'
     '                       retransmit_period = 0.02; rc_info_period = 0.01; log_pruning_period = 0.03; };\n', (0.001, 10000, 30000)),
    # libconfig accepts ':' and an L suffix; comments of all three kinds; a setting of the same name OUTSIDE the group does not count
    ('hb_period = 7.0; // not ours\ndare_global_config : { /* hb_period = 9; */ hb_period : 0.002; elec_timeout_low = 2000L;\n'
     '  elec_timeout_high=6000L; # done\n};\n', (0.002, 2000, 6000)),
    # no group: the defaults of config-dare.c stay
    ('db_name = "x";\nport = 1;\n', (0.01, 100000, 300000)),
])
def test_dare_global_config_is_read_like_libconfig(harness, tmp_path, text, want):
    """read_dare_config (the three values of config-dare.c:12-52 this engine uses) must accept what libconfig accepts, not
    only the layout of target/nodes.local.cfg -- benchmarks/run_gpu.sh writes several settings per line."""
    cfg = tmp_path / "node.cfg"
    cfg.write_text(text)
    out = subprocess.run([harness, "config", str(tmp_path), str(cfg), "0", "0"], capture_output=True, text=True, timeout=30)
    assert out.returncode == 0, out.stdout + out.stderr
    hb, lo, hi = want
    assert f"hb_period={hb:g} elec_timeout_low={lo} elec_timeout_high={hi}" in out.stdout, out.stdout
