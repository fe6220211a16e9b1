"""The election policy of libapus_dare.so (apus_b200/csrc/dare_entry.c: elect -- the restatement of the reference's
start_election / poll_vote_requests / poll_vote_count, dare_server.c:1264-1743) as N survivor PROCESSES on a box without
a GPU: tests/election/harness.c includes the product source and links it against tests/election/mock_engine.c, where the
control words (SID, vote requests, vote acks, leader announcement) live in files mapped by every process instead of HBM.

Checked, over repeated runs with the randomised election timeouts the reference uses: exactly one leader per winning
term; every other survivor follows it in that term; nobody whose log is behind a voter's gets that voter's vote
(up-to-date rule on (term, idx), dare_server.c:1640-1652) -- so the winner holds the most advanced log among a majority;
followers are adjusted to the winner's log; servers that never answered are disconnected by the winner."""
import os
import re
import struct
import subprocess
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("election") / "harness")
    subprocess.run(["gcc", "-O1", "-g", "-std=gnu99", "-Wall", "-Wno-unused-function", f"-I{ROOT}/include", "-o", exe,
                    os.path.join(HERE, "election", "harness.c"), os.path.join(HERE, "election", "mock_engine.c"), "-lpthread"],
                   check=True)
    return exe


BLK = struct.Struct("<4Q16Q" + "8Q" * 13 + "4Q5Q")      # mock_engine.c: blk_t


def read_blk(d, i):
    raw = open(os.path.join(d, f"ctl{i}.bin"), "rb").read()[:BLK.size]
    f = BLK.unpack(raw)
    tail = f[4 + 16 + 8 * 13:]
    return dict(sid=f[0], leader_sid=f[1], adj_end=f[2], adj_count=f[3], last_idx=tail[0], last_term=tail[1], commit=tail[2],
                end=tail[3], role_leader=tail[4], role_term=tail[5], launches=tail[6], adjusted_by=tail[7] - 1, disconnected=tail[8])


def run_election(harness, n, dead, logs, absent=(), elec=(2000, 6000), timeout=60, env=None):
    """logs: {idx: (last_idx, last_term)}; returns ({idx: result dict}, dir-free block dicts)"""
    with tempfile.TemporaryDirectory() as d:
        procs = {}
        env = dict(os.environ, **(env or {}))
        for i in range(n):
            if i == dead or i in absent:
                continue
            li, lt = logs[i]
            procs[i] = subprocess.Popen([harness, d, str(i), str(n), str(dead), "1", str(li), str(lt), str(64 * li), str(64 * li),
                                         str(elec[0]), str(elec[1])], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                                        env=env)
        # `up<i>` of servers that never start: the harness only waits for the ones that exist
        for i in absent:
            open(os.path.join(d, f"up{i}"), "w").close()
        res = {}
        for i, p in procs.items():
            out = p.communicate(timeout=timeout)[0]
            m = re.search(r"RESULT idx=(\d+) role=(\w+) leader=(\d+) term=(\d+) rc=(\d+)", out)
            assert m, f"replica {i} printed no result:\n{out[-1500:]}"
            res[i] = dict(role=m.group(2), leader=int(m.group(3)), term=int(m.group(4)), rc=int(m.group(5)), log=out)
        blks = {i: read_blk(d, i) for i in procs}
    return res, blks


def check(res, blks, logs, n):
    leaders = [i for i, r in res.items() if r["role"] == "leader"]
    top_term = max(r["term"] for r in res.values())
    win = [i for i in leaders if res[i]["term"] == top_term]
    assert len(win) == 1, {i: (r["role"], r["leader"], r["term"]) for i, r in res.items()}
    w = win[0]
    # one leader per term: no two survivors claim leadership in the same term
    assert len({res[i]["term"] for i in leaders}) == len(leaders)
    for i, r in res.items():
        if r["rc"] != 0:
            # a survivor that answered after the winner's grace period was treated as failed (check_failure_count): it is
            # out of the group -- it keeps standing, nobody answers, it gives up -- and would have to join again
            assert f"REMOVE SERVER p{i}" in res[w]["log"] and blks[w]["disconnected"] & (1 << i), r["log"][-800:]
            continue
        if i != w and r["role"] == "follower":
            assert r["leader"] == w and r["term"] == top_term, (i, r["role"], r["leader"], r["term"], w, top_term)
            assert blks[i]["adjusted_by"] == w and blks[i]["role_leader"] == w and blks[i]["role_term"] == top_term
            assert blks[i]["last_idx"] == blks[w]["last_idx"]          # the follower holds the winner's log now
    followers = [i for i, r in res.items() if r["role"] == "follower" and r["leader"] == w]
    assert 1 + len(followers) >= n // 2 + 1                         # the winner leads a majority
    # up-to-date rule: the winner's log is not behind the log of anybody who voted for it (= its followers)
    key = lambda i: (logs[i][1], logs[i][0])                         # noqa: E731   (term, idx)
    assert all(key(w) >= key(i) for i in followers), (w, logs)
    assert "] LEADER" in res[w]["log"]                               # the line benchmarks/run.sh:52 greps for
    return w, top_term


@pytest.mark.timeout(240)
def test_three_replicas_leader_dies(harness):
    for _ in range(6):
        logs = {1: (40, 1), 2: (40, 1)}
        res, blks = run_election(harness, 3, 0, logs)
        w, term = check(res, blks, logs, 3)
        assert term >= 2


@pytest.mark.timeout(300)
def test_five_replicas_behind_server_cannot_win(harness):
    for trial in range(6):
        logs = {0: (100, 1), 1: (100, 1), 3: (70, 1), 4: (100, 1)}
        res, blks = run_election(harness, 5, 2, logs)
        w, term = check(res, blks, logs, 5)
        assert w != 3, "a server whose log is behind three others won their votes"


@pytest.mark.timeout(300)
def test_older_term_log_loses_to_newer_term(harness):
    """(term, idx) order: an entry of a newer term beats a longer log of an older term (dare_server.c:1640-1652)."""
    for trial in range(4):
        logs = {1: (90, 2), 2: (120, 1), 3: (90, 2), 4: (90, 2)}
        res, blks = run_election(harness, 5, 0, logs)
        w, term = check(res, blks, logs, 5)
        assert w != 2


@pytest.mark.timeout(300)
def test_silent_server_is_disconnected_by_the_winner(harness):
    """Two of five are gone (the leader and one that never answers): the three others still form a majority; the winner
    treats the silent one as failed (check_failure_count, dare_server.c:1189-1228)."""
    for trial in range(3):
        logs = {1: (55, 1), 2: (55, 1), 3: (55, 1)}
        res, blks = run_election(harness, 5, 0, logs, absent=(4,))
        w, term = check(res, blks, logs, 5)
        assert blks[w]["disconnected"] & (1 << 4)
        assert "REMOVE SERVER p4" in res[w]["log"]


@pytest.mark.timeout(300)
def test_slow_winner_is_followed_not_fought(harness):
    """The winner needs 60 ms to take over -- longer than its voter's patience (hb_timeout 20 ms + a random election
    timeout): the voter stands in a higher term meanwhile.  The winner no longer answers vote requests (it leads), so that
    candidacy can never win; when it times out the voter must give way to the announced leader instead of raising its term
    for ever (the give-way rule in elect(), case (a))."""
    for trial in range(4):
        logs = {1: (40, 1), 2: (40, 1)}
        res, blks = run_election(harness, 3, 0, logs, env={"MOCK_TAKEOVER_DELAY_US": "60000"})
        leaders = [i for i, r in res.items() if r["role"] == "leader"]
        assert len(leaders) == 1, {i: (r["role"], r["term"]) for i, r in res.items()}
        w = leaders[0]
        for i, r in res.items():
            assert r["rc"] == 0, r["log"][-600:]
            if i != w:
                assert r["role"] == "follower" and r["leader"] == w and r["term"] == res[w]["term"]
                assert blks[i]["role_leader"] == w and blks[i]["adjusted_by"] == w
                assert "Start election" in r["log"], "the scenario needs the voter to have stood meanwhile"


def test_false_positive_keeps_following(harness):
    """The failure detector fired but the leader's beat still moves (its kernel paused around a change of the peer set, or
    this replica's context was not scheduled): no election, no term bump -- the replica relaunches its kernel and keeps
    following (the reference's "false possitive" branch, dare_server.c:781-796)."""
    logs = {1: (40, 1), 2: (40, 1)}
    res, blks = run_election(harness, 3, 0, logs, env={"MOCK_LEADER_ALIVE": "1"})
    for i, r in res.items():
        assert (r["rc"], r["role"], r["leader"], r["term"]) == (0, "follower", 0, 1), r["log"][-400:]
        assert "false positive => p0 is alive" in r["log"] and "Start election" not in r["log"]
        assert blks[i]["launches"] == 1 and blks[i]["sid"] == (1 << 9) | (1 << 8) | 0     # kernel relaunched, SID untouched
    # a beat of ANOTHER term is not the leader I follow: the election goes ahead
    res, blks = run_election(harness, 3, 0, logs, env={"MOCK_LEADER_ALIVE": "7"})
    w, term = check(res, blks, logs, 3)
    assert term >= 2
