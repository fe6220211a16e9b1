"""Join / reconfiguration into an emptied slot and the snapshot callbacks (N3): a follower process is
killed, the leader's failure detector removes it with a CONFIG entry, a replacement started with server_type=join asks to
be let in; the leader snapshots the proxy's state (get_db_size / create_db_snapshot), sends its log over NVLink, appends
the CONFIG entry that puts the slot back; the joiner loads the snapshot (apply_db_snapshot) and follows the log.
Reference: dare_ibv_ud.c:952-1087, dare_server.c:598-721, 1883-1937, proxy.c:300-339."""
import hashlib
import json
import os
import signal
import subprocess
import sys
import tempfile
import time

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFPROXY = os.path.join(ROOT, "oracle", "_ref", "libref_proxy.so")


def wait_file(path, timeout, procs, what):
    t0 = time.time()
    while not os.path.exists(path):
        assert time.time() - t0 < timeout, f"timed out waiting for {what}"
        for p in procs:
            if not (p.poll() is None or p.returncode == 0 or p.returncode == -9):
                raise AssertionError(f"a replica died (rc {p.returncode}) while waiting for {what}:\n"
                                     + (p.communicate()[0] or b"").decode(errors="replace")[-3000:])
        time.sleep(0.01)


def test_follower_replaced_by_a_joiner():
    import __graft_entry__ as g
    g.build()
    if not os.path.exists(REFPROXY):
        pytest.skip("oracle/_ref/libref_proxy.so absent (built only where /root/reference exists)")
    import apus_b200
    nd = 1          # all replica processes on GPU 0: a dead process's region must stay mapped (tools/failover_drill.py explains)
    n, nconn, nreqA, nreqB, plen = 3, 2, 300, 400, 64
    with tempfile.TemporaryDirectory() as d:
        # fewer GPUs than processes: contexts are time-sliced, the failure detector needs a generous timeout
        to = "20000" if nd >= n else "1500000"
        env = dict(os.environ, apus_rendezvous=os.path.join(d, "rdv"), apus_log_size=str(1 << 22), APUS_NO_BUILD="1",
                   apus_hb_period_us="1000", apus_hb_timeout_us=to)

        def start(i, how):
            e = dict(env, apus_gpu=str(i % nd))
            return subprocess.Popen([sys.executable, os.path.join(HERE, "join_worker.py"), str(i), str(n), str(nconn), str(nreqA),
                                     str(nreqB), str(plen), d, how], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)

        procs = [start(i, "start") for i in range(n)]
        try:
            wait_file(os.path.join(d, "phaseA_done"), 120, procs, "phase A")
            os.kill(procs[2].pid, signal.SIGKILL)                                  # a follower dies
            t0 = time.time()
            while "REMOVE SERVER p2" not in open(os.path.join(d, "dare0.log")).read():
                assert time.time() - t0 < 60, "the leader never removed the dead follower:\n" + open(os.path.join(d, "dare0.log")).read()[-1500:]
                time.sleep(0.02)
            joiner = start(2, "join")                                              # its replacement
            procs.append(joiner)
            t0 = time.time()
            while "p2 joined" not in open(os.path.join(d, "dare0.log")).read():
                assert time.time() - t0 < 90, "join did not complete:\n" + open(os.path.join(d, "dare0.log")).read()[-1500:]
                assert joiner.poll() is None, joiner.communicate()[0].decode(errors="replace")[-2000:]
                time.sleep(0.02)
            open(os.path.join(d, "phaseB_go"), "w").close()
            res = {}
            for tag in ("0", "1", "2j"):
                wait_file(os.path.join(d, f"result{tag}.json"), 120, [procs[0], procs[1], joiner], f"result {tag}")
                res[tag] = json.load(open(os.path.join(d, f"result{tag}.json")))
            logs = {t: open(os.path.join(d, f"dare{t}.log")).read() for t in ("0", "1", "2j")}
        except AssertionError as ex:
            logs_txt = ""
            for f in sorted(os.listdir(d)):
                if f.startswith("dare") and f.endswith(".log"):
                    logs_txt += f"\n--- {f}\n" + open(os.path.join(d, f)).read()[-1800:]
            raise AssertionError(str(ex) + logs_txt) from None
        finally:
            for p in procs:
                if p.poll() is None:
                    p.kill()
    # the leader snapshotted through the proxy callback, the joiner loaded it
    assert res["0"]["db_dumps"] == 1
    assert "snapshot of" in logs["2j"] and "joined: leader p0" in logs["2j"]
    # every replica holds the same log: phase A, CONFIG removing p2, CONFIG adding it back, phase B
    ents = {t: [(e["idx"], e["term"], e["type"], e["sender"], e["sha"]) for e in r["entries"]] for t, r in res.items()}
    assert ents["1"] == ents["0"] and ents["2j"] == ents["0"]
    cfgs = [e for e in res["0"]["entries"] if e["type"] == 2]
    masks = [int.from_bytes(bytes.fromhex(e["data"])[12:16], "little") for e in cfgs]
    assert masks == [0b111, 0b011, 0b111], masks
    assert res["0"]["highest_rec"] == 2 * (2 * nconn) + nreqA + nreqB
    # the joiner replayed phase B completely and in order, like the follower that never left
    expect = sorted(hashlib.sha256(b"".join(bytes((((100000 + i) * 31 + k) & 0xFF) for k in range(plen)) for i in range(nreqB) if i % nconn == c)).hexdigest()
                    for c in range(nconn))
    for t in ("1", "2j"):
        shas = [x["sha"] for x in res[t]["replay"]]
        assert all(s in shas for s in expect), f"replica {t} did not replay phase B"
    for r in res.values():
        assert r["offsets"]["end"] == res["0"]["offsets"]["end"] and r["offsets"]["commit"] == res["0"]["offsets"]["end"]
