"""Drop-in test of the upper boundary: the reference's UNMODIFIED src/proxy/proxy.c
(oracle/_ref/libref_proxy.so, built from /root/reference by oracle/Makefile refproxy) runs on
top of libapus_dare.so + libapus_gpu.so, one process per replica as APUS deploys it
(benchmarks/run.sh:26), peers mapped with CUDA IPC.  The leader's "application" threads call
proxy_on_accept/read/close and block until commit; followers replay every committed request
into their local application (a TCP sink here) in log order."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFPROXY = os.path.join(ROOT, "oracle", "_ref", "libref_proxy.so")


@pytest.mark.parametrize("n,nconn,nreq,plen", [(3, 2, 400, 64), (3, 3, 150, 700)])
def test_reference_proxy_on_gpu_engine(n, nconn, nreq, plen):
    import __graft_entry__ as g
    g.build()
    if not os.path.exists(REFPROXY):
        pytest.skip("oracle/_ref/libref_proxy.so absent (built only where /root/reference exists)")
    with tempfile.TemporaryDirectory() as d:
        env = dict(os.environ, apus_rendezvous=os.path.join(d, "rdv"), apus_log_size=str(1 << 22))
        procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "proxy_worker.py"), str(i), str(n), str(nconn),
                                   str(nreq), str(plen), d], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
                 for i in range(n)]
        outs = [p.communicate(timeout=240)[0].decode(errors="replace") for p in procs]
        res = []
        for i in range(n):
            path = os.path.join(d, f"result{i}.json")
            assert os.path.exists(path), f"replica {i} produced no result:\n{outs[i][-2000:]}"
            res.append(json.load(open(path)))
        logs = [open(os.path.join(d, f"dare{i}.log")).read() for i in range(n)]
    # leader: every request (CONNECTs + SENDs + CLOSEs) was committed before its hook returned
    assert res[0]["highest_rec"] == 2 * nconn + nreq
    assert "] LEADER" in logs[0]                       # what benchmarks/run.sh:52 greps for
    # store_cmd once per appended entry; record sizes of SURVEY.md H4: 4 B (CONNECT/CLOSE), 24 B (SEND)
    for r in res:
        assert r["db_records"] == 2 * nconn + nreq, r
        assert r["db_sizes"] == {"4": 2 * nconn, "24": nreq}, r
    # followers replayed the byte streams, per connection, in order
    expect = []
    for c in range(nconn):
        buf = b"".join(bytes(((i * 31 + k) & 0xFF) for k in range(plen)) for i in range(nreq) if i % nconn == c)
        expect.append(hashlib.sha256(buf).hexdigest())
    for r in res[1:]:
        assert r["conns"] == nconn and r["bytes"] == nreq * plen, r
        assert sorted(r["sha"]) == sorted(expect)
    print("closed-loop latency through the unmodified proxy.c: p50 %.1f us, p99 %.1f us" % (res[0]["p50_us"], res[0]["p99_us"]))
