"""Drop-in test of the whole operator surface (BASELINE config 3 in miniature):
an UNMODIFIED redis-server 2.8.17 is started per replica exactly as benchmarks/run.sh:26 starts it --
    server_type=start server_idx=i group_size=N config_path=<libconfig file> dare_log_file=<log>
    LD_PRELOAD=interpose.so redis-server --port <p_i>
-- where interpose.so is the reference's unmodified spec_hooks.cpp + proxy.c + db-interface.c +
config-proxy.c (real libconfig 1.4.9 and BerkeleyDB 5.1.29 from the vendored tarballs) linked against
libapus_dare.so + libapus_gpu.so instead of libdare.a/-lev/-libverbs (oracle/build_refapp.sh).
redis-benchmark / redis-cli talk to the leader's port only; every read() of the leader is committed through
the GPU log before Redis sees it, followers replay the committed byte streams into their own Redis.
The three data sets must then be identical (DEBUG DIGEST = Redis's own SHA-1 of the keyspace)."""
import os
import signal
import subprocess
import tempfile
import time

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(420)]
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "oracle", "_ref")
SERVER, BENCH, CLI, INTERPOSE = (os.path.join(REF, x) for x in ("redis-server", "redis-benchmark", "redis-cli",
                                                                   "interpose.so"))
BASE_PORT = 18880


def cli(port, *args, stdin=None, timeout=120):
    out = subprocess.run([CLI, "-p", str(port), *args], input=stdin, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         timeout=timeout)
    return out.stdout.decode(errors="replace").strip()


def wait_for(pred, timeout, what):
    t0 = time.time()
    while time.time() - t0 < timeout:
        v = pred()
        if v:
            return v
        time.sleep(0.2)
    raise AssertionError(f"timed out after {timeout}s waiting for {what() if callable(what) else what}")


def test_redis_replicated_through_gpu_log():
    import __graft_entry__ as g
    g.build()
    for f in (SERVER, BENCH, CLI, INTERPOSE):
        if not os.path.exists(f):
            pytest.skip(f"{f} absent (built only where /root/reference exists: oracle/build_refapp.sh)")
    import apus_b200
    ndev = apus_b200.lib().apus_device_count()
    n = 3
    # replicas that share one GPU are time-sliced by the driver (milliseconds per commit): keep the load small there
    nbench, nlist = (20000, 2000) if ndev >= n else (1200, 150)
    print(run_redis_group(n, ndev, nbench, nlist))


def run_redis_group(n, ndev, nbench, nlist, stagger=0.0, order=None, startup_timeout=120):
    procs, d = [None] * n, tempfile.mkdtemp(prefix="apus-redis-")
    try:
        for i in (order or range(n)):
            wd = os.path.join(d, f"node{i}")
            os.makedirs(wd)
            with open(os.path.join(wd, "node.cfg"), "w") as f:      # target/nodes.local.cfg, one file per replica
                f.write(f'db_name = "node_test{i}";\nreq_log = 0;\nip_address = "127.0.0.1";\nport = {BASE_PORT + i};\n')
            env = dict(os.environ, server_type="start", server_idx=str(i), group_size=str(n),
                       config_path=os.path.join(wd, "node.cfg"), dare_log_file=os.path.join(wd, "dare.log"),
                       LD_PRELOAD=INTERPOSE, apus_rendezvous=os.path.join(d, "rdv"), apus_log_size=str(1 << 24),
                       apus_segv_trace="1")
            procs[i] = subprocess.Popen([SERVER, "--port", str(BASE_PORT + i), "--save", "", "--bind", "127.0.0.1"],
                                        cwd=wd, env=env, stdout=open(os.path.join(wd, "redis.out"), "w"),
                                        stderr=subprocess.STDOUT)
            if stagger:
                time.sleep(stagger)

        def log(i):
            p = os.path.join(d, f"node{i}", "dare.log")
            return open(p).read() if os.path.exists(p) else ""

        # benchmarks/run.sh:52 finds the leader by grepping for "] LEADER"
        def state():
            return "\n".join(f"--- replica {i}: rc={procs[i].poll()}\n{log(i)[-600:]}\n"
                             f"{open(os.path.join(d, f'node{i}', 'redis.out')).read()[-2500:]}" for i in range(n))

        def up():
            assert all(p.poll() is None for p in procs), "a redis-server died during start-up:\n" + state()
            return "] LEADER" in log(0) and all(" up on GPU " in log(i) for i in range(n))

        wait_for(up, startup_timeout, lambda: "the replicas to come up:\n" + state())
        lp = BASE_PORT
        # 1. the reference's config-3 load: redis-benchmark SET, 128 B values, 16 concurrent clients
        t0 = time.time()
        out = subprocess.run([BENCH, "-p", str(lp), "-t", "set", "-d", "128", "-c", "16", "-n", str(nbench), "-r", "100000",
                              "-q"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300).stdout.decode()
        dt = time.time() - t0
        assert "requests per second" in out, out
        # 2. an order-sensitive stream on ONE connection (the reference orders per connection, proxy.c:121-133)
        cmds = "".join(f"RPUSH mylist {i}\nINCR ctr\n" for i in range(nlist)).encode()
        cli(lp, stdin=cmds, timeout=300)
        want_list = cli(lp, "LRANGE", "mylist", "0", "-1")
        want_size = cli(lp, "DBSIZE")
        want_digest = cli(lp, "DEBUG", "DIGEST")
        assert want_list.split() == [str(i) for i in range(nlist)]
        assert cli(lp, "GET", "ctr") == str(nlist)
        assert len(want_digest) == 40 and want_digest != "0" * 40
        # followers: same keyspace, same list order, same digest -- replayed from the GPU log, never sent by a client
        for i in range(1, n):
            port = BASE_PORT + i
            wait_for(lambda: cli(port, "DEBUG", "DIGEST") == want_digest, 120,
                     lambda: f"follower {i} to converge (leader DBSIZE {want_size}, follower {cli(port, 'DBSIZE')})\n"
                     + state())
            assert cli(port, "DBSIZE") == want_size
            assert cli(port, "LRANGE", "mylist", "0", "-1") == want_list
            assert cli(port, "GET", "ctr") == str(nlist)
        summary = (f"redis-benchmark through the GPU log, {n} replicas on {min(n, ndev)} GPU(s): "
                   f"{out.strip().splitlines()[-1]} ({nbench} SETs in {dt:.1f} s)")
        # 3. the reference's shutdown drill (kill -2, benchmarks/run.sh:78): every process must be gone afterwards.
        # Who handles SIGINT is a race the reference has too -- dare_server_init installs int_handler
        # (dare_server.c:186-187) while redis-server's main installs its own SIGINT/SIGTERM handler -- and on the
        # leader redis's shutdown closes its LISTENING socket through the close() hook, which proxy.c:141-146
        # dereferences as an unknown connection (a reference bug: the leader dies by SIGSEGV instead of exiting).
        for p in procs:
            p.send_signal(signal.SIGINT)
        for i in range(1, n):
            wait_for(lambda: procs[i].poll() is not None or "SIGINT detected" in log(i), 60,
                     lambda: f"follower {i} to shut down on SIGINT\n" + state())
        return summary
    finally:
        for p in procs:
            if p is not None and p.poll() is None:
                p.kill()
        for p in procs:
            try:
                if p is not None:
                    p.wait(timeout=20)
            except subprocess.TimeoutExpired:
                pass
        subprocess.run(["rm", "-rf", d])
