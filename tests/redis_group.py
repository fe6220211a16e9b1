"""Shared driver of the redis drop-in scenario (tests/test_gpu_redis_dropin.py: the GPU engine under the reference's
interposer; tests/test_refstack_redis.py: the reference's own stack under the same interposer, on the shim NIC).
An UNMODIFIED redis-server 2.8.17 is started per replica exactly as benchmarks/run.sh:26 starts it --
    server_type=start server_idx=i group_size=N config_path=<libconfig file> dare_log_file=<log>
    LD_PRELOAD=interpose.so redis-server --port <p_i>
redis-benchmark / redis-cli talk to the leader's port only; followers are fed through the replicated log.
Test infrastructure."""
import os
import signal
import subprocess
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "oracle", "_ref")
SERVER, BENCH, CLI, INTERPOSE, INTERPOSE_REF = (os.path.join(REF, x) for x in (
    "redis-server", "redis-benchmark", "redis-cli", "interpose.so", "interpose_ref.so"))
BASE_PORT = 18880

CFG = """db_name = "node_test{i}";
req_log = 0;
ip_address = "127.0.0.1";
port = {port};
dare_global_config = {{
    hb_period = 0.01;
    elec_timeout_low = 100000;
    elec_timeout_high = 300000;
    retransmit_period = 0.04;
    rc_info_period = 0.05;
    log_pruning_period = 0.05;
}};
"""


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


def run_redis_group(n, ndev, nbench, nlist, stagger=0.0, order=None, startup_timeout=120, stack="gpu", base_port=BASE_PORT):
    """stack = "gpu": interpose.so on libapus_dare/libapus_gpu, replica 0 leads; "refstack": interpose_ref.so, the
    reference's own election decides.  Returns a one-line summary."""
    BASE_PORT = base_port
    procs, d = [None] * n, tempfile.mkdtemp(prefix="apus-redis-")
    try:
        for i in (order or range(n)):
            wd = os.path.join(d, f"node{i}")
            os.makedirs(wd)
            with open(os.path.join(wd, "node.cfg"), "w") as f:      # target/nodes.local.cfg, one file per replica
                f.write(CFG.format(i=i, port=BASE_PORT + i))
            env = dict(os.environ, server_type="start", server_idx=str(i), group_size=str(n),
                       config_path=os.path.join(wd, "node.cfg"), dare_log_file=os.path.join(wd, "dare.log"))
            if stack == "gpu":
                env.update(LD_PRELOAD=INTERPOSE, apus_rendezvous=os.path.join(d, "rdv"), apus_log_size=str(1 << 24),
                           apus_segv_trace="1")
            else:
                env.update(LD_PRELOAD=INTERPOSE_REF, APUS_SHIM_DIR=os.path.join(d, "shim"))
                env.pop("mgid", None)
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

        def leader_idx():
            who = [i for i in range(n) if "] LEADER" in log(i)]
            return who[-1] if who else None

        def up():
            assert all(p.poll() is None for p in procs), "a redis-server died during start-up:\n" + state()
            if stack == "gpu":
                return "] LEADER" in log(0) and all(" up on GPU " in log(i) for i in range(n))
            return leader_idx() is not None

        wait_for(up, startup_timeout, lambda: "the replicas to come up:\n" + state())
        lead = leader_idx()
        if stack != "gpu":
            time.sleep(1.0)                                  # followers grant log access to the elected leader
            lead = leader_idx()
        lp = BASE_PORT + lead
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
        followers = [i for i in range(n) if i != lead]
        for i in followers:
            port = BASE_PORT + i
            wait_for(lambda: cli(port, "DEBUG", "DIGEST") == want_digest, 120,
                     lambda: f"follower {i} to converge (leader DBSIZE {want_size}, follower {cli(port, 'DBSIZE')})\n"
                     + state())
            assert cli(port, "DBSIZE") == want_size
            assert cli(port, "LRANGE", "mylist", "0", "-1") == want_list
            assert cli(port, "GET", "ctr") == str(nlist)
        where = f"the GPU log, {n} replicas on {min(n, ndev)} GPU(s)" if stack == "gpu" else \
            f"the reference's own stack on the shim NIC, {n} replica processes"
        summary = (f"redis-benchmark through {where} (leader p{lead}): "
                   f"{out.strip().splitlines()[-1]} ({nbench} SETs in {dt:.1f} s)")
        # 3. the reference's shutdown drill (kill -2, benchmarks/run.sh:78): every process must be gone afterwards.
        # Who handles SIGINT is a race the reference has too -- dare_server_init installs int_handler
        # (dare_server.c:186-187) while redis-server's main installs its own SIGINT/SIGTERM handler -- and on the
        # leader redis's shutdown closes its LISTENING socket through the close() hook, which proxy.c:141-146
        # dereferences as an unknown connection (a reference bug: the leader dies by SIGSEGV instead of exiting).
        for p in procs:
            p.send_signal(signal.SIGINT)
        for i in followers:
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
