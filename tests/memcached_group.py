"""Shared driver of the memcached drop-in scenario (BASELINE config 4 in miniature): an UNMODIFIED memcached 1.4.21
(oracle/build_memcached.sh), four worker threads, is started per replica the way benchmarks/run.sh:26 starts an application
-- server_type / server_idx / group_size / config_path / dare_log_file, LD_PRELOAD of the reference's unmodified interposer --
on the GPU engine (interpose.so) or on the reference's own stack (interpose_ref.so, shim NIC).  Clients talk to the leader
only: 16 connections, each with its own key range, 1 KB values, sets and gets half and half (what the reference's memslap
run does, apps/memcached/run).  Followers are fed through the replicated log and must end up holding every key with the
value the leader holds.  Test infrastructure."""
import os
import signal
import socket
import subprocess
import tempfile
import threading
import time

import redis_group as RG

MEMCACHED = os.path.join(RG.REF, "memcached")


class Conn:
    def __init__(self, port):
        self.s = socket.create_connection(("127.0.0.1", port))
        self.s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        self.s.settimeout(30)
        self.buf = b""

    def line(self):
        while b"\r\n" not in self.buf:
            d = self.s.recv(65536)
            assert d, "memcached closed the connection"
            self.buf += d
        l, self.buf = self.buf.split(b"\r\n", 1)
        return l

    def take(self, n):
        while len(self.buf) < n:
            d = self.s.recv(65536)
            assert d, "memcached closed the connection"
            self.buf += d
        out, self.buf = self.buf[:n], self.buf[n:]
        return out

    def set(self, key, value):
        self.s.sendall(b"set %s 0 0 %d\r\n%s\r\n" % (key, len(value), value))
        return self.line()

    def get(self, key):
        self.s.sendall(b"get %s\r\n" % key)
        l = self.line()
        if l == b"END":
            return None
        n = int(l.split()[3])
        v = self.take(n + 2)[:n]
        assert self.line() == b"END"
        return v

    def close(self):
        self.s.close()


def value_of(c, k, vlen):
    seed = (c * 1000003 + k * 7919) & 0xFFFFFFFF
    return bytes(((seed >> (8 * (j & 3))) + j * 31) & 0xFF for j in range(vlen)).replace(b"\r", b"r").replace(b"\n", b"n")


def run_memcached_group(n, ndev, nconn=16, nkeys=200, vlen=1024, stack="gpu", base_port=21300, startup_timeout=120):
    procs, d = [None] * n, tempfile.mkdtemp(prefix="apus-memcached-")
    try:
        for i in range(n):
            wd = os.path.join(d, f"node{i}")
            os.makedirs(wd)
            with open(os.path.join(wd, "node.cfg"), "w") as f:
                f.write(RG.CFG.format(i=i, port=base_port + i))
            env = dict(os.environ, server_type="start", server_idx=str(i), group_size=str(n),
                       config_path=os.path.join(wd, "node.cfg"), dare_log_file=os.path.join(wd, "dare.log"))
            if stack == "gpu":
                env.update(LD_PRELOAD=RG.INTERPOSE, apus_rendezvous=os.path.join(d, "rdv"), apus_log_size=str(1 << 26),
                           apus_segv_trace="1")
            else:
                env.update(LD_PRELOAD=RG.INTERPOSE_REF, APUS_SHIM_DIR=os.path.join(d, "shim"))
                env.pop("mgid", None)
            procs[i] = subprocess.Popen([MEMCACHED, "-u", "root", "-p", str(base_port + i), "-U", "0", "-t", "4", "-l", "127.0.0.1",
                                         "-m", "256"], cwd=wd, env=env, stdout=open(os.path.join(wd, "app.out"), "w"),
                                        stderr=subprocess.STDOUT)

        def log(i):
            p = os.path.join(d, f"node{i}", "dare.log")
            return open(p, errors="replace").read() if os.path.exists(p) else ""

        def state():
            return "\n".join(f"--- replica {i}: rc={procs[i].poll()}\n{log(i)[-600:]}\n"
                             f"{open(os.path.join(d, f'node{i}', 'app.out'), errors='replace').read()[-1500:]}" for i in range(n))

        def leader_idx():
            who = [i for i in range(n) if "] LEADER" in log(i)]
            return who[-1] if who else None

        def up():
            assert all(p.poll() is None for p in procs), "a memcached died during start-up:\n" + state()
            if stack == "gpu":
                return "] LEADER" in log(0) and all(" up on GPU " in log(i) for i in range(n))
            return leader_idx() is not None

        RG.wait_for(up, startup_timeout, lambda: "the replicas to come up:\n" + state())
        if stack != "gpu":
            time.sleep(1.0)                                  # followers grant log access to the elected leader
        lead = leader_idx()
        lp = base_port + lead
        errors, ops = [], [0] * nconn

        def client(c):
            try:
                k = Conn(lp)
                for i in range(nkeys):                       # set / get half and half, this connection's own keys
                    key = b"c%d:k%d" % (c, i)
                    assert k.set(key, value_of(c, i, vlen)) == b"STORED"
                    assert k.get(key) == value_of(c, i, vlen)
                    ops[c] += 2
                k.close()
            except Exception as e:                           # noqa: BLE001 - surfaced below
                errors.append(f"connection {c}: {type(e).__name__}: {e}")

        t0 = time.time()
        th = [threading.Thread(target=client, args=(c,)) for c in range(nconn)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.time() - t0
        assert not errors, "\n".join(errors[:4]) + "\n" + state()
        # followers: every key, the leader's value -- replayed from the replicated log, never sent by a client
        followers = [i for i in range(n) if i != lead]
        for i in followers:
            k = Conn(base_port + i)
            RG.wait_for(lambda: k.get(b"c%d:k%d" % (nconn - 1, nkeys - 1)) is not None, 60,
                        lambda: f"follower {i} to catch up\n" + state())
            for c in range(nconn):
                for j in range(nkeys):
                    got = k.get(b"c%d:k%d" % (c, j))
                    if got is None:                          # the last requests of other connections may still be in flight
                        got = RG.wait_for(lambda: k.get(b"c%d:k%d" % (c, j)), 30, lambda: f"follower {i} key c{c}:k{j}\n" + state())
                    assert got == value_of(c, j, vlen), f"follower {i}: c{c}:k{j} differs"
            k.close()
        where = f"the GPU log, {n} replicas on {min(n, ndev)} GPU(s)" if stack == "gpu" else \
            f"the reference's own stack on the shim NIC, {n} replica processes"
        summary = (f"memcached set/get 50/50, {vlen} B values, {nconn} connections through {where} (leader p{lead}): "
                   f"{sum(ops) / dt:.0f} ops/s ({sum(ops)} operations in {dt:.2f} s, a Python client)")
        for p in procs:
            p.send_signal(signal.SIGINT)
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


def main():
    """Load generator for benchmarks/run_gpu.sh --app=memcached: python tests/memcached_group.py <port> <connections> <keys per connection> <value bytes>"""
    import sys
    port, nconn, nkeys, vlen = (int(x) for x in sys.argv[1:5])
    ops, errors = [0] * nconn, []

    def client(c):
        try:
            k = Conn(port)
            for i in range(nkeys):
                key = b"c%d:k%d" % (c, i)
                assert k.set(key, value_of(c, i, vlen)) == b"STORED"
                assert k.get(key) == value_of(c, i, vlen)
                ops[c] += 2
            k.close()
        except Exception as e:                               # noqa: BLE001
            errors.append(f"connection {c}: {type(e).__name__}: {e}")

    t0 = time.time()
    th = [threading.Thread(target=client, args=(c,)) for c in range(nconn)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.time() - t0
    print(f"memcached set/get 50/50, {vlen} B values, {nconn} connections: {sum(ops) / dt:.0f} ops/s ({sum(ops)} operations in {dt:.2f} s)"
          + ("; ERRORS: " + "; ".join(errors[:3]) if errors else ""))
    sys.exit(1 if errors else 0)


if __name__ == "__main__":
    main()
