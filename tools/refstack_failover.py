#!/usr/bin/env python
"""Leader-failover drill on the REFERENCE's own stack (BASELINE.json configs[4], benchmarks/reconf_bench.sh analogue):
N replica processes of oracle/_ref/libref_stack.so on the verbs shim, the elected leader under load; the leader
process is killed and the time to the next "] LEADER" line in a survivor's log is reported -- the reference-side
number the GPU engine's control plane has to be compared with.
Test infrastructure; no GPU involved.

    python tools/refstack_failover.py [replicas=5] [trials=5] [signal=KILL|INT]
"""
import json
import os
import re
import signal
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "refstack_worker.py")
STAMP = re.compile(r"\[(\d+):(\d+)\] \[T(\d+)\] LEADER")


def leaders(path):
    if not os.path.exists(path):
        return []
    return [(int(s) + int(us) * 1e-6, int(t)) for s, us, t in STAMP.findall(open(path, errors="replace").read())]


def trial(n, sig):
    d = tempfile.mkdtemp(prefix="apus-failover-")
    env = dict(os.environ, REFSTACK_NO_IMAGE="1", REFSTACK_RUN_TIMEOUT="60")
    procs = [subprocess.Popen([sys.executable, WORKER, str(i), str(n), "1", "2000000", "64", d, "1"], env=env,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(n)]
    try:
        t0 = time.time()
        while not os.path.exists(os.path.join(d, "leader.json")):
            assert time.time() - t0 < 60, "no leader"
            time.sleep(0.01)
        lead = json.load(open(os.path.join(d, "leader.json")))["idx"]
        time.sleep(1.5)                                         # the leader is committing requests now
        t_kill = time.time()
        procs[lead].send_signal(sig)
        new = None
        while time.time() - t_kill < 30 and new is None:
            for i in range(n):
                if i == lead:
                    continue
                for ts, term in leaders(os.path.join(d, f"node{i}", "dare.log")):
                    if ts > t_kill - 0.001:
                        new = (i, ts - t_kill, term)
            time.sleep(0.002)
        return lead, new
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        for p in procs:
            p.wait()
        subprocess.run(["rm", "-rf", d])


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    trials = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    sig = signal.SIGINT if (len(sys.argv) > 3 and sys.argv[3] == "INT") else signal.SIGKILL
    print(f"# reference stack on the verbs shim, {n} replicas, leader killed with {sig.name} under a closed-loop 64 B load; "
          f"hb_period 10 ms, election timeout 100-300 ms (target/nodes.local.cfg)")
    out = []
    for k in range(trials):
        lead, new = trial(n, sig)
        if new is None:
            print(f"trial {k}: leader p{lead} killed, no new leader within 30 s")
        else:
            print(f"trial {k}: leader p{lead} killed -> p{new[0]} is LEADER of term {new[2]} after {1e3 * new[1]:.1f} ms", flush=True)
            out.append(new[1])
    if out:
        out.sort()
        print(f"# recovery (kill -> next LEADER line): min {1e3 * out[0]:.0f} ms, median {1e3 * out[len(out) // 2]:.0f} ms, max {1e3 * out[-1]:.0f} ms")
