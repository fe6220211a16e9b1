"""One replica PROCESS of the reference's own software stack on the verbs shim (oracle/_ref/libref_stack.so =
the reference's unmodified src/dare/*.c + proxy.c + db-interface.c + config-*.c, real libev / libconfig /
BerkeleyDB; oracle/verbs_shim is the NIC).  Test infrastructure: it produces the reference's log images for
tests/test_oracle_vs_refstack.py and is the CPU reference arm of bench.py.

    refstack_worker.py <idx> <n> <nconn> <nreq> <plen> <outdir> [threads]

Every replica starts the same way (server_type=start, benchmarks/run.sh:26).  Whoever wins the election drives
the workload exactly as src/spec_hooks.cpp would: proxy_on_accept per connection, proxy_on_read per request
(returns once the request is committed), proxy_on_close -- refstack_drive() in oracle/ref_stack_access.c.  With
`threads` > 1 the connections are spread over that many application threads (memcached-style); request i carries
payload bytes (i*31+k)&0xFF; plen < 0 asks for ragged lengths up to -plen.  REFSTACK_PRUNE=<seconds> sets
log_pruning_period (default 0.05 as in target/nodes.local.cfg; a large value keeps HEAD entries out of the log).
At the end every replica dumps entries[0..end) of its log and its offsets."""
import ctypes as C
import hashlib
import json
import os
import socket
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STACK = os.path.join(ROOT, "oracle", "_ref", os.environ.get("REFSTACK_LIB", "libref_stack.so"))

CFG = """db_name = "node_test{idx}";
req_log = 0;
ip_address = "127.0.0.1";
port = {port};
dare_global_config = {{
    hb_period = 0.01;
    elec_timeout_low = 100000;
    elec_timeout_high = 300000;
    retransmit_period = 0.04;
    rc_info_period = 0.05;
    log_pruning_period = {prune};
}};
"""


def main():
    idx, n, nconn, nreq, plen, outdir = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]),
                                          int(sys.argv[5]), sys.argv[6])
    nthreads = int(sys.argv[7]) if len(sys.argv) > 7 else 1
    received = {}
    lock = threading.Lock()

    def sink(port_holder):                      # the follower's "application": a TCP sink that records the replay
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind(("127.0.0.1", 0))
        srv.listen(256)
        port_holder.append(srv.getsockname()[1])

        def serve(conn, k):
            h, nbytes = hashlib.sha256(), 0
            while True:
                d = conn.recv(1 << 16)
                if not d:
                    break
                h.update(d)
                nbytes += len(d)
                with lock:
                    received[k] = (nbytes, h.copy())
        k = 0
        while True:
            conn, _ = srv.accept()
            with lock:
                received[k] = (0, hashlib.sha256())
            threading.Thread(target=serve, args=(conn, k), daemon=True).start()
            k += 1

    ph = []
    threading.Thread(target=sink, args=(ph,), daemon=True).start()
    while not ph:
        time.sleep(0.01)
    wd = os.path.join(outdir, f"node{idx}")
    os.makedirs(wd, exist_ok=True)
    os.chdir(wd)
    with open("node.cfg", "w") as f:
        f.write(CFG.format(idx=idx, port=ph[0], prune=float(os.environ.get("REFSTACK_PRUNE", "0.05"))))
    os.environ.update(server_idx=str(idx), group_size=str(n), server_type="start",
                      dare_log_file=os.path.join(wd, "dare.log"), APUS_SHIM_DIR=os.path.join(outdir, "shim"))
    os.environ.pop("mgid", None)

    st = C.CDLL(STACK, mode=C.RTLD_GLOBAL)
    st.proxy_init.restype = C.c_void_p
    st.proxy_init.argtypes = [C.c_char_p, C.c_char_p]
    st.proxy_on_read.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_int]
    st.proxy_on_accept.argtypes = [C.c_void_p, C.c_int]
    st.proxy_on_close.argtypes = [C.c_void_p, C.c_int]
    st.refstack_offsets.argtypes = [C.POINTER(C.c_uint64)]
    st.refstack_log_read.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]

    proxy = st.proxy_init(b"node.cfg", None)              # spec_hooks.cpp:33
    assert proxy, "proxy_init failed"

    def offsets():
        o = (C.c_uint64 * 8)()
        if st.refstack_offsets(o):
            return None
        return dict(zip(("head", "apply", "commit", "end", "tail", "len", "sid", "term"), [int(x) for x in o]))

    leader_file = os.path.join(outdir, "leader.json")
    done_file = os.path.join(outdir, "done.json")
    deadline = time.time() + float(os.environ.get("REFSTACK_ELECTION_TIMEOUT", "60"))
    i_lead = False
    while time.time() < deadline and not os.path.exists(leader_file):
        if st.refstack_is_leader():
            i_lead = True
            break
        time.sleep(0.005)
    result = {"idx": idx, "leader": i_lead}
    if not i_lead and not os.path.exists(leader_file):
        print("no leader was elected within the timeout", file=sys.stderr)
        os._exit(3)
    if i_lead:
        with open(leader_file + ".tmp", "w") as f:
            json.dump({"idx": idx}, f)
        os.rename(leader_file + ".tmp", leader_file)
        time.sleep(float(os.environ.get("REFSTACK_SETTLE", "0.5")))     # let the followers grant log access
        lat = (C.c_uint64 * max(nreq, 1))()
        secs = C.c_double(0.0)
        st.refstack_drive.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_double)]
        steps = []
        for _ in range(int(os.environ.get("REFSTACK_STEPS", "1"))):       # bench.py: warm-up + timed steps, same group
            rc = st.refstack_drive(proxy, nthreads, nconn, nreq, plen, lat, C.byref(secs))   # spec_hooks.cpp:116-174
            assert rc == 0
            dt = secs.value
            allat = sorted(int(x) for x in lat[:nreq])
            steps.append(dict(seconds=dt, requests=nreq, threads=nthreads, ops_per_s=(nreq + 2 * nconn) / dt if dt > 0 else 0.0,
                              p50_us=allat[len(allat) // 2] / 1e3 if allat else 0.0,
                              p99_us=allat[int(len(allat) * 0.99)] / 1e3 if allat else 0.0,
                              pct_us={str(q): allat[min(len(allat) - 1, int(len(allat) * q / 100))] / 1e3
                                      for q in (10, 25, 50, 75, 90, 95, 99, 99.9)} if allat else {},
                              max_us=allat[-1] / 1e3 if allat else 0.0))
        result.update(steps[-1])
        result["steps"] = steps
        # quiescence: the prune timer may still append one HEAD entry (never two in a row, dare_log.h:472-478)
        quiet = max(0.25, 5 * float(os.environ.get("REFSTACK_PRUNE", "0.05"))) if float(os.environ.get("REFSTACK_PRUNE", "0.05")) < 10 else 0.25
        last, since = None, time.time()
        while time.time() - since < quiet:
            o = offsets()
            cur = (o["end"], o["commit"], o["apply"])
            if cur != last or not (o["end"] == o["commit"] == o["apply"]):
                last, since = cur, time.time()
            time.sleep(0.005)
        with open(done_file + ".tmp", "w") as f:
            json.dump({"end": o["end"], "term": o["term"]}, f)
        os.rename(done_file + ".tmp", done_file)
    else:
        t0 = time.time()
        while not os.path.exists(done_file) and time.time() - t0 < float(os.environ.get("REFSTACK_RUN_TIMEOUT", "300")):
            time.sleep(0.02)
    # wait until this replica holds (and, for followers, has applied) everything the leader committed
    want = json.load(open(done_file)) if os.path.exists(done_file) else None
    t0 = time.time()
    while want and time.time() - t0 < 20:
        o = offsets()
        if o and o["end"] == want["end"] and (i_lead or o["apply"] == want["end"]):
            break
        time.sleep(0.01)
    time.sleep(0.2)
    o = offsets() or {}
    result["offsets"] = o
    if hasattr(st, "shim_transport_stats"):                  # how this replica's RDMA operations travelled (verbs_shim.c)
        ts = (C.c_uint64 * 2)()
        st.shim_transport_stats(ts)
        result["rdma_ops"] = {"memcpy": int(ts[0]), "process_vm": int(ts[1])}
    if o and not os.environ.get("REFSTACK_NO_IMAGE"):
        end = o["end"] if o["end"] != o["len"] else 0
        img = C.create_string_buffer(max(end, 1))
        if end:
            st.refstack_log_read(0, end, img)
        with open(os.path.join(outdir, f"image{idx}.bin"), "wb") as f:
            f.write(img.raw[:end])
    st.refstack_highest_rec.restype = C.c_uint64
    st.refstack_highest_rec.argtypes = [C.c_void_p]
    st.refstack_cur_rec.restype = C.c_uint64
    st.refstack_cur_rec.argtypes = [C.c_void_p]
    st.refstack_records_len.restype = C.c_uint32
    result["proxy"] = {"highest_rec": int(st.refstack_highest_rec(proxy)), "cur_rec": int(st.refstack_cur_rec(proxy)),
                       "records_len": int(st.refstack_records_len())}
    with lock:
        result["replay"] = {"conns": len(received), "bytes": sum(v[0] for v in received.values()),
                            "sha": sorted(v[1].hexdigest() for v in received.values())}
    # what the reference logged UP TO HERE (tear-down follows: replicas exit at slightly different times and the last ones
    # log the others as failed)
    try:
        text = open(os.path.join(wd, "dare.log"), errors="replace").read()
    except OSError:
        text = ""
    result["log_marks"] = {"removals": text.count("REMOVE SERVER"), "leaderships": text.count("] LEADER")}
    with open(os.path.join(outdir, f"result{idx}.json.tmp"), "w") as f:
        json.dump(result, f)
    os.rename(os.path.join(outdir, f"result{idx}.json.tmp"), os.path.join(outdir, f"result{idx}.json"))
    t0 = time.time()
    while time.time() - t0 < 30 and not all(os.path.exists(os.path.join(outdir, f"result{i}.json")) for i in range(n)):
        time.sleep(0.02)
    os._exit(0)


if __name__ == "__main__":
    main()
