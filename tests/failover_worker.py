"""One replica process of the leader-failover drill (BASELINE config 5; benchmarks/reconf_bench.sh:249-343 analogue).

    failover_worker.py <idx> <n> <nconn> <nreq2> <plen> <outdir>

Every process = libapus_gpu.so + libapus_dare.so + the reference's UNMODIFIED proxy.c (oracle/_ref/libref_proxy.so),
started the way benchmarks/run.sh:26 starts a replica.  Replica 0 leads first and issues requests in a closed loop until
the drill kills it (SIGKILL); whoever wins the election then issues <nreq2> more requests over <nconn> new connections.
Followers replay everything into a TCP sink.  At the end every survivor dumps the entries of its log.
"""
import ctypes as C
import hashlib
import json
import os
import socket
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    idx, n, nconn, nreq2, plen, outdir = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]),
                                           int(sys.argv[5]), sys.argv[6])
    received = {}
    lock = threading.Lock()

    def sink(port_holder):
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind(("127.0.0.1", 0))
        srv.listen(256)
        port_holder.append(srv.getsockname()[1])

        def serve(conn, k):
            h, nb = hashlib.sha256(), 0
            with lock:
                received[k] = [0, h, b""]
            while True:
                d = conn.recv(1 << 16)
                if not d:
                    break
                h.update(d)
                nb += len(d)
                with lock:
                    received[k][0] = nb
                    if len(received[k][2]) < 8:
                        received[k][2] = (received[k][2] + d)[:8]

        k = 0
        while True:
            conn, _ = srv.accept()
            threading.Thread(target=serve, args=(conn, k), daemon=True).start()
            k += 1

    ph = []
    threading.Thread(target=sink, args=(ph,), daemon=True).start()
    while not ph:
        time.sleep(0.01)
    os.environ["stub_port"] = str(ph[0])
    os.environ["server_idx"] = str(idx)
    os.environ["group_size"] = str(n)
    os.environ["server_type"] = "start"
    os.environ["dare_log_file"] = os.path.join(outdir, f"dare{idx}.log")

    gpu = C.CDLL(os.path.join(ROOT, "apus_b200", "libapus_gpu.so"), mode=C.RTLD_GLOBAL)
    dare = C.CDLL(os.path.join(ROOT, "apus_b200", "libapus_dare.so"), mode=C.RTLD_GLOBAL)
    px = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_proxy.so"), mode=C.RTLD_GLOBAL)
    px.proxy_init.restype = C.c_void_p
    px.proxy_init.argtypes = [C.c_char_p, C.c_char_p]
    px.proxy_on_read.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_int]
    px.proxy_on_accept.argtypes = [C.c_void_p, C.c_int]
    px.proxy_on_close.argtypes = [C.c_void_p, C.c_int]
    px.stub_highest_rec.restype = C.c_uint64
    px.stub_highest_rec.argtypes = [C.c_void_p]
    dare.apus_dare_replica.restype = C.c_void_p
    dare.apus_dare_term.restype = C.c_uint64

    os.chdir(outdir)
    proxy = px.proxy_init(b"nodes.local.cfg", None)
    assert proxy
    result = {"idx": idx}
    t_up = time.time()
    while not dare.apus_dare_replica():
        assert time.time() - t_up < 60
        time.sleep(0.01)

    def drive(first_fd, count, tag):
        """closed loop over nconn connections; count < 0: until killed.  Progress goes to a file the drill reads."""
        for c in range(nconn):
            px.proxy_on_accept(proxy, first_fd + c)
        i, t_last = 0, time.time()
        lat = []
        prog = os.path.join(outdir, f"progress_{tag}.txt")
        while count < 0 or i < count:
            c = i % nconn
            payload = bytes(((i * 31 + k) & 0xFF) for k in range(plen))
            buf = C.create_string_buffer(payload, plen)
            a = time.perf_counter_ns()
            px.proxy_on_read(proxy, buf, plen, first_fd + c)
            lat.append(time.perf_counter_ns() - a)
            i += 1
            if i % 64 == 0 or time.time() - t_last > 0.005:
                t_last = time.time()
                with open(prog + ".tmp", "w") as f:
                    f.write(f"{i} {time.time()}\n")
                os.rename(prog + ".tmp", prog)
        for c in range(nconn):
            px.proxy_on_close(proxy, first_fd + c)
        return lat

    if idx == 0:
        t0 = time.time()
        while not dare.is_leader():
            assert time.time() - t0 < 60, "leader never came up"
            time.sleep(0.005)
        with open(os.path.join(outdir, "phase1_started"), "w") as f:
            f.write(str(time.time()))
        drive(100, -1, "p1")                                  # until SIGKILL
        return
    # a survivor: wait for the drill to end, or become the leader
    led = False
    done_file = os.path.join(outdir, "done.json")
    while not os.path.exists(done_file):
        if dare.is_leader() and not led:
            led = True
            t_lead = time.time()
            with open(os.path.join(outdir, "new_leader.json.tmp"), "w") as f:
                json.dump({"idx": idx, "t_leader": t_lead, "term": int(dare.apus_dare_term())}, f)
            os.rename(os.path.join(outdir, "new_leader.json.tmp"), os.path.join(outdir, "new_leader.json"))
            lat = drive(300, nreq2, "p2")
            lat.sort()
            result.update(leader=True, t_leader=t_lead, t_first_commit=t_lead, highest_rec=int(px.stub_highest_rec(proxy)),
                          p50_us=lat[len(lat) // 2] / 1e3, p99_us=lat[int(len(lat) * 0.99)] / 1e3)
            time.sleep(0.3)
            with open(done_file + ".tmp", "w") as f:
                json.dump({"leader": idx}, f)
            os.rename(done_file + ".tmp", done_file)
        time.sleep(0.002)
    time.sleep(0.5)                                           # let the replay drain
    # dump the entries of my log (through the C ABI, the kernel still running)
    rep = C.c_void_p(dare.apus_dare_replica())
    gpu.apus_log_offsets.argtypes = [C.c_void_p, C.c_void_p]
    gpu.apus_log_read.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    offs = (C.c_uint64 * 8)()
    assert gpu.apus_log_offsets(rep, offs) == 0
    head, apply, commit, end = int(offs[0]), int(offs[1]), int(offs[2]), int(offs[3])
    # the follower's header `end` is what its kernel saw last; read generously up to the leader-announced commit
    L = int(offs[7])
    stop = end if end != L else 0
    img = (C.c_uint8 * max(stop, 1))()
    if stop:
        assert gpu.apus_log_read(rep, 0, stop, img) == 0
    raw = bytes(img[:stop])
    ents, off = [], 0
    while off + 64 <= stop:
        typ = raw[off + 26]
        ln = raw[off + 48] | (raw[off + 49] << 8)
        stride = 64 if typ in (0, 2, 3) else 64 + ln
        if off + stride > stop:
            break
        e = raw[off:off + stride]
        ents.append({"idx": int.from_bytes(e[0:8], "little"), "term": int.from_bytes(e[8:16], "little"),
                     "req": int.from_bytes(e[16:24], "little"), "clt": int.from_bytes(e[24:26], "little"), "type": typ,
                     "sender": e[27], "len": ln if typ not in (0, 2, 3) else 0,
                     "data": e[48:64].hex() if typ == 2 else "",
                     "sha": hashlib.sha256(e[:28] + e[41:]).hexdigest()[:12]})      # reply bytes masked
        off += stride
    result.update(offsets={"head": head, "apply": apply, "commit": commit, "end": end}, entries=ents)
    with lock:
        result["replay"] = [{"bytes": v[0], "sha": v[1].hexdigest(), "first": v[2].hex()} for _, v in sorted(received.items())]
    with open(os.path.join(outdir, f"result{idx}.json.tmp"), "w") as f:
        json.dump(result, f)
    os.rename(os.path.join(outdir, f"result{idx}.json.tmp"), os.path.join(outdir, f"result{idx}.json"))
    t0 = time.time()
    while time.time() - t0 < 30 and not all(os.path.exists(os.path.join(outdir, f"result{i}.json")) for i in range(1, n)):
        time.sleep(0.05)
    os._exit(0)


if __name__ == "__main__":
    main()
