"""One replica process of the join drill (tests/test_gpu_join.py): a follower dies, the leader removes it from the
configuration, a replacement started with server_type=join takes its slot (snapshot through the proxy callbacks, log over
NVLink), the group goes on.

    join_worker.py <idx> <n> <nconn> <nreqA> <nreqB> <plen> <outdir> <start|join>
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
    idx, n, nconn, nreqA, nreqB, plen, outdir, how = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]),
                                                       int(sys.argv[5]), int(sys.argv[6]), sys.argv[7], sys.argv[8])
    tag = f"{idx}{'j' if how == 'join' else ''}"
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
                received[k] = [0, h]
            while True:
                d = conn.recv(1 << 16)
                if not d:
                    break
                h.update(d)
                nb += len(d)
                with lock:
                    received[k][0] = nb

        k = 0
        while True:
            conn, _ = srv.accept()
            threading.Thread(target=serve, args=(conn, k), daemon=True).start()
            k += 1

    ph = []
    threading.Thread(target=sink, args=(ph,), daemon=True).start()
    while not ph:
        time.sleep(0.01)
    os.environ.update(stub_port=str(ph[0]), server_idx=str(idx), group_size=str(n), server_type=how,
                      dare_log_file=os.path.join(outdir, f"dare{tag}.log"))
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
    px.stub_db_count.restype = C.c_uint32
    px.stub_db_dumps.restype = C.c_uint32
    dare.apus_dare_replica.restype = C.c_void_p
    os.chdir(outdir)
    proxy = px.proxy_init(b"nodes.local.cfg", None)
    assert proxy
    t_up = time.time()
    while not dare.apus_dare_replica():
        assert time.time() - t_up < 90
        time.sleep(0.01)
    result = {"idx": idx, "how": how}

    def drive(first_fd, count, base):
        for c in range(nconn):
            px.proxy_on_accept(proxy, first_fd + c)
        for i in range(count):
            payload = bytes((((base + i) * 31 + k) & 0xFF) for k in range(plen))
            buf = C.create_string_buffer(payload, plen)
            px.proxy_on_read(proxy, buf, plen, first_fd + (i % nconn))
        for c in range(nconn):
            px.proxy_on_close(proxy, first_fd + c)

    done_file = os.path.join(outdir, "done.json")
    if idx == 0:
        t0 = time.time()
        while not dare.is_leader():
            assert time.time() - t0 < 90, "leader never came up"
            time.sleep(0.005)
        drive(100, nreqA, 0)
        with open(os.path.join(outdir, "phaseA_done"), "w") as f:
            f.write(str(int(px.stub_highest_rec(proxy))))
        while not os.path.exists(os.path.join(outdir, "phaseB_go")):
            # keep a trickle of load going so that commits keep flowing while the group changes
            time.sleep(0.01)
        drive(300, nreqB, 100000)
        result["highest_rec"] = int(px.stub_highest_rec(proxy))
        result["db_dumps"] = int(px.stub_db_dumps())
        time.sleep(0.5)
        with open(done_file + ".tmp", "w") as f:
            json.dump({"leader": 0}, f)
        os.rename(done_file + ".tmp", done_file)
    else:
        while not os.path.exists(done_file):
            time.sleep(0.01)
    time.sleep(0.7)
    rep = C.c_void_p(dare.apus_dare_replica())
    gpu.apus_log_offsets.argtypes = [C.c_void_p, C.c_void_p]
    gpu.apus_log_read.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    offs = (C.c_uint64 * 8)()
    assert gpu.apus_log_offsets(rep, offs) == 0
    head, apply, commit, end, L = int(offs[0]), int(offs[1]), int(offs[2]), int(offs[3]), int(offs[7])
    stop = end if end != L else 0
    img = (C.c_uint8 * max(stop, 1))()
    if stop:
        assert gpu.apus_log_read(rep, 0, stop, img) == 0
    raw = bytes(img[:stop])
    ents, off = [], head
    while off + 64 <= stop:
        typ = raw[off + 26]
        ln = raw[off + 48] | (raw[off + 49] << 8)
        stride = 64 if typ in (0, 2, 3) else 64 + ln
        if off + stride > stop:
            break
        e = raw[off:off + stride]
        ents.append({"idx": int.from_bytes(e[0:8], "little"), "term": int.from_bytes(e[8:16], "little"), "type": typ, "sender": e[27],
                     "data": e[48:64].hex() if typ == 2 else "", "sha": hashlib.sha256(e[:28] + e[41:]).hexdigest()[:12]})
        off += stride
    result.update(offsets={"head": head, "apply": apply, "commit": commit, "end": end}, entries=ents,
                  db_records=int(px.stub_db_count()))
    with lock:
        result["replay"] = [{"bytes": v[0], "sha": v[1].hexdigest()} for _, v in sorted(received.items())]
    with open(os.path.join(outdir, f"result{tag}.json.tmp"), "w") as f:
        json.dump(result, f)
    os.rename(os.path.join(outdir, f"result{tag}.json.tmp"), os.path.join(outdir, f"result{tag}.json"))
    time.sleep(3.0)
    os._exit(0)


if __name__ == "__main__":
    main()
