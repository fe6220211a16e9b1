"""One replica process of the proxy drop-in test (tests/test_gpu_proxy_dropin.py).

Loads, in this order, libapus_gpu.so, libapus_dare.so (our engine entry) and
oracle/_ref/libref_proxy.so (the reference's UNMODIFIED src/proxy/proxy.c + test doubles for
BerkeleyDB/libconfig) and then does what src/spec_hooks.cpp does: proxy_init() at start-up,
proxy_on_accept / proxy_on_read / proxy_on_close around the application's socket calls.
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
    idx, n, nconn, nreq, plen, outdir = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]),
                                          int(sys.argv[5]), sys.argv[6])
    nthreads = int(sys.argv[7]) if len(sys.argv) > 7 else 0      # > 0: the C driver of oracle/app_driver.inc (stub_drive)
    nsteps = int(sys.argv[8]) if len(sys.argv) > 8 else 1
    group = int(sys.argv[9]) if len(sys.argv) > 9 else n        # apus_colocate_followers: one process, `group` replicas
    leader = idx == 0
    received = {}
    lock = threading.Lock()

    def sink(port_holder):
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind(("127.0.0.1", 0))
        srv.listen(64)
        port_holder.append(srv.getsockname()[1])

        def serve(conn, k):
            h, nbytes = hashlib.sha256(), 0
            with lock:
                received[k] = [0, h]
            while True:
                d = conn.recv(1 << 16)
                if not d:
                    break
                h.update(d)
                nbytes += len(d)
                with lock:
                    received[k][0] = nbytes

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
    os.environ["group_size"] = str(group)
    os.environ["server_type"] = "start"
    os.environ["dare_log_file"] = os.path.join(outdir, f"dare{idx}.log")

    C.CDLL(os.path.join(ROOT, "apus_b200", "libapus_gpu.so"), mode=C.RTLD_GLOBAL)
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
    px.stub_db_size.restype = C.c_uint32
    px.stub_db_size.argtypes = [C.c_uint32]

    os.chdir(outdir)
    proxy = px.proxy_init(b"nodes.local.cfg", None)      # spec_hooks.cpp:33
    assert proxy
    result = {"idx": idx}
    if leader:
        t0 = time.time()
        while not dare.is_leader():
            assert time.time() - t0 < 60, "leader never came up"
            time.sleep(0.01)
        if nthreads > 0:
            # the same multi-threaded closed-loop driver the reference arm runs on its own stack (refstack_drive)
            px.stub_drive.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_double), C.c_int]
            lat = (C.c_uint64 * max(nreq, 1))()
            secs = C.c_double(0.0)
            steps = []
            for _ in range(nsteps):
                assert px.stub_drive(proxy, nthreads, nconn, nreq, plen, lat, C.byref(secs), 3) == 0
                al = sorted(int(x) for x in lat[:nreq])
                steps.append(dict(seconds=secs.value, requests=nreq, threads=nthreads,
                                  p50_us=al[len(al) // 2] / 1e3, p99_us=al[int(len(al) * 0.99)] / 1e3))
            result["steps"] = steps
            result["highest_rec"] = int(px.stub_highest_rec(proxy))
            result["p50_us"], result["p99_us"] = steps[-1]["p50_us"], steps[-1]["p99_us"]
            nreq_total = nreq * nsteps
        lat = []
        for c in range(nconn if nthreads == 0 else 0):
            px.proxy_on_accept(proxy, 100 + c)            # spec_hooks.cpp:116
        for i in range(nreq if nthreads == 0 else 0):
            c = i % nconn
            payload = bytes(((i * 31 + k) & 0xFF) for k in range(plen))
            buf = C.create_string_buffer(payload, plen)
            a = time.perf_counter_ns()
            px.proxy_on_read(proxy, buf, plen, 100 + c)   # spec_hooks.cpp:174: returns once committed
            lat.append(time.perf_counter_ns() - a)
        for c in range(nconn if nthreads == 0 else 0):
            px.proxy_on_close(proxy, 100 + c)             # spec_hooks.cpp:150
        if nthreads == 0:
            result["highest_rec"] = int(px.stub_highest_rec(proxy))
            lat.sort()
            result["p50_us"] = lat[len(lat) // 2] / 1e3
            result["p99_us"] = lat[int(len(lat) * 0.99)] / 1e3
    else:
        want = nreq * plen * nsteps
        t0 = time.time()
        while True:
            with lock:
                got = sum(v[0] for v in received.values())
                nc = len(received)
            if got >= want and nc >= nconn * nsteps:
                break
            if time.time() - t0 > float(os.environ.get("PROXY_RUN_TIMEOUT", "90")):
                break
            time.sleep(0.01)
        time.sleep(0.3)
        with lock:
            result["conns"] = len(received)
            result["bytes"] = sum(v[0] for v in received.values())
            result["sha"] = [received[k][1].hexdigest() for k in sorted(received)]
    cnt = int(px.stub_db_count())
    sizes = {}
    for i in range(cnt):
        s = int(px.stub_db_size(i))
        sizes[s] = sizes.get(s, 0) + 1
    result["db_records"] = cnt
    result["db_sizes"] = sizes
    with open(os.path.join(outdir, f"result{idx}.json.tmp"), "w") as f:
        json.dump(result, f)
    os.rename(os.path.join(outdir, f"result{idx}.json.tmp"), os.path.join(outdir, f"result{idx}.json"))
    # keep the kernels alive until every replica has reported
    t0 = time.time()
    while time.time() - t0 < 60 and not all(os.path.exists(os.path.join(outdir, f"result{i}.json")) for i in range(n)):
        time.sleep(0.05)
    os._exit(0)


if __name__ == "__main__":
    main()
