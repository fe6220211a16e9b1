"""ctypes binding of include/apus_gpu.h (the C ABI is the product boundary).

`Group` mirrors how the reference deploys a Paxos group: `group_size` replicas,
`server_idx` 0..n-1 (env vars of benchmarks/run.sh:26), one of them the leader.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libapus_gpu.so")

APUS_OK, APUS_ERROR, APUS_RETRY = 0, 1, -1
NOOP, CSM, CONFIG, HEAD, CONNECT, SEND, CLOSE = 0, 1, 2, 3, 4, 5, 6
RING_HOST_MAPPED, RING_DEVICE = 0, 1
LOG_SIZE = 16384 * 4096
MAX_SERVERS = 13
F_FENCED_ACK, F_DEVICE_STATS, F_AUTOPRUNE, F_FOLLOWER_WALK, F_EXPLICIT = 0x1, 0x2, 0x4, 0x8, 0x80000000
F_HOST_APPLY, F_NO_EXPRESS, F_PROFILE, F_FABRIC = 0x10, 0x20, 0x40, 0x100
UINT64_MAX = (1 << 64) - 1

u64, u32, u16, u8, i64, i32 = C.c_uint64, C.c_uint32, C.c_uint16, C.c_uint8, C.c_int64, C.c_int32


class ApusError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("struct_size", u32), ("device", i32), ("server_idx", u8), ("group_size", u8),
                ("leader_idx", u8), ("ring_mode", u8), ("flags", u32), ("term", u64),
                ("log_size", u64), ("ring_slots", u32), ("ring_bytes", u32), ("leader_ctas", u32), ("reserved", u32),
                ("hb_period_us", u32), ("hb_timeout_us", u32)]


class PeerHandle(C.Structure):
    _fields_ = [("bytes", u8 * 128)]


class LogOffsets(C.Structure):
    _fields_ = [(k, u64) for k in ("head", "apply", "commit", "end", "tail", "old_end", "old_commit", "len")]


class Stats(C.Structure):
    _fields_ = [(k, u64) for k in ("tickets_submitted", "tickets_consumed", "tickets_committed",
                                   "entries_acked", "bytes_replicated", "batches", "kernel_launches",
                                   "lat_samples", "auto_heads", "entries_published")] + [("phase_ns", u64 * 8), ("turn_ns", u64 * 8)]


_lib = None

EXPORTS = [
    "apus_abi_version", "apus_last_error", "apus_device_count", "apus_replica_create",
    "apus_replica_destroy", "apus_replica_export", "apus_replica_connect", "apus_replicas_launch",
    "apus_replica_wait", "apus_replica_last_launch_ms", "apus_replicas_stop", "apus_submit",
    "apus_submit_batch", "apus_submit_defer", "apus_submit_flush", "apus_committed_tickets",
    "apus_progress", "apus_wait_committed", "apus_closed_loop", "apus_log_offsets", "apus_log_read", "apus_get_stats",
    "apus_latency_samples", "apus_set_head", "apus_remote_apply_offsets",
    "apus_submit_uniform", "apus_submit_synth", "apus_synth_byte", "apus_submit_release", "apus_set_applied",
    "apus_log_read_range", "apus_leader_suspect", "apus_last_commit_ns",
    "apus_ctl_read", "apus_ctl_set_sid", "apus_ctl_reset_votes", "apus_ctl_clear_vote_request", "apus_ctl_send_vote_request",
    "apus_ctl_send_vote_ack", "apus_ctl_last_entry", "apus_ctl_adjust_follower", "apus_replica_set_role",
    "apus_replica_disconnect", "apus_follower_beats", "apus_device_numa_node", "apus_group_multicast", "apus_ctl_heartbeat",
]


def load_library(path=LIB_PATH):
    """Load libapus_gpu.so; raises (no fallback) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise ApusError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(make -C apus_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.apus_abi_version.restype = C.c_int
    L.apus_last_error.restype = C.c_char_p
    L.apus_device_count.restype = C.c_int
    L.apus_replica_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.apus_replica_destroy.argtypes = [vp]
    L.apus_replica_destroy.restype = None
    L.apus_replica_export.argtypes = [vp, C.POINTER(PeerHandle)]
    L.apus_replica_connect.argtypes = [vp, u8, C.POINTER(PeerHandle)]
    L.apus_replicas_launch.argtypes = [C.POINTER(vp), C.c_int, u64]
    L.apus_replica_wait.argtypes = [vp, i64]
    L.apus_replica_last_launch_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.apus_replicas_stop.argtypes = [C.POINTER(vp), C.c_int]
    L.apus_submit.argtypes = [vp, u8, u16, u64, vp, u16, C.POINTER(u64)]
    L.apus_submit_batch.argtypes = [vp, u32, vp, vp, vp, vp, vp, C.c_size_t, C.POINTER(u64)]
    L.apus_submit_defer.argtypes = [vp, C.c_int]
    L.apus_submit_flush.argtypes = [vp]
    L.apus_committed_tickets.argtypes = [vp]
    L.apus_committed_tickets.restype = u64
    L.apus_wait_committed.argtypes = [vp, u64, i64]
    L.apus_closed_loop.argtypes = [vp, u32, u16, u16, u64, vp]
    L.apus_progress.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.apus_log_offsets.argtypes = [vp, C.POINTER(LogOffsets)]
    L.apus_log_read.argtypes = [vp, u64, u64, vp]
    L.apus_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.apus_latency_samples.argtypes = [vp, vp, u32, C.POINTER(u32)]
    L.apus_set_head.argtypes = [vp, u64]
    L.apus_remote_apply_offsets.argtypes = [vp, C.POINTER(u64)]
    if hasattr(L, "apus_submit_uniform"):       # ABI 2
        L.apus_submit_uniform.argtypes = [vp, u32, u8, u16, u64, u16, vp, C.c_size_t, C.POINTER(u64)]
        L.apus_submit_synth.argtypes = [vp, u32, u8, u16, u64, u16, u32, C.POINTER(u64)]
        L.apus_synth_byte.argtypes = [u32, u64, u32]
        L.apus_synth_byte.restype = u8
        L.apus_submit_release.argtypes = [vp, u64]
        L.apus_set_applied.argtypes = [vp, u64]
        L.apus_log_read_range.argtypes = [vp, u64, u64, vp, u64, C.POINTER(u64)]
        L.apus_leader_suspect.argtypes = [vp]
        L.apus_leader_suspect.restype = u64
        L.apus_last_commit_ns.argtypes = [vp]
        L.apus_last_commit_ns.restype = u64
    _lib = L
    return L


def lib():
    return load_library()


def _ck(rc, what):
    if rc == APUS_OK:
        return
    msg = lib().apus_last_error().decode(errors="replace")
    if rc == APUS_RETRY:
        raise BlockingIOError(f"{what}: {msg}")
    raise ApusError(f"{what}: {msg}")


class Replica:
    def __init__(self, device, server_idx, group_size, leader_idx=0, term=1, log_size=0,
                 ring_mode=RING_HOST_MAPPED, ring_slots=0, ring_bytes=0, flags=None, leader_ctas=0,
                 hb_period_us=0, hb_timeout_us=0):
        cfg = Config()
        cfg.leader_ctas = leader_ctas
        cfg.hb_period_us, cfg.hb_timeout_us = hb_period_us, hb_timeout_us
        cfg.struct_size = C.sizeof(Config) if lib().apus_abi_version() >= 2 else 48
        cfg.device, cfg.server_idx, cfg.group_size, cfg.leader_idx = device, server_idx, group_size, leader_idx
        cfg.ring_mode, cfg.term, cfg.log_size = ring_mode, term, log_size
        cfg.ring_slots, cfg.ring_bytes = ring_slots, ring_bytes
        cfg.flags = 0 if flags is None else (flags | F_EXPLICIT)
        self.h = C.c_void_p()
        _ck(lib().apus_replica_create(C.byref(cfg), C.byref(self.h)), "apus_replica_create")
        self.cfg = cfg
        self.device, self.idx, self.n, self.leader = device, server_idx, group_size, leader_idx
        self.log_len = log_size or LOG_SIZE

    @property
    def is_leader(self):
        return self.idx == self.leader

    def close(self):
        if self.h:
            lib().apus_replica_destroy(self.h)
            self.h = C.c_void_p()

    def export(self) -> bytes:
        ph = PeerHandle()
        _ck(lib().apus_replica_export(self.h, C.byref(ph)), "apus_replica_export")
        return bytes(ph.bytes)

    def connect(self, peer_idx, blob: bytes):
        ph = PeerHandle()
        C.memmove(ph.bytes, blob, 128)
        _ck(lib().apus_replica_connect(self.h, peer_idx, C.byref(ph)), "apus_replica_connect")

    def wait(self, timeout_ms=-1):
        _ck(lib().apus_replica_wait(self.h, timeout_ms), "apus_replica_wait")

    def last_launch_ms(self):
        ms = C.c_float()
        _ck(lib().apus_replica_last_launch_ms(self.h, C.byref(ms)), "apus_replica_last_launch_ms")
        return float(ms.value)

    def submit(self, typ, conn, req_id, payload=b""):
        t = u64()
        buf = (u8 * max(len(payload), 1)).from_buffer_copy(bytes(payload) or b"\0")
        _ck(lib().apus_submit(self.h, typ, conn, req_id, C.cast(buf, C.c_void_p), len(payload), C.byref(t)),
            "apus_submit")
        return int(t.value)

    def submit_batch(self, types, conns, req_ids, lens, payloads, stride):
        """numpy arrays: types u8[n], conns u16[n], req_ids u64[n], lens u16[n], payloads u8[n*stride]."""
        n = len(types)
        types = np.ascontiguousarray(types, dtype=np.uint8)
        conns = np.ascontiguousarray(conns, dtype=np.uint16)
        req_ids = np.ascontiguousarray(req_ids, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint16)
        pp = None
        if payloads is not None:
            payloads = np.ascontiguousarray(payloads, dtype=np.uint8)
            pp = payloads.ctypes.data
        t = u64()
        _ck(lib().apus_submit_batch(self.h, n, types.ctypes.data, conns.ctypes.data, req_ids.ctypes.data,
                                    lens.ctypes.data, pp, stride, C.byref(t)), "apus_submit_batch")
        return int(t.value)

    def submit_uniform(self, n, typ, conn, first_req_id, length, payloads, stride=None):
        """n requests of one shape, filled by the engine's host threads (apus_submit_uniform)."""
        pp = None
        if payloads is not None and length:
            payloads = np.ascontiguousarray(payloads, dtype=np.uint8)
            pp = payloads.ctypes.data
        t = u64()
        _ck(lib().apus_submit_uniform(self.h, n, typ, conn, first_req_id, length, pp, length if stride is None else stride,
                                      C.byref(t)), "apus_submit_uniform")
        return int(t.value)

    def submit_synth(self, n, typ, conn, first_req_id, length, seed):
        """n device-generated requests written straight into the HBM submission ring."""
        t = u64()
        _ck(lib().apus_submit_synth(self.h, n, typ, conn, first_req_id, length, seed, C.byref(t)), "apus_submit_synth")
        return int(t.value)

    def release(self, ticket):
        _ck(lib().apus_submit_release(self.h, ticket), "apus_submit_release")

    def set_applied(self, offset):
        _ck(lib().apus_set_applied(self.h, offset), "apus_set_applied")

    def read_range(self, start, stop, cap=1 << 20):
        out = np.empty(cap, dtype=np.uint8)
        got = u64()
        _ck(lib().apus_log_read_range(self.h, start, stop, out.ctypes.data, cap, C.byref(got)), "apus_log_read_range")
        return out[: got.value].copy()

    def leader_suspect(self):
        return int(lib().apus_leader_suspect(self.h))

    def last_commit_ns(self):
        return int(lib().apus_last_commit_ns(self.h))

    def defer(self, on=True):
        _ck(lib().apus_submit_defer(self.h, 1 if on else 0), "apus_submit_defer")

    def flush(self):
        _ck(lib().apus_submit_flush(self.h), "apus_submit_flush")

    def committed(self):
        return int(lib().apus_committed_tickets(self.h))

    def progress(self):
        off, cnt = u64(), u64()
        _ck(lib().apus_progress(self.h, C.byref(off), C.byref(cnt)), "apus_progress")
        return int(off.value), int(cnt.value)

    def wait_committed(self, ticket, timeout_us=10_000_000):
        _ck(lib().apus_wait_committed(self.h, ticket, timeout_us), "apus_wait_committed")

    def closed_loop(self, n, payload_len, conn, first_req_id):
        """n requests, one in flight; returns host-clock latencies in ns (uint32 array)."""
        out = np.empty(n, dtype=np.uint32)
        _ck(lib().apus_closed_loop(self.h, n, payload_len, conn, first_req_id, out.ctypes.data), "apus_closed_loop")
        return out

    def offsets(self):
        o = LogOffsets()
        _ck(lib().apus_log_offsets(self.h, C.byref(o)), "apus_log_offsets")
        return {k: int(getattr(o, k)) for k, _ in LogOffsets._fields_}

    def image(self, start=0, stop=None):
        stop = self.log_len if stop is None else stop
        out = np.empty(stop - start, dtype=np.uint8)
        if stop > start:
            _ck(lib().apus_log_read(self.h, start, stop - start, out.ctypes.data), "apus_log_read")
        return out

    def stats(self):
        s = Stats()
        _ck(lib().apus_get_stats(self.h, C.byref(s)), "apus_get_stats")
        d = {k: int(getattr(s, k)) for k, _ in Stats._fields_ if k not in ("phase_ns", "turn_ns")}
        d["phase_ns"] = [int(x) for x in s.phase_ns]
        d["turn_ns"] = [int(x) for x in s.turn_ns]
        return d

    def latency_ns(self, max_samples=65536):
        out = np.empty(max_samples, dtype=np.uint32)
        n = u32()
        _ck(lib().apus_latency_samples(self.h, out.ctypes.data, max_samples, C.byref(n)), "apus_latency_samples")
        return out[: n.value].copy()

    def set_head(self, head):
        _ck(lib().apus_set_head(self.h, head), "apus_set_head")

    def remote_apply_offsets(self):
        arr = (u64 * MAX_SERVERS)()
        _ck(lib().apus_remote_apply_offsets(self.h, arr), "apus_remote_apply_offsets")
        return [int(x) for x in arr]


def synth_payload(seed: int, req_id: int, length: int) -> bytes:
    """Payload of the device-generated request `req_id` (apus_submit_synth), computed on the host with numpy --
    the same integer mix the fill kernel runs (apus_engine.cu: synth_word)."""
    if length == 0:
        return b""
    w = np.arange((length + 3) // 4, dtype=np.uint64)
    x = (np.uint64(seed) ^ np.uint64((req_id * 0x9E3779B1) & 0xFFFFFFFF) ^ np.uint64(((req_id >> 32) * 0x7F4A7C15) & 0xFFFFFFFF)
         ^ ((w * np.uint64(0x85EBCA77)) & np.uint64(0xFFFFFFFF)))
    M = np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7FEB352D)) & M
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846CA68B)) & M
    x ^= x >> np.uint64(16)
    return x.astype(np.uint32).view(np.uint8)[:length].tobytes()


def pin_to_device_node(device: int):
    """Run this process on the CPUs of the NUMA node next to `device` (its pinned rings are then allocated there too).
    Returns the node, or None when the topology is not exposed."""
    try:
        lib().apus_device_numa_node.argtypes = [C.c_int]
        node = int(lib().apus_device_numa_node(device))
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:                                        # noqa: BLE001 - best effort
        pass
    return None


def cid_image(n: int) -> bytes:
    """dare_cid_t of a fresh stable group (dare_server.c:285-291)."""
    return (0).to_bytes(8, "little") + bytes([n, 0, 0, 0]) + ((1 << n) - 1).to_bytes(4, "little")


class Group:
    """All replicas of one Paxos group inside this process (tests, 1..8 GPUs)."""

    def __init__(self, n, devices=None, leader=0, term=1, log_size=0, ring_mode=RING_HOST_MAPPED,
                 ring_slots=0, ring_bytes=0, flags=None, leader_ctas=0, hb_period_us=0, hb_timeout_us=0):
        ndev = lib().apus_device_count()
        if ndev <= 0:
            raise ApusError("no CUDA device visible: the engine has no CPU fallback")
        if devices is None:
            devices = [i % ndev for i in range(n)]
        self.n, self.leader_idx, self.devices = n, leader, list(devices)
        self.replicas = [Replica(devices[i], i, n, leader, term, log_size, ring_mode, ring_slots, ring_bytes, flags,
                                 leader_ctas, hb_period_us, hb_timeout_us) for i in range(n)]
        blobs = [r.export() for r in self.replicas]
        for r in self.replicas:
            for j, b in enumerate(blobs):
                if j != r.idx:
                    r.connect(j, b)
        self.tickets = 0

    @property
    def leader(self) -> Replica:
        return self.replicas[self.leader_idx]

    def close(self):
        for r in self.replicas:
            r.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        try:
            self.stop()
        except Exception:
            pass
        self.close()

    def launch(self, target=None):
        """One fused launch per device; `target` = cumulative tickets (None -> all submitted)."""
        target = self.tickets if target is None else target
        by_dev = {}
        for r in self.replicas:
            by_dev.setdefault(r.device, []).append(r)
        # followers first, so a leader never waits for a CTA that is not scheduled yet
        for dev, rs in sorted(by_dev.items(), key=lambda kv: any(r.is_leader for r in kv[1])):
            arr = (C.c_void_p * len(rs))(*[r.h for r in rs])
            _ck(lib().apus_replicas_launch(arr, len(rs), target), "apus_replicas_launch")

    def wait(self, timeout_ms=60_000):
        for r in self.replicas:
            r.wait(timeout_ms)

    def stop(self):
        arr = (C.c_void_p * self.n)(*[r.h for r in self.replicas])
        _ck(lib().apus_replicas_stop(arr, self.n), "apus_replicas_stop")

    def multicast(self):
        """Bind the replicas' regions (created with F_FABRIC, one GPU each) to an NVSwitch multicast object."""
        arr = (C.c_void_p * self.n)(*[r.h for r in self.replicas])
        lib().apus_group_multicast.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        _ck(lib().apus_group_multicast(arr, self.n), "apus_group_multicast")

    def prologue(self):
        """Blank CONFIG entry the election winner appends (dare_server.c:1412-1421);
        none when the group has a single member (dare_server.c:416-424)."""
        if self.n == 1:
            return 0
        self.tickets = self.leader.submit(CONFIG, 0, 0, cid_image(self.n))
        return self.tickets

    def submit(self, typ, conn, req_id, payload=b""):
        self.tickets = self.leader.submit(typ, conn, req_id, payload)
        return self.tickets

    def submit_stream(self, stream):
        """stream: iterable of (type, connection_id, req_id, payload) -- tailq_entry_t fields."""
        self.leader.defer(True)
        try:
            for typ, conn, rid, payload in stream:
                self.tickets = self.leader.submit(typ, conn, rid, payload)
        finally:
            self.leader.flush()
            self.leader.defer(False)
        return self.tickets

    def submit_uniform(self, n_req, length, conn, first_req_id, payloads=None, typ=SEND):
        """n_req requests of `length` bytes on one connection (batch ABI call)."""
        types = np.full(n_req, typ, dtype=np.uint8)
        conns = np.full(n_req, conn, dtype=np.uint16)
        req_ids = np.arange(first_req_id, first_req_id + n_req, dtype=np.uint64)
        lens = np.full(n_req, length, dtype=np.uint16)
        t0 = self.leader.submit_batch(types, conns, req_ids, lens, payloads, length)
        self.tickets = t0 + n_req - 1
        return self.tickets

    def run(self, timeout_ms=60_000):
        """Launch for everything submitted so far and wait until it is committed everywhere."""
        self.launch()
        self.wait(timeout_ms)
