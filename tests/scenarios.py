"""Scenario scripts shared by the golden generator (run against the compiled
reference header) and test_golden.py (run against the restatement, anywhere)."""
import hashlib

import numpy as np

import orc as O
import streams as S


def _h(obj) -> str:
    return hashlib.sha256(repr(obj).encode()).hexdigest()[:16]


def single_log(oracle, name, length, stream, term=1, head_follows=0):
    log = O.Log(oracle, length)
    rets = []
    for k, (typ, clt, rid, payload) in enumerate(stream):
        rets.append(log.append(term, rid, clt, typ, O.cmd_image(payload)))
        if head_follows and k % head_follows == 0:
            o = log.offsets()
            if o["end"] != o["len"]:
                log.set_offsets(head=o["tail"], apply=o["tail"], commit=o["tail"])
    off = log.offsets()
    img = log.image()
    log.close()
    return dict(name=name, offsets=off, rets=_h(rets), n_rets=len(rets), last_ret=rets[-1],
                fnv=f"{O.fnv1a(img) if length <= (1 << 20) else 0:016x}",
                sha=hashlib.sha256(img.tobytes()).hexdigest())


def cluster(oracle, name, n, length, stream, prune_every=0):
    c = O.Cluster(oracle, n, leader=0, term=1, length=length)
    c.prologue()
    cido = [(O.u64)(0) for _ in range(n)]
    heads = []
    for k, (typ, clt, rid, payload) in enumerate(stream):
        c.submit(typ, clt, rid, O.cmd_image(payload))
        if k % 3 == 2:
            c.round()
        if prune_every and k % prune_every == prune_every - 1:
            c.round()
            heads.append(int(c.prune()))
            c.round()
            for i in range(1, n):
                c.poll_head(i, cido[i])
    c.round()
    out = dict(name=name, n=n,
               offsets=[c.offsets(i) for i in range(n)],
               sha=[hashlib.sha256(c.image(i).tobytes()).hexdigest() for i in range(n)],
               applied=[_h(c.applied(i)) for i in range(n)],
               n_applied=[len(c.applied(i)) for i in range(n)],
               heads=_h(heads), bytes_replicated=int(c.bytes_replicated()),
               update_state=int(c.update_state_calls()))
    c.close()
    return out


def all_scenarios(oracle):
    res = []
    res.append(single_log(oracle, "kat7", O.LOG_SIZE,
                          [(O.SEND, 0x0100, i + 1, S.payload_kat(i, ln))
                           for i, ln in enumerate([64, 64, 64, 100, 4096])]
                          + [(O.CONNECT, 0x0100, 6, b"")]))
    res.append(single_log(oracle, "uniform64_x4096", O.LOG_SIZE, S.uniform_stream(4096, 64)))
    res.append(single_log(oracle, "uniform1024_x512_c16", O.LOG_SIZE, S.uniform_stream(512, 1024, conns=16)))
    res.append(single_log(oracle, "ragged300_x2000", O.LOG_SIZE, S.ragged_stream(2000, 300, seed=5, close_every=100)))
    res.append(single_log(oracle, "ragged_wrap_16k", 16384, S.ragged_stream(1500, 200, seed=6), head_follows=5))
    res.append(single_log(oracle, "ragged_wrap_4k", 4096, S.ragged_stream(800, 90, seed=7), head_follows=3))
    res.append(single_log(oracle, "max_len_65535", O.LOG_SIZE,
                          [(O.CONNECT, 1, 1, b"")] + [(O.SEND, 1, 2 + i, bytes([i]) * 65535) for i in range(3)]))
    for n in (1, 3, 5, 7):
        res.append(cluster(oracle, f"cluster{n}_ragged", n, 1 << 20, S.ragged_stream(500, 256, seed=10 + n, close_every=60)))
    res.append(cluster(oracle, "cluster5_uniform64", 5, 1 << 20, S.uniform_stream(2000, 64, conns=4)))
    res.append(cluster(oracle, "cluster3_wrap_prune", 3, 16384, S.ragged_stream(900, 150, conns=2, seed=77), prune_every=5))
    res.append(cluster(oracle, "cluster5_wrap_prune", 5, 32768, S.ragged_stream(1200, 333, conns=3, seed=78), prune_every=7))
    return res
