#!/usr/bin/env python
"""Committed ops/s over request size x group size with ONE REPLICA PER GPU (north_star sweep: 64 B - 4 KB requests at
3/5/7 replica GPUs), in one process, value mode (requests generated on the device, bounded launches), with the NVLink
byte counters of the leader GPU (nvidia-smi nvlink -gt d: driver-level Tx/Rx KiB per link) read around every
configuration -- measured egress next to the algorithmic bytes (N-1)*(64+L) per op.

    python tools/sweep_spread.py [--sizes 64,256,1024,4096] [--replicas 3,5,7] [--ctas 16] [--out gpurun_out/sweep.txt]
"""
import argparse
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NVLINK_PEER_GBS = 770.0


def nvlink_kib(gpu):
    """(tx, rx) KiB summed over the links of one GPU"""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu)], capture_output=True, text=True, timeout=10).stdout
    except Exception:
        return None
    tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    return (tx, rx) if (tx or rx or "Data Tx" in out) else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="64,256,1024,4096")
    ap.add_argument("--replicas", default="3,5,7")
    ap.add_argument("--ctas", default="16")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--out", default="")
    ap.add_argument("--multicast", action="store_true", help="fabric mode: VMM regions bound to an NVSwitch multicast object, multimem.st in T5")
    a = ap.parse_args()
    import apus_b200 as A
    from apus_b200 import engine as E
    nd = A.lib().apus_device_count()
    lines = [("# MULTICAST: one multimem.st per 16 B chunk, the NVSwitch fans it out (leader egress 1x; algorithmic bytes unchanged)\n"
              if a.multicast else "") + f"# one replica per GPU ({nd} GPUs visible), value mode, device-generated requests, bounded launches; "
             f"roofline = algorithmic bytes (N-1)*(64+L)*ops/s against {NVLINK_PEER_GBS} GB/s measured NVLink peer copy",
             "replicas payload ctas batch ops_per_s alg_GBps frac_nvlink kernel_ms nvlink_tx_GB_leader tx_over_algorithmic T5_share"]
    print("\n".join(lines), flush=True)
    for n in [int(x) for x in a.replicas.split(",")]:
        if n > nd:
            continue
        for L in [int(x) for x in a.sizes.split(",")]:
            for ctas in [int(x) for x in a.ctas.split(",")]:
                stride = 64 + L
                batch = (1 << 19) if L <= 256 else max(8192, (48 << 20) // stride)
                K, W = a.steps, 2
                img = 0 if (2 + L) <= 80 else (2 + L + 15) // 16 * 16
                total = (K + W) * batch + 16
                slots = 1 << max(16, (total - 1).bit_length())
                rb = ((total * img + (1 << 20)) + 4095) // 4096 * 4096
                if rb // 16 > 0xFFFFFF:
                    continue
                flags = E.F_DEVICE_STATS | E.F_AUTOPRUNE | (E.F_FABRIC if a.multicast else 0)
                g = A.Group(n, devices=list(range(n)), log_size=0, ring_mode=A.RING_DEVICE, ring_slots=slots, ring_bytes=rb,
                            flags=flags, leader_ctas=ctas)
                try:
                    if a.multicast:
                        g.multicast()
                    g.prologue()
                    g.submit(E.CONNECT, 0, 1, b"")
                    g.run()
                    req, targets = 2, []
                    g.leader.defer(True)
                    for _ in range(K + W):
                        t0 = g.leader.submit_synth(batch, E.SEND, 0, req, L, 0xA5A50000 + L)
                        g.tickets = t0 + batch - 1
                        req += batch
                        targets.append(g.tickets)
                    g.leader.flush(); g.leader.defer(False)
                    for s in range(W):
                        g.launch(targets[s]); g.wait(120_000)
                    c0 = nvlink_kib(0)
                    st0 = g.leader.stats()
                    kms = 0.0
                    t0 = time.perf_counter()
                    for s in range(W, W + K):
                        g.launch(targets[s]); g.wait(120_000)
                        kms += g.leader.last_launch_ms()
                    t1 = time.perf_counter()
                    c1 = nvlink_kib(0)
                    st1 = g.leader.stats()
                    ops = K * batch / (t1 - t0)
                    alg = ops * (n - 1) * stride / 1e9
                    tx = ((c1[0] - c0[0]) * 1024 / 1e9) if (c0 and c1) else float("nan")
                    algb = K * batch * (n - 1) * stride / 1e9
                    ph = [b - a_ for a_, b in zip(st0["phase_ns"], st1["phase_ns"])]
                    t5 = ph[5] / max(1, sum(ph[:7]))
                    line = (f"{n} {L} {ctas} {batch} {ops:.0f} {alg:.1f} {alg / NVLINK_PEER_GBS:.3f} {kms / K:.3f} {tx:.2f} "
                            f"{tx / algb:.3f} {t5:.2f}")
                except Exception as ex:                      # noqa: BLE001
                    line = f"{n} {L} {ctas} FAILED {type(ex).__name__}: {str(ex)[:120]}"
                finally:
                    g.close()
                lines.append(line)
                print(line, flush=True)
    if a.out:
        with open(a.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
