"""BASELINE config 4 in miniature against the REFERENCE ITSELF: unmodified memcached 1.4.21 processes (four worker threads
each) under the reference's unmodified interposer on the reference's own src/dare stack (oracle/_ref/interpose_ref.so, verbs
shim NIC).  Sixteen client connections set and get 1 KB values on the leader; every follower ends up with every key.  It is
the CPU-side counterpart of tests/test_gpu_memcached_dropin.py (same driver, tests/memcached_group.py)."""
import os

import pytest

import memcached_group as MG
import redis_group as RG

pytestmark = [pytest.mark.timeout(300)]


def test_memcached_replicated_through_reference_stack():
    for f in (MG.MEMCACHED, RG.INTERPOSE_REF):
        if not os.path.exists(f):
            pytest.skip(f"{f} absent (built only where /root/reference exists: oracle/build_memcached.sh, build_refapp.sh)")
    for attempt in range(2):                       # (the reference's start-up election may remove a slow replica: one more try)
        try:
            print(MG.run_memcached_group(3, 0, nconn=16, nkeys=60, vlen=1024, stack="refstack", base_port=21360 + 10 * attempt,
                                         startup_timeout=60))
            break
        except AssertionError as e:
            if attempt:
                raise
            print(f"first attempt failed on the reference stack ({str(e)[:200]}); trying once more")
