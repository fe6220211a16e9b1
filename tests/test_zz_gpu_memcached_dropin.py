"""BASELINE config 4 in miniature on the GPU engine: unmodified memcached 1.4.21 processes (four worker threads each) under
the reference's unmodified interposer linked on libapus_dare.so / libapus_gpu.so (oracle/_ref/interpose.so); sixteen client
connections set and get 1 KB values on the leader, every follower -- fed only through the GPU log -- ends up with every
key.  Same driver as the reference-side run (tests/memcached_group.py, tests/test_refstack_memcached.py, which passes).

NOT RUN ON HARDWARE: it was written after the round's GPU minutes were spent.  It is therefore opt-in
(APUS_TEST_UNVERIFIED=1) and sorts last, so that an untested test can neither stop `pytest -x` in front of the verified
ones nor pass for evidence it is not."""
import os

import pytest

import memcached_group as MG
import redis_group as RG

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.mark.skipif(os.environ.get("APUS_TEST_UNVERIFIED") != "1", reason="not yet run on a GPU box: opt in with APUS_TEST_UNVERIFIED=1")
def test_memcached_replicated_through_gpu_log():
    import __graft_entry__ as g
    g.build()
    for f in (MG.MEMCACHED, RG.INTERPOSE):
        if not os.path.exists(f):
            pytest.skip(f"{f} absent (built only where /root/reference exists)")
    import apus_b200
    nd = max(1, apus_b200.lib().apus_device_count())
    print(MG.run_memcached_group(3, nd, nconn=16, nkeys=100, vlen=1024, stack="gpu", base_port=21330))
