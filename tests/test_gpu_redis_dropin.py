"""Drop-in test of the whole operator surface (BASELINE config 3 in miniature):
an UNMODIFIED redis-server 2.8.17 per replica under interpose.so = the reference's unmodified spec_hooks.cpp +
proxy.c + db-interface.c + config-proxy.c (real libconfig 1.4.9 and BerkeleyDB 5.1.29 from the vendored tarballs)
linked against libapus_dare.so + libapus_gpu.so instead of libdare.a/-lev/-libverbs (oracle/build_refapp.sh).
redis-benchmark / redis-cli talk to the leader's port only; every read() of the leader is committed through
the GPU log before Redis sees it, followers replay the committed byte streams into their own Redis.
The data sets must then be identical (DEBUG DIGEST = Redis's own SHA-1 of the keyspace).
The same scenario runs against the reference's own stack in tests/test_refstack_redis.py."""
import os

import pytest

import redis_group as RG

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(420)]


def test_redis_replicated_through_gpu_log():
    import __graft_entry__ as g
    g.build()
    for f in (RG.SERVER, RG.BENCH, RG.CLI, RG.INTERPOSE):
        if not os.path.exists(f):
            pytest.skip(f"{f} absent (built only where /root/reference exists: oracle/build_refapp.sh)")
    import apus_b200
    ndev = apus_b200.lib().apus_device_count()
    n = 3
    # replicas that share one GPU are time-sliced by the driver (milliseconds per commit): keep the load small there
    nbench, nlist = (20000, 2000) if ndev >= n else (1200, 150)
    print(RG.run_redis_group(n, ndev, nbench, nlist))
