"""Diagnostic driver for tests/test_gpu_redis_dropin.py: start-up order / stagger variants (GPU box)."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import redis_group as T
import apus_b200
ndev = apus_b200.lib().apus_device_count()
for name, kw in (("simultaneous", {}), ("staggered 5s", dict(stagger=5.0)), ("followers first", dict(stagger=5.0, order=[2, 1, 0]))):
    print("=====", name, flush=True)
    try:
        print(T.run_redis_group(3, ndev, 300, 40, startup_timeout=45, **kw), flush=True)
    except BaseException:
        print(traceback.format_exc()[-3500:], flush=True)
