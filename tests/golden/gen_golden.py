#!/usr/bin/env python
"""Generate tests/golden/oracle_golden.json from the COMPILED REFERENCE HEADER
(oracle/_ref/libapus_ref.so = /root/reference/src/include/dare/dare_log.h built
by oracle/Makefile).  Run in the build container, where /root/reference exists:

    python tests/golden/gen_golden.py

The fixture pins the restated oracle on machines that have no reference tree
(the GPU box): tests/test_golden.py replays the same scenarios through
oracle/liboracle_port.so and compares offsets, return values, apply traces and
SHA-256 of every log image.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import orc as O          # noqa: E402
import scenarios         # noqa: E402

if __name__ == "__main__":
    O.build_oracle()
    assert O.have_ref(), "needs /root/reference to build oracle/_ref"
    ref = O.Oracle("ref")
    out = dict(source="oracle/_ref/libapus_ref.so (reference dare_log.h, unmodified)",
               reference_commit="896959f", scenarios=scenarios.all_scenarios(ref))
    path = os.path.join(HERE, "oracle_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path, len(out["scenarios"]), "scenarios")
