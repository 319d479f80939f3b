"""Regenerate tests/golden/bgp_golden.npz from the C++ oracle.

    python tests/golden/make_golden.py

These are REGRESSION pins of our own stream definition + restatement (the reference has no golden
vectors for this path and cannot run here: SURVEY.md 8c); the oracle itself is pinned by the Philox
known answers, the reference's behavioural tests and oracle/oracle_np.py.  The GPU parity tests compare
against these vectors too, so they do not need the oracle to travel.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_lib  # noqa: E402
from smm_jl_b200 import configs  # noqa: E402

out = {"normals_bits": oracle_lib.normals(1234, 0, 0, 1 << 28, 64).view(np.uint64),
       "zig_normals_bits": oracle_lib.zig_normals(1234, 0, 0, 1 << 28, 4096).view(np.uint64)}
for tag, cfg, n in (("c1", configs.c1_serial_normal(40), 40), ("mv", configs.mvnormal(8, 10), 10)):
    r = oracle_lib.run(cfg, n)
    for f in r.trace.FLOAT_FIELDS + r.trace.INT_FIELDS:
        out[f"{tag}_{f}"] = getattr(r.trace, f)
    out[f"{tag}_sigma"] = r.sigma
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bgp_golden.npz"), **out)
print("wrote bgp_golden.npz", {k: v.shape for k, v in out.items()})
