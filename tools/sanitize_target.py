"""Tiny runs of every kernel path for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from smm_jl_b200 import configs, _lib
for mode in (2, 1, 0):
    cfg = configs.mvnormal(6, 6, n_params=4, n_sim=600, exchange_mode=mode, sigma_update_steps=2)
    with _lib.BGPHandle(cfg) as h:
        h.step(6)
        tr = h.read_trace(1, 6)
        print("mode", mode, "values", np.round(tr.value[-1], 6))
cfg = configs.c1_serial_normal(4, n_sim=300, exchange_mode=2)
with _lib.BGPHandle(cfg) as h:
    h.step(4)
    v, m, st = h.eval_batch(np.array([[0.1, 0.2], [0.3, -0.4]]), noseed=1, rep0=3)
    print("c1 ok", v)
