import sys
sys.path.insert(0, ".")
import numpy as np
from smm_jl_b200 import configs, _lib
from oracle import oracle_lib as O
N, S, n = 64, 2000, 6
for tag, kw in [("noswap", dict(min_improve=[1e9] * N)), ("swap", dict())]:
    cfg = configs.mvnormal(N, n, exchange_mode=1, n_sim=S, **kw)
    with _lib.BGPHandle(cfg) as h:
        h.step(n)
        tr = h.read_trace(1, n)
    ref = O.run(cfg, n, n_threads=8).trace
    for it in range(n):
        same = (tr.exchanged[it] == ref.exchanged[it])
        noex = (tr.exchanged[it] == 0) & (ref.exchanged[it] == 0)
        dv = np.abs(tr.value[it] - ref.value[it]) / np.abs(ref.value[it])
        print(tag, "iter", it + 1, "exch mismatches", int((~same).sum()), "| among chains untouched in both: value mismatches",
              int((dv[noex] > 1e-9).sum()), "acc mismatches", int((tr.accepted[it] != ref.accepted[it])[noex].sum()),
              "param mismatches", int((np.abs(tr.params[it] - ref.params[it]).max(axis=1) > 0)[noex].sum()))
