import sys
sys.path.insert(0, ".")
from smm_jl_b200 import configs, _lib
with _lib.BGPHandle(configs.mvnormal(256, 10, exchange_mode=1)) as h:
    for v, name in [(0, "persistent kernel barrier"), (1, "no fences (lower bound)"), (2, "release arrive only"), (3, "classic threadfence")]:
        print(f"variant {v} {name}: {h.barrier_bench(v, 2000):.2f} us per barrier")
