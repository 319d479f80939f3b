import sys, time
sys.path.insert(0, ".")
from smm_jl_b200 import configs, _lib
for n in (1, 2, 3):
    cfg = configs.c1_serial_normal(n, exchange_mode=1)
    t = time.time()
    with _lib.BGPHandle(cfg) as h:
        print("created", flush=True)
        try:
            h.step(n)
            print("stepped", n, time.time() - t, flush=True)
        except Exception as e:
            print("ERR", e, flush=True)
