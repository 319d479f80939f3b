"""Run each mode for a few seconds while sampling SM clock / power (is the kernel power- or clock-limited?)."""
import subprocess, sys, threading, time, statistics
sys.path.insert(0, ".")
from smm_jl_b200 import configs, _lib

def sample(stop, rows):
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu",
                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        rows.append(line.strip())
        if stop.is_set():
            break
    p.terminate()

for mode in (0, 1, "rng"):
    rows, stop = [], threading.Event()
    th = threading.Thread(target=sample, args=(stop, rows), daemon=True)
    if mode == "rng":
        th.start(); time.sleep(0.2)
        t0 = time.time()
        while time.time() - t0 < 2.0:
            ms, rate = _lib.rng_throughput(20000, 148 * 8)
        print("rng-only", rate / 1e9, "G normals/s")
    else:
        cfg = configs.mvnormal(256, 20100, exchange_mode=mode, sigma_update_steps=10**9)
        with _lib.BGPHandle(cfg) as h:
            h.step(100)
            th.start(); time.sleep(0.2)
            ms = h.step(20000)
            print("mode", mode, "us/iter", ms / 20000 * 1e3)
    stop.set(); time.sleep(0.1)
    clk = [float(r.split(",")[0]) for r in rows if r]
    pw = [float(r.split(",")[1]) for r in rows if r]
    caps = sum(1 for r in rows if "Active" in r.split(",")[2] and "Not" not in r.split(",")[2])
    n = len(clk)
    top = sorted(range(n), key=lambda i: -pw[i])[: max(n // 2, 1)]
    print("   samples", n, "clock under load: median %.0f min %.0f" % (statistics.median(clk[i] for i in top), min(clk[i] for i in top)),
          "power max %.0f W median(load) %.0f W" % (max(pw), statistics.median(pw[i] for i in top)), "sw_power_cap active samples", caps,
          "temp", rows[-1].split(",")[-1] if rows else None)
