"""Probe: how long do NVML queries take and do they stall concurrent CUDA work? (bench.py ClockSampler tuning)"""
import time, threading, torch, pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
x = torch.zeros(1 << 20, device="cuda")
torch.cuda.synchronize()
def t(f, n=20):
    a = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - a) / n * 1e3
print("clock_sm ms", t(lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
print("reasons ms", t(lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
print("maxclock ms", t(lambda: pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)))
# effect on a launch loop
def loop(n=2000):
    torch.cuda.synchronize(); a = time.perf_counter()
    worst = 0
    for _ in range(n):
        b = time.perf_counter(); x.add_(1.0); torch.cuda.synchronize(); worst = max(worst, time.perf_counter() - b)
    return (time.perf_counter() - a) / n * 1e3, worst * 1e3
print("launch loop alone: avg ms, worst ms", loop())
for what in ("clock", "reasons", "both"):
    stop = threading.Event()
    def s():
        while not stop.is_set():
            if what in ("clock", "both"): pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            if what in ("reasons", "both"): pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            stop.wait(0.02)
    th = threading.Thread(target=s, daemon=True); th.start()
    print(f"launch loop with {what} sampler @20ms: avg ms, worst ms", loop())
    stop.set(); th.join()
