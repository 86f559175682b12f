import os, time, torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "torch default threads", torch.get_num_threads())
try:
    print("cgroup cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e: print("no cgroup", e)
a = torch.randn(2048, 2048); b = torch.randn(2048, 8192)
for n in (4, 8, 16, 32, 64, 128):
    torch.set_num_threads(n)
    a @ b
    t = time.perf_counter()
    for _ in range(3): a @ b
    dt = (time.perf_counter() - t) / 3
    print(n, "threads:", round(2*2048*2048*8192/dt/1e9, 1), "GFLOP/s")
