"""H2D / D2H bandwidth of the box with pinned memory through the library's own allocator (context for the e2e number)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
def main():
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, f in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        f(); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(5): f()
        torch.cuda.synchronize()
        print(name, "%.1f GB/s" % (5 * n / (time.perf_counter() - t) / 1e9))
    print("cpus", os.cpu_count())
if __name__ == "__main__":
    main()
