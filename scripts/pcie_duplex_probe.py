"""How fast are host<->device copies on this box, one direction at a time and both at once?  (decides whether xnb_step_host can hide
its position download behind the upload of the next chunk)"""
import torch

n = 100 * 1024 * 1024 // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory(); h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda"); d_out = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


gb = n * 8 / 1e9
t1, t2, t3 = timed(h2d), timed(d2h), timed(both)
print("H2D %.1f GB/s  D2H %.1f GB/s  both at once: %.1f GB/s aggregate (%.2f ms for 2 x 100 MB; serial would be %.2f ms)" % (gb / t1 * 1e3, gb / t2 * 1e3, 2 * gb / t3 * 1e3, t3, t1 + t2))
