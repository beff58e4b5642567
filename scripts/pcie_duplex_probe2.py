"""chunked upload -> tiny kernel -> download of the same chunk, three streams with event hand-offs: does the download of chunk c overlap
the upload of chunk c+1?  (A) download into a separate pinned buffer, (B) into the buffer the upload reads from (other region)"""
import torch

n = 2048000; K = 8; per = n // K
hin = [torch.randn(n, dtype=torch.float64).pin_memory() for _ in range(6)]
hout = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
d = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(6)]
s0, s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def run(same, chunks=True, download=True):
    ev0 = torch.cuda.Event(); ev0.record(torch.cuda.current_stream())
    for s in (s0, s1, s2):
        s.wait_event(ev0)
    rng = [(c * per, (c + 1) * per) for c in range(K)] if chunks else [(0, n)]
    for a, b in rng:
        with torch.cuda.stream(s1):
            for f in range(6):
                d[f][a:b].copy_(hin[f][a:b], non_blocking=True)
            e1 = torch.cuda.Event(); e1.record(s1)
        with torch.cuda.stream(s0):
            s0.wait_event(e1)
            for f in range(3):
                d[f][a:b].add_(d[3 + f][a:b], alpha=0.005)
            e2 = torch.cuda.Event(); e2.record(s0)
        if download:
            with torch.cuda.stream(s2):
                s2.wait_event(e2)
                for f in range(3):
                    (hin[f] if same else hout[f])[a:b].copy_(d[f][a:b], non_blocking=True)
    for s in (s0, s1, s2):
        torch.cuda.current_stream().wait_stream(s)


def timed(**kw):
    run(**kw); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run(**kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10


print("upload only (98 MB)                         %.3f ms" % timed(same=False, download=False))
print("upload + download, one chunk, separate bufs %.3f ms" % timed(same=False, chunks=False))
print("upload + download, 8 chunks, separate bufs  %.3f ms" % timed(same=False))
print("upload + download, 8 chunks, same buffers   %.3f ms" % timed(same=True))
