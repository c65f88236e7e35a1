"""Host->device upload of a large pageable NumPy array: the driver's own staging against a chunked copy through
page-locked buffers filled by torch's (multi-threaded) CPU copy.  python tools/prof_h2d.py [MB]"""
import sys, time
import numpy as np, torch

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 296
n = mb * (1 << 20) // 8
src = np.random.default_rng(0).normal(size=n)
dev = torch.device("cuda:0")
print("torch threads", torch.get_num_threads())
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    d = torch.as_tensor(src).to(dev); torch.cuda.synchronize()
    print(f"pageable .to(): {(time.perf_counter() - t) * 1e3:.1f} ms")
for chunk_mb in (16, 32, 64):
    c = chunk_mb * (1 << 20) // 8
    stage = [torch.empty(c, dtype=torch.float64, pin_memory=True) for _ in range(2)]
    ev = [torch.cuda.Event() for _ in range(2)]
    dst = torch.empty(n, dtype=torch.float64, device=dev)
    s = torch.as_tensor(src)
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        for i, o in enumerate(range(0, n, c)):
            b = i & 1
            if i >= 2:
                ev[b].synchronize()
            m = min(c, n - o)
            stage[b][:m].copy_(s[o:o + m])
            dst[o:o + m].copy_(stage[b][:m], non_blocking=True)
            ev[b].record()
        torch.cuda.synchronize()
        print(f"chunked {chunk_mb} MB via pinned staging: {(time.perf_counter() - t) * 1e3:.1f} ms")
    assert torch.equal(dst.cpu(), s)
t = time.perf_counter(); p = torch.empty(n, dtype=torch.float64, pin_memory=True); print(f"pinned alloc {mb} MB: {(time.perf_counter() - t) * 1e3:.1f} ms")
t = time.perf_counter(); p.copy_(torch.as_tensor(src)); print(f"cpu copy into pinned: {(time.perf_counter() - t) * 1e3:.1f} ms")
torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(p, non_blocking=True); torch.cuda.synchronize(); print(f"pinned -> device: {(time.perf_counter() - t) * 1e3:.1f} ms")
out = torch.empty(n * 6, dtype=torch.float64, device=dev); hp = torch.empty(n * 6, dtype=torch.float64, pin_memory=True)
torch.cuda.synchronize(); t = time.perf_counter(); hp.copy_(out, non_blocking=True); torch.cuda.synchronize(); print(f"device -> pinned {6 * mb} MB: {(time.perf_counter() - t) * 1e3:.1f} ms")
