"""How to bring an 818 MB result block to host memory: pageable .cpu(), one pinned allocation, or chunks through two persistent pinned buffers."""
import time, numpy as np, torch
x = torch.randn(1023, 20, 1, 100, 100, device="cuda")
torch.cuda.synchronize()
def t(fn, name):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); print(f"{name}: {1e3 * (time.perf_counter() - t0):.1f} ms"); return r
t(lambda: x.cpu().numpy(), "pageable .cpu() (first)")
t(lambda: x.cpu().numpy(), "pageable .cpu() (second)")
def pinned():
    h = torch.empty(x.shape, dtype=x.dtype, pin_memory=True); h.copy_(x, non_blocking=True); torch.cuda.synchronize(); return h.numpy()
t(pinned, "pinned alloc + copy (first)")
t(pinned, "pinned alloc + copy (second, cached block)")
CH = 64 << 20
bufs = [torch.empty(CH // 4, dtype=torch.float32, pin_memory=True) for _ in range(2)]
def chunked():
    flat = x.reshape(-1); n = flat.numel(); out = np.empty(n, np.float32); step = CH // 4
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream()); evs = [None, None]
    offs = list(range(0, n, step))
    with torch.cuda.stream(s):
        for i, o in enumerate(offs[:2]):
            m = min(step, n - o); bufs[i % 2][:m].copy_(flat[o:o + m], non_blocking=True); evs[i % 2] = torch.cuda.Event(); evs[i % 2].record(s)
    for i, o in enumerate(offs):
        m = min(step, n - o); evs[i % 2].synchronize()
        out[o:o + m] = bufs[i % 2][:m].numpy()
        if i + 2 < len(offs):
            o2 = offs[i + 2]; m2 = min(step, n - o2)
            with torch.cuda.stream(s):
                bufs[i % 2][:m2].copy_(flat[o2:o2 + m2], non_blocking=True); evs[i % 2] = torch.cuda.Event(); evs[i % 2].record(s)
    return out.reshape(x.shape)
r = t(chunked, "chunked via 2 x 64 MB pinned (first)")
t(chunked, "chunked via 2 x 64 MB pinned (second)")
print("equal:", np.array_equal(r, x.cpu().numpy()))
