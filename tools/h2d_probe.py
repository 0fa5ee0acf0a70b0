"""Developer tool (GPU): host->device bandwidth of a 128 MiB upload from (a) torch pinned memory, (b) write-combined
pinned memory (cudaHostAllocWriteCombined), (c) the same split over two streams."""
import ctypes, torch, numpy as np, time
rt = ctypes.CDLL("libcudart.so.12")
n = 128 << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
dev2 = torch.empty(n, dtype=torch.uint8, device="cuda")
def bw(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return n * reps / (time.perf_counter() - t0) / 1e9
pin = torch.empty(n, dtype=torch.uint8).pin_memory(); pin.fill_(3)
print("torch pinned           %.1f GB/s" % bw(lambda: dev.copy_(pin, non_blocking=True)))
for flags, name in ((0, "cudaHostAlloc default "), (4, "cudaHostAlloc WC      "), (1, "cudaHostAlloc portable")):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flags)) == 0
    ctypes.memset(p, 5, n)
    s = torch.cuda.current_stream().cuda_stream
    f = lambda: rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), p, ctypes.c_size_t(n), 1, ctypes.c_void_p(s))
    print("%s %.1f GB/s" % (name, bw(f)))
    if flags == 4:
        s2 = torch.cuda.Stream()
        h = n // 2
        def two():
            rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), p, ctypes.c_size_t(h), 1, ctypes.c_void_p(s))
            rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr() + h), ctypes.c_void_p(p.value + h), ctypes.c_size_t(h), 1, ctypes.c_void_p(s2.cuda_stream))
        print("WC, two streams        %.1f GB/s" % bw(two))
    rt.cudaFreeHost(p)
# with a concurrent D2H of 33 MB (the e2e step's readback)
out = torch.empty(33177600, dtype=torch.uint8).pin_memory(); s3 = torch.cuda.Stream()
def both():
    dev.copy_(pin, non_blocking=True)
    with torch.cuda.stream(s3): out.copy_(dev2[:33177600], non_blocking=True)
print("torch pinned + D2H 33MB %.1f GB/s (H2D bytes only)" % bw(both))
