"""Development probe: H2D bandwidth from write-combined pinned memory vs default pinned memory (cudaHostAlloc flags)."""
import ctypes, time, sys
import torch
rt = ctypes.CDLL("libcudart.so.12")
n = 66 * 1024 * 1024
d = torch.empty(n, dtype=torch.uint8, device="cuda")
src = torch.randint(0, 255, (n,), dtype=torch.uint8)
for name, flags in (("default", 0), ("write-combined", 4), ("portable|mapped", 3), ("default again", 0)):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flags)) == 0
    t0 = time.perf_counter()
    ctypes.memmove(p.value, src.data_ptr(), n)
    t_fill = time.perf_counter() - t0
    st = torch.cuda.current_stream().cuda_stream
    def go():
        rc = rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), p, ctypes.c_size_t(n), ctypes.c_int(1), ctypes.c_void_p(st))
        assert rc == 0
    for _ in range(3): go()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): go()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    ok = bool((d[:1 << 20].cpu() == src[:1 << 20]).all())
    print("%-16s host fill %.1f ms (%.1f GB/s)  H2D %.2f ms, %.1f GB/s  data ok=%s" % (name, t_fill * 1e3, n / t_fill / 1e9, dt * 1e3, n / dt / 1e9, ok))
    rt.cudaFreeHost(p)
