"""What cudaHostRegister costs on touched host memory (4 KiB pages and transparent huge pages), per GB."""
import ctypes as C, mmap, time
import numpy as np, torch
torch.cuda.init()
rt = torch.cuda.cudart()
libc = C.CDLL("libc.so.6")
libc.posix_memalign.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_size_t]
libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
for huge in (False, True):
    for gb in (1, 4):
        n = gb << 30
        p = C.c_void_p()
        assert libc.posix_memalign(C.byref(p), 2 << 20, n) == 0
        if huge:
            libc.madvise(p, n, 14)  # MADV_HUGEPAGE
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n,))
        t0 = time.perf_counter(); a[::4096] = 1; touch = time.perf_counter() - t0
        t0 = time.perf_counter(); r = rt.cudaHostRegister(p.value, n, 0); reg = time.perf_counter() - t0
        t0 = time.perf_counter(); rt.cudaHostUnregister(p.value); unreg = time.perf_counter() - t0
        print(f"huge={huge} {gb} GB: touch {touch*1e3:.0f} ms, register {reg*1e3:.0f} ms (rc {int(r)}), unregister {unreg*1e3:.0f} ms", flush=True)
        libc.free(p)
