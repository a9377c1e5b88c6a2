"""How fast can fresh result memory be made writable on this box?  4 KB first-touch faults vs transparent huge pages
(madvise), one thread vs several.  Decides how engine.sweep_host allocates and pre-faults the history array."""
import ctypes, mmap, os, time
import numpy as np
from concurrent.futures import ThreadPoolExecutor

libc = ctypes.CDLL("libc.so.6", use_errno=True)
MADV_HUGEPAGE = 14
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), "| defrag:",
      open("/sys/kernel/mm/transparent_hugepage/defrag").read().strip(), "| cpus", os.cpu_count())
GB = 1.5


def fresh(nbytes, huge):
    a = np.empty(nbytes // 8 + (1 << 18))
    addr = a.ctypes.data
    if huge:
        lo = (addr + (2 << 20) - 1) & ~((2 << 20) - 1)
        ln = (addr + a.nbytes - lo) & ~((2 << 20) - 1)
        rc = libc.madvise(ctypes.c_void_p(lo), ctypes.c_size_t(ln), MADV_HUGEPAGE)
        if rc != 0:
            print("madvise failed", ctypes.get_errno())
    return a


def touch(a, nthreads, stride):
    flat = a.reshape(-1)
    cut = [flat.size * j // nthreads for j in range(nthreads + 1)]
    t = time.perf_counter()
    if nthreads == 1:
        flat[::stride] = 0.0
    else:
        with ThreadPoolExecutor(nthreads) as ex:
            list(ex.map(lambda j: flat[cut[j]:cut[j + 1]:stride].fill(0.0), range(nthreads)))
    return time.perf_counter() - t


src = np.random.rand(int(GB * (1 << 30)) // 8)
for huge in (False, True):
    for nt in (1, 2, 4, 8):
        a = fresh(src.nbytes, huge)
        dt = touch(a, nt, 512)
        print("huge=%d threads=%d: touch %.1f GB in %.3f s = %.1f GB/s" % (huge, nt, GB, dt, GB / dt))
        del a
    for nt in (1, 4, 8):
        a = fresh(src.nbytes, huge)
        flat = a.reshape(-1)[:src.size]
        cut = [flat.size * j // nt for j in range(nt + 1)]
        t = time.perf_counter()
        with ThreadPoolExecutor(nt) as ex:
            list(ex.map(lambda j: np.copyto(flat[cut[j]:cut[j + 1]], src[cut[j]:cut[j + 1]]), range(nt)))
        dt = time.perf_counter() - t
        print("huge=%d threads=%d: copy into fresh memory %.3f s = %.1f GB/s" % (huge, nt, dt, GB / dt))
        t = time.perf_counter()
        with ThreadPoolExecutor(nt) as ex:
            list(ex.map(lambda j: np.copyto(flat[cut[j]:cut[j + 1]], src[cut[j]:cut[j + 1]]), range(nt)))
        dt = time.perf_counter() - t
        print("                     copy into touched memory %.3f s = %.1f GB/s" % (dt, GB / dt))
        del a
