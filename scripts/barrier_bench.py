import ctypes as ct, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpkrylov_b200 import _lib
L = ct.CDLL(_lib.LIB_PATH)
a, b = ct.c_double(), ct.c_double()
for it in (10, 1000):
    rc = L.cpk_debug_barrier_cycles(0, it, ct.byref(a), ct.byref(b))
    print("iters", it, "rc", rc, "sync cycles", a.value, "(%.2f us @1.9GHz)" % (a.value / 1900), "reduce cycles", b.value, "(%.2f us)" % (b.value / 1900))
