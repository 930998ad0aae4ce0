"""oracle/refint.py -- TEST INFRASTRUCTURE ONLY.

ctypes view of oracle/_ref/libpfft_refint.so, i.e. the reference's *own*
integer layer (kernel/block.c, kernel/partrafo.c:99-199, kernel/procmesh.c,
gcell/gcells_plan.c:51-76, api/api-basic.c:60-121) compiled by oracle/Makefile
against single-process stub headers.  It is used (a) by tests/golden/gen_golden.py
to capture golden vectors in the build container and (b) by the CPU tests as a
live cross-check whenever the prebuilt file is present.  Nothing in the product
imports this module.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpfft_refint.so")

TRANSPOSED_NONE, TRANSPOSED_IN, TRANSPOSED_OUT = 0, 1, 2
SHIFTED_IN, SHIFTED_OUT = 4, 8
PADDED_R2C = 1 << 11

INT = C.c_ssize_t  # ptrdiff_t


def available():
    return os.path.exists(LIB_PATH)


class RefInt:
    def __init__(self):
        self.lib = C.CDLL(LIB_PATH)
        self.lib.oracle_stub_set_world.argtypes = [C.c_int, C.c_int]
        self.world = C.c_void_p.in_dll(self.lib, "oracle_stub_world")

    def _world_ptr(self):
        return C.c_void_p(C.addressof(self.world))

    def procmesh(self, np_):
        size = 1
        for p in np_:
            size *= p
        self.lib.oracle_stub_set_world(size, 0)
        comm = C.c_void_p()
        arr = (C.c_int * len(np_))(*np_)
        rc = self.lib.pfft_create_procmesh(len(np_), self._world_ptr(), arr, C.byref(comm))
        assert rc == 0
        return comm

    def local_block(self, kind, ni, no, np_, pid, flags, iblock=None, oblock=None):
        """kind in {'dft','dft_r2c','dft_c2r','r2r'}; returns (lni, lis, lno, los)."""
        d = len(ni)
        comm = self.procmesh(np_)
        V = INT * d
        lni, lis, lno, los = V(), V(), V(), V()
        ib = (INT * len(iblock))(*iblock) if iblock is not None else None
        ob = (INT * len(oblock))(*oblock) if oblock is not None else None
        fn = getattr(self.lib, "pfft_local_block_many_" + kind)
        fn.restype = None
        fn(C.c_int(d), V(*ni), V(*no), ib, ob, comm, C.c_int(pid), C.c_uint(flags),
           lni, lis, lno, los)
        return list(lni), list(lis), list(lno), list(los)

    def local_size_gc(self, local_n, local_start, howmany, gc_below, gc_above):
        d = len(local_n)
        V = INT * d
        ngc, gcs = V(), V()
        fn = self.lib.pfft_local_size_many_gc
        fn.restype = INT
        mem = fn(C.c_int(d), V(*local_n), V(*local_start), INT(howmany),
                 V(*gc_below), V(*gc_above), ngc, gcs)
        return int(mem), list(ngc), list(gcs)

    def init_input(self, kind, n, local_n, local_start):
        """kind in {'complex','complex_hermitian','real'} -> list of python numbers."""
        import numpy as np
        d = len(n)
        V = INT * d
        tot = 1
        for x in local_n:
            tot *= x
        if kind == "real":
            buf = np.zeros(tot, dtype=np.float64)
        else:
            buf = np.zeros(tot, dtype=np.complex128)
        fn = getattr(self.lib, "pfft_init_input_" + kind)
        fn.restype = None
        fn(C.c_int(d), V(*n), V(*local_n), V(*local_start), buf.ctypes.data_as(C.c_void_p))
        return buf
