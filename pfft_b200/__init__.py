"""pfft_b200 -- Python host layer over libpfft_b200.so (ctypes).

The product is the C library behind PFFT's C API (include/pfft.h); this module only
mirrors that API for Python callers (tests, bench.py): same function names minus the
`pfft_` prefix, same argument meaning and order, same flag values
(reference api/pfft.h:69-572).  Arrays are passed as

  * torch CUDA tensors            -> used in place (device pointer),
  * numpy arrays                  -> host pointers, staged through HBM inside execute,
  * arrays from alloc_complex/real -> CUDA managed memory (what pfft_alloc_* returns).

There is no Python or CPU implementation of any transform here: if the shared library
is missing the import fails.
"""
import ctypes as C
import json
import os
import uuid

import numpy as np

from . import _lib

INT = C.c_ssize_t

# ---- flags (reference api/pfft.h:528-572) ------------------------------------------
FORWARD, BACKWARD = -1, +1
TRANSPOSED_NONE, TRANSPOSED_IN, TRANSPOSED_OUT = 0, 1 << 0, 1 << 1
SHIFTED_NONE, SHIFTED_IN, SHIFTED_OUT = 0, 1 << 2, 1 << 3
MEASURE, ESTIMATE, PATIENT, EXHAUSTIVE = 0, 1 << 4, 1 << 5, 1 << 6
NO_TUNE, TUNE = 0, 1 << 7
PRESERVE_INPUT, DESTROY_INPUT, BUFFERED_INPLACE = 1 << 8, 1 << 9, 1 << 10
PADDED_R2C = PADDED_C2R = 1 << 11
GC_TRANSPOSED_NONE, GC_TRANSPOSED, GC_SENDRECV, GC_RMA, GC_R2C, GC_PADDED = 0, 1, 2, 4, 8, 16
R2HC, HC2R, DHT, REDFT00, REDFT01, REDFT10, REDFT11, RODFT00, RODFT01, RODFT10, RODFT11 = range(11)

_KIND_NAME = {"c2c": "dft", "r2c": "dft_r2c", "c2r": "dft_c2r", "r2r": "r2r"}


def lib():
    return _lib.load()


def _vec(v, ctype=INT):
    if v is None:
        return None
    return (ctype * len(v))(*[int(x) for x in v])


def _sym_addr(name):
    return C.addressof(C.c_char.in_dll(lib(), name))


class Comm:
    """An MPI communicator handle of the library's MPI (minimpi in this image)."""

    def __init__(self, handle, owned=False):
        self.handle = C.c_void_p(handle)
        self.owned = owned

    @property
    def rank(self):
        r = C.c_int()
        lib().MPI_Comm_rank(self.handle, C.byref(r))
        return r.value

    @property
    def size(self):
        r = C.c_int()
        lib().MPI_Comm_size(self.handle, C.byref(r))
        return r.value

    def barrier(self):
        lib().MPI_Barrier(self.handle)

    def allreduce_max(self, x):
        a, b = C.c_double(float(x)), C.c_double()
        lib().MPI_Allreduce(C.byref(a), C.byref(b), 1, 9, 1, self.handle)   # MPI_DOUBLE, MPI_MAX
        return b.value

    def free(self):
        if self.owned and self.handle:
            lib().MPI_Comm_free(C.byref(self.handle))
            self.owned = False


_initialized = False


def bootstrap(jobname, rank, size):
    """Join a multi-rank job without a launcher (call before init)."""
    return lib().minimpi_bootstrap(jobname.encode(), int(rank), int(size))


def init(use_torch_distributed=None):
    """MPI_Init + pfft_init.  Under torchrun (RANK/WORLD_SIZE set) with an initialised
    torch.distributed process group the job name is agreed through it; otherwise the
    environment prepared by pfftrun (or a single rank) is used."""
    global _initialized
    if _initialized:
        return
    L = lib()
    if use_torch_distributed is None:
        use_torch_distributed = "PFFT_MPI_JOB" not in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1
    if use_torch_distributed:
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("pfft_b200.init: initialise torch.distributed first (or launch with pfftrun)")
        name = [uuid.uuid4().hex[:16] if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(name, src=0)
        bootstrap(name[0], dist.get_rank(), dist.get_world_size())
    L.MPI_Init(None, None)
    L.pfft_init()
    _initialized = True


def finalize():
    global _initialized
    if _initialized:
        lib().MPI_Finalize()
        _initialized = False


def comm_world():
    return Comm(_sym_addr("minimpi_comm_world_obj"))


def comm_self():
    return Comm(_sym_addr("minimpi_comm_self_obj"))


def create_procmesh(np_, comm=None):
    """pfft_create_procmesh (reference kernel/procmesh.c:36-64): raises if prod(np) != comm size."""
    comm = comm or comm_world()
    out = C.c_void_p()
    rc = lib().pfft_create_procmesh(len(np_), comm.handle, _vec(np_, C.c_int), C.byref(out))
    if rc != 0:
        raise ValueError("process mesh %s does not match the communicator size %d" % (list(np_), comm.size))
    return Comm(out.value, owned=True)


def last_error():
    f = lib().pfftb200_last_error
    f.restype = C.c_char_p
    return f().decode()


def _prefix(dtype):
    dtype = np.dtype(dtype)
    if dtype in (np.dtype(np.float64), np.dtype(np.complex128)):
        return "pfft_", np.float64
    if dtype in (np.dtype(np.float32), np.dtype(np.complex64)):
        return "pfftf_", np.float32
    raise TypeError("unsupported dtype %s" % dtype)


def local_size(kind, n, comm, flags=0, ni=None, no=None, howmany=1, iblock=None, oblock=None):
    """pfft_local_size_many_{dft,dft_r2c,dft_c2r,r2r}: returns (alloc_local, local_ni,
    local_i_start, local_no, local_o_start)."""
    d = len(n)
    V = INT * d
    a, b, c_, e = V(), V(), V(), V()
    fn = getattr(lib(), "pfft_local_size_many_" + _KIND_NAME[kind])
    fn.restype = INT
    alloc = fn(C.c_int(d), _vec(n), _vec(ni if ni is not None else n), _vec(no if no is not None else n),
               INT(howmany), _vec(iblock), _vec(oblock), comm.handle, C.c_uint(flags), a, b, c_, e)
    return int(alloc), list(a), list(b), list(c_), list(e)


def local_block(kind, n, comm, pid, flags=0, ni=None, no=None, iblock=None, oblock=None):
    d = len(n)
    V = INT * d
    a, b, c_, e = V(), V(), V(), V()
    fn = getattr(lib(), "pfft_local_block_many_" + _KIND_NAME[kind])
    fn.restype = None
    fn(C.c_int(d), _vec(ni if ni is not None else n), _vec(no if no is not None else n), _vec(iblock), _vec(oblock),
       comm.handle, C.c_int(pid), C.c_uint(flags), a, b, c_, e)
    return list(a), list(b), list(c_), list(e)


class ManagedArray:
    """Memory from pfft_alloc_* (CUDA managed) viewed as a numpy array."""

    def __init__(self, count, dtype):
        self.dtype = np.dtype(dtype)
        self.count = int(count)
        f = lib().pfft_malloc
        f.restype = C.c_void_p
        self.ptr = f(C.c_size_t(max(self.count, 1) * self.dtype.itemsize))
        if not self.ptr:
            raise MemoryError("pfft_malloc failed")
        buf = (C.c_char * (max(self.count, 1) * self.dtype.itemsize)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=self.count)

    def free(self):
        if self.ptr:
            self.array = None
            lib().pfft_free(C.c_void_p(self.ptr))
            self.ptr = None


def alloc_complex(count, dtype=np.complex128):
    return ManagedArray(count, dtype)


def alloc_real(count, dtype=np.float64):
    return ManagedArray(count, dtype)


def _pointer(x):
    """(address, keepalive) of a torch tensor, numpy array, ManagedArray or int."""
    if x is None:
        return None, None
    if isinstance(x, ManagedArray):
        return x.ptr, x
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("arrays must be C-contiguous")
        return x.ctypes.data, x
    if isinstance(x, int):
        return x, None
    if hasattr(x, "data_ptr"):
        if not x.is_contiguous():
            raise ValueError("tensors must be contiguous")
        return x.data_ptr(), x
    raise TypeError("cannot take the address of %r" % type(x))


class Plan:
    """pfft_plan.  Created by plan_dft / plan_dft_r2c / plan_dft_c2r / plan_r2r."""

    def __init__(self, handle, prefix, keep):
        self.handle = C.c_void_p(handle)
        self.prefix = prefix
        self._keep = keep

    def execute(self, in_=None, out=None):
        """pfft_execute / pfft_execute_dft (new-array execute): blocking, like the reference."""
        L = lib()
        if in_ is None and out is None:
            getattr(L, self.prefix + "execute")(self.handle)
        else:
            pi, _ = _pointer(in_)
            po, _ = _pointer(out)
            getattr(L, self.prefix + "execute_dft")(self.handle, C.c_void_p(pi), C.c_void_p(po))

    def execute_async(self, in_=None, out=None):
        pi, _ = _pointer(in_)
        po, _ = _pointer(out)
        lib().pfftb200_execute_async(self.handle, C.c_void_p(pi), C.c_void_p(po))

    def describe(self):
        f = lib().pfftb200_plan_describe
        f.restype = C.c_size_t
        need = f(self.handle, None, C.c_size_t(0))
        buf = C.create_string_buffer(need + 8)
        f(self.handle, buf, C.c_size_t(need + 8))
        return json.loads(buf.value.decode())

    def stage_times_ms(self):
        out = (C.c_double * 64)()
        n = lib().pfftb200_stage_times(self.handle, out, 64)
        return list(out)[:n]

    def exchange_times_ms(self):
        out = (C.c_double * 64)()
        n = lib().pfftb200_exchange_times(self.handle, out, 64)
        return list(out)[:n]

    def enable_stage_timing(self, on):
        lib().pfftb200_enable_stage_timing(self.handle, int(bool(on)))

    def destroy(self):
        if self.handle:
            getattr(lib(), self.prefix + "destroy_plan")(self.handle)
            self.handle = None


def _plan(kind, n, in_, out, comm, sign, flags, dtype, ni, no, howmany, iblock, oblock, kinds, skip):
    prefix, _ = _prefix(dtype)
    d = len(n)
    pi, k1 = _pointer(in_)
    po, k2 = _pointer(out)
    fn = getattr(lib(), prefix + "plan_many_" + _KIND_NAME[kind] + "_skipped")
    fn.restype = C.c_void_p
    last = _vec(kinds, C.c_int) if kind == "r2r" else C.c_int(sign)
    h = fn(C.c_int(d), _vec(n), _vec(ni if ni is not None else n), _vec(no if no is not None else n), INT(howmany),
           _vec(iblock), _vec(oblock), _vec(skip, C.c_int), C.c_void_p(pi), C.c_void_p(po), comm.handle, last,
           C.c_uint(flags))
    if not h:
        return None
    return Plan(h, prefix, (k1, k2))


def plan_dft(n, in_, out, comm, sign, flags=0, dtype=np.complex128, ni=None, no=None, howmany=1, iblock=None,
             oblock=None, skip=None):
    """pfft_plan_many_dft[_skipped]; returns None where the reference returns NULL (see last_error())."""
    return _plan("c2c", n, in_, out, comm, sign, flags, dtype, ni, no, howmany, iblock, oblock, None, skip)


def plan_dft_r2c(n, in_, out, comm, sign=FORWARD, flags=0, dtype=np.float64, **kw):
    return _plan("r2c", n, in_, out, comm, sign, flags, dtype, kw.get("ni"), kw.get("no"), kw.get("howmany", 1),
                 kw.get("iblock"), kw.get("oblock"), None, kw.get("skip"))


def plan_dft_c2r(n, in_, out, comm, sign=BACKWARD, flags=0, dtype=np.float64, **kw):
    return _plan("c2r", n, in_, out, comm, sign, flags, dtype, kw.get("ni"), kw.get("no"), kw.get("howmany", 1),
                 kw.get("iblock"), kw.get("oblock"), None, kw.get("skip"))


def plan_r2r(n, in_, out, comm, kinds, flags=0, dtype=np.float64, **kw):
    return _plan("r2r", n, in_, out, comm, -1, flags, dtype, kw.get("ni"), kw.get("no"), kw.get("howmany", 1),
                 kw.get("iblock"), kw.get("oblock"), kinds, kw.get("skip"))


def init_input(kind, n, local_n, local_start, data, dtype=np.float64):
    """pfft_init_input_{complex,complex_hermitian,real}: kind in those three names."""
    prefix, _ = _prefix(dtype)
    p, _k = _pointer(data)
    getattr(lib(), prefix + "init_input_" + kind)(C.c_int(len(n)), _vec(n), _vec(local_n), _vec(local_start), C.c_void_p(p))


def clear_input(kind, n, local_n, local_start, data, dtype=np.float64):
    prefix, _ = _prefix(dtype)
    p, _k = _pointer(data)
    getattr(lib(), prefix + "clear_input_" + kind)(C.c_int(len(n)), _vec(n), _vec(local_n), _vec(local_start), C.c_void_p(p))


def check_output(kind, n, local_n, local_start, data, comm, dtype=np.float64):
    prefix, real = _prefix(dtype)
    p, _k = _pointer(data)
    fn = getattr(lib(), prefix + "check_output_" + kind)
    fn.restype = C.c_double if real == np.float64 else C.c_float
    return float(fn(C.c_int(len(n)), _vec(n), _vec(local_n), _vec(local_start), C.c_void_p(p), comm.handle))


def describe_schedule(kind, n, np_, pid, flags=0, ni=None, no=None, howmany=1, iblock=None, oblock=None, sign=-1,
                      kinds=None, skip=None):
    """Planner introspection without MPI/CUDA (pfftb200_describe_schedule)."""
    d = len(n)
    fn = lib().pfftb200_describe_schedule
    fn.restype = C.c_size_t
    args = [C.c_int({"c2c": 0, "r2c": 1, "c2r": 2, "r2r": 3}[kind]), C.c_int(d), _vec(n),
            _vec(ni if ni is not None else n), _vec(no if no is not None else n), INT(howmany), _vec(iblock),
            _vec(oblock), C.c_int(len(np_)), _vec(np_, C.c_int), C.c_int(pid), C.c_int(sign), _vec(kinds, C.c_int),
            _vec(skip, C.c_int), C.c_uint(flags)]
    need = fn(*args, None, C.c_size_t(0))
    buf = C.create_string_buffer(need + 8)
    fn(*args, buf, C.c_size_t(need + 8))
    return json.loads(buf.value.decode())


def describe_kernels(kind, n, np_, pid, flags=0, ni=None, no=None, howmany=1, iblock=None, oblock=None, sign=-1,
                     kinds=None, skip=None, precision="double"):
    """Kernel family and tile geometry of every stage of rank `pid`, without MPI/CUDA (pfftb200_describe_kernels)."""
    d = len(n)
    fn = lib().pfftb200_describe_kernels
    fn.restype = C.c_size_t
    args = [C.c_int(0 if precision == "double" else 1), C.c_int({"c2c": 0, "r2c": 1, "c2r": 2, "r2r": 3}[kind]), C.c_int(d),
            _vec(n), _vec(ni if ni is not None else n), _vec(no if no is not None else n), INT(howmany), _vec(iblock),
            _vec(oblock), C.c_int(len(np_)), _vec(np_, C.c_int), C.c_int(pid), C.c_int(sign), _vec(kinds, C.c_int),
            _vec(skip, C.c_int), C.c_uint(flags)]
    need = fn(*args, None, C.c_size_t(0))
    buf = C.create_string_buffer(need + 8)
    fn(*args, buf, C.c_size_t(need + 8))
    return json.loads(buf.value.decode())


def describe_exchange_ordering(kind, n, np_, pid, flags=0, ni=None, no=None, howmany=1, iblock=None, oblock=None, sign=-1,
                               kinds=None, skip=None):
    """Device-side ordering of the p2p transport for rank `pid`, without MPI/CUDA (pfftb200_describe_exchange_ordering)."""
    d = len(n)
    fn = lib().pfftb200_describe_exchange_ordering
    fn.restype = C.c_size_t
    args = [C.c_int({"c2c": 0, "r2c": 1, "c2r": 2, "r2r": 3}[kind]), C.c_int(d), _vec(n),
            _vec(ni if ni is not None else n), _vec(no if no is not None else n), INT(howmany), _vec(iblock),
            _vec(oblock), C.c_int(len(np_)), _vec(np_, C.c_int), C.c_int(pid), C.c_int(sign), _vec(kinds, C.c_int),
            _vec(skip, C.c_int), C.c_uint(flags)]
    need = fn(*args, None, C.c_size_t(0))
    buf = C.create_string_buffer(need + 8)
    fn(*args, buf, C.c_size_t(need + 8))
    return json.loads(buf.value.decode())


def launch_count():
    f = lib().pfftb200_launch_count
    f.restype = C.c_ulonglong
    return int(f())


def set_stream(cuda_stream_ptr):
    lib().pfftb200_set_stream(C.c_void_p(cuda_stream_ptr))


def set_transport(name):
    return lib().pfftb200_set_transport(name.encode())


def version():
    f = lib().pfftb200_version
    f.restype = C.c_char_p
    return f().decode()
