"""ctypes binding of libp3dfft.3.so -- the B200-native P3DFFT++ transform path.

This is plumbing for tests and bench.py: the product is the C/C++ library (include/p3dfft.h,
include/Cwrap.h).  The functions below call the reference-compatible C ABI (reference
build/wrap.C:88-794) one to one; arrays are passed as raw pointers, either host (numpy) or device
(torch ``data_ptr()``).  There is no Python or CPU implementation of the transforms here: if the
shared library has not been built, loading raises.

The directory name contains a dot, so import it through ``__graft_entry__.load_package()``.
"""
import ctypes
import json
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("P3DFFT_B200_LIB") or os.path.join(_HERE, "lib", "libp3dfft.3.so")  # (override: A/B builds of the same library)
EMU_LIB_PATH = os.environ.get("P3DFFT_B200_EMU_LIB") or os.path.join(ROOT, "tools", "cuda_emu", "_build", "libp3dfft_emu.so")


class Grid(ctypes.Structure):
    """CDataGrid_struct (include/Cwrap.h; reference include/Cwrap.h:82-95)."""

    _fields_ = [
        ("nd", c_int),
        ("Gdims", c_int * 3),
        ("dim_conj_sym", c_int),
        ("MemOrder", c_int * 3),
        ("Ldims", c_int * 3),
        ("Dmap", c_int * 3),
        ("pgrid", c_int),
        ("grid_id", c_int * 3),
        ("GlobStart", c_int * 3),
        ("taskid", c_int),
        ("numtasks", c_int),
        ("ProcDims", c_int * 3),
        ("mpi_comm_glob", c_int),
    ]


TYPE_NAMES = [
    "EMPTY_TYPE_SINGLE", "EMPTY_TYPE_DOUBLE", "EMPTY_TYPE_SINGLE_COMPLEX", "EMPTY_TYPE_DOUBLE_COMPLEX",
    "R2CFFT_S", "R2CFFT_D", "C2RFFT_S", "C2RFFT_D",
    "CFFT_FORWARD_S", "CFFT_FORWARD_D", "CFFT_BACKWARD_S", "CFFT_BACKWARD_D",
] + [f"{k}{n}_{v}" for n in (1, 2, 3, 4) for k in ("DCT", "DST") for v in ("REAL_S", "REAL_D", "COMPLEX_S", "COMPLEX_D")]

C_SYMBOLS = [
    "p3dfft_setup", "p3dfft_cleanup", "p3dfft_init_3Dtype", "p3dfft_plan_1Dtrans", "p3dfft_plan_3Dtrans", "find_grid",
    "p3dfft_init_proc_grid", "p3dfft_init_data_grid", "p3dfft_free_data_grid", "p3dfft_free_proc_grid", "p3dfft_inv_mo",
    "p3dfft_write_buf", "p3dfft_exec_1Dtrans_double", "p3dfft_exec_1Dtrans_single", "p3dfft_exec_3Dtrans_double",
    "p3dfft_exec_3Dtrans_single", "p3dfft_exec_3Dderiv_double", "p3dfft_exec_3Dderiv_single",
    "p3dfft_compute_deriv_single", "p3dfft_compute_deriv_double",
    # Fortran twins (include/Fwrap.h)
    "p3dfft_init_3Dtype_f", "p3dfft_plan_1Dtrans_f", "p3dfft_plan_3Dtrans_f", "p3dfft_init_proc_grid_f",
    "p3dfft_init_data_grid_f", "p3dfft_exec_1Dtrans_double_f", "p3dfft_exec_1Dtrans_single_f",
    "p3dfft_exec_3Dtrans_double_f", "p3dfft_exec_3Dtrans_single_f", "p3dfft_exec_3Dderiv_double_f",
    "p3dfft_exec_3Dderiv_single_f", "p3dfft_compute_deriv_single_f", "p3dfft_compute_deriv_double_f",
]
EXT_SYMBOLS = [
    "p3dfft_b200_version", "p3dfft_b200_set_stream", "p3dfft_b200_sync", "p3dfft_b200_kernel_launches",
    "p3dfft_b200_describe_plan3d", "p3dfft_b200_describe_plan1d", "p3dfft_b200_enable_timers",
    "p3dfft_b200_stage_times", "p3dfft_b200_have_device",
]
CU_SYMBOLS = [
    "p3dfftcu_last_error", "p3dfftcu_device_count", "p3dfftcu_init", "p3dfftcu_malloc", "p3dfftcu_free",
    "p3dfftcu_memset", "p3dfftcu_memcpy", "p3dfftcu_stream_sync", "p3dfftcu_pointer_is_device",
    "p3dfftcu_stage_create", "p3dfftcu_stage_destroy", "p3dfftcu_stage_exec", "p3dfftcu_stage_exec_capped", "p3dfftcu_stage_variant",
    "p3dfftcu_stream_create", "p3dfftcu_stream_destroy", "p3dfftcu_stream_wait_event", "p3dfftcu_num_sms",
    "p3dfftcu_deriv", "p3dfftcu_event_create", "p3dfftcu_event_destroy", "p3dfftcu_event_record",
    "p3dfftcu_event_elapsed", "p3dfftcu_ipc_export", "p3dfftcu_ipc_open", "p3dfftcu_ipc_close",
    "p3dfftcu_peer_barrier", "p3dfftcu_launch_count",
]


def _ptr(a):
    """raw address of a numpy array, a torch tensor, an int or None"""
    if a is None:
        return c_void_p(0)
    if isinstance(a, int):
        return c_void_p(a)
    if hasattr(a, "data_ptr"):
        return c_void_p(a.data_ptr())
    return c_void_p(a.ctypes.data)


def _i3(v):
    return (c_int * 3)(*[int(x) for x in v])


class Library:
    """One loaded copy of the shared library.  ``emulated=True`` loads the CPU-thread emulation build
    (tools/cuda_emu), which exists only to validate kernel index arithmetic without a GPU."""

    def __init__(self, path=None, emulated=False):
        self.emulated = emulated
        path = path or (EMU_LIB_PATH if emulated else LIB_PATH)
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "There is no fallback implementation.")
        self.path = path
        self.dll = ctypes.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_DEEPBIND)  # product and emulation builds define the same symbols
        d = self.dll
        d.p3dfft_init_data_grid.restype = POINTER(Grid)
        d.p3dfft_b200_version.restype = c_char_p
        d.p3dfft_b200_kernel_launches.restype = c_longlong
        d.p3dfft_b200_describe_plan3d.restype = c_size_t
        d.p3dfft_b200_describe_plan1d.restype = c_size_t
        d.p3dfftcu_last_error.restype = c_char_p
        d.p3dfftcu_stage_variant.restype = c_char_p
        d.p3dfftcu_launch_count.restype = c_longlong
        self._setup_done = False

    # ---- reference C API
    def setup(self):
        self.dll.p3dfft_setup()
        self._setup_done = True
        self.types = {n: c_int.in_dll(self.dll, "P3DFFT_" + n).value for n in TYPE_NAMES}
        return self

    def cleanup(self):
        self.dll.p3dfft_cleanup()
        self._setup_done = False

    def have_device(self):
        return bool(self.dll.p3dfft_b200_have_device())

    def init_proc_grid(self, pdims, comm=0):
        return self.dll.p3dfft_init_proc_grid(_i3(pdims), c_int(comm))

    def init_data_grid(self, gdims, dim_conj_sym, pgrid, dmap, mem_order):
        return self.dll.p3dfft_init_data_grid(_i3(gdims), c_int(dim_conj_sym), c_int(pgrid), _i3(dmap), _i3(mem_order))

    def free_data_grid(self, g):
        self.dll.p3dfft_free_data_grid(g)

    def init_3Dtype(self, types):
        ids = [self.types[t] if isinstance(t, str) else int(t) for t in types]
        return self.dll.p3dfft_init_3Dtype(_i3(ids))

    def plan_3Dtrans(self, g1, g2, type3d):
        return self.dll.p3dfft_plan_3Dtrans(g1, g2, c_int(type3d))

    def plan_1Dtrans(self, g1, g2, type_id, dim):
        tid = self.types[type_id] if isinstance(type_id, str) else int(type_id)
        return self.dll.p3dfft_plan_1Dtrans(g1, g2, c_int(tid), c_int(dim))

    def exec_3Dtrans(self, plan, a_in, a_out, ow=0, single=False):
        f = self.dll.p3dfft_exec_3Dtrans_single if single else self.dll.p3dfft_exec_3Dtrans_double
        f(c_int(plan), _ptr(a_in), _ptr(a_out), c_int(ow))

    def exec_3Dderiv(self, plan, a_in, a_out, idir, ow=0, single=False):
        f = self.dll.p3dfft_exec_3Dderiv_single if single else self.dll.p3dfft_exec_3Dderiv_double
        f(c_int(plan), _ptr(a_in), _ptr(a_out), c_int(idir), c_int(ow))

    def exec_1Dtrans(self, plan, a_in, a_out, ow=0, single=False):
        f = self.dll.p3dfft_exec_1Dtrans_single if single else self.dll.p3dfft_exec_1Dtrans_double
        f(c_int(plan), _ptr(a_in), _ptr(a_out), c_int(ow))

    def compute_deriv(self, a_in, a_out, grid, idir, single=False):
        f = self.dll.p3dfft_compute_deriv_single if single else self.dll.p3dfft_compute_deriv_double
        f(_ptr(a_in), _ptr(a_out), grid, c_int(idir))

    # ---- extensions
    def version(self):
        return self.dll.p3dfft_b200_version().decode()

    def set_stream(self, stream_handle):
        self.dll.p3dfft_b200_set_stream(c_void_p(stream_handle))

    def sync(self):
        self.dll.p3dfft_b200_sync()

    def kernel_launches(self):
        return int(self.dll.p3dfft_b200_kernel_launches())

    def describe_plan3d(self, plan):
        n = self.dll.p3dfft_b200_describe_plan3d(c_int(plan), None, c_size_t(0))
        buf = ctypes.create_string_buffer(n + 8)
        self.dll.p3dfft_b200_describe_plan3d(c_int(plan), buf, c_size_t(n + 8))
        return json.loads(buf.value.decode())

    def describe_plan1d(self, plan):
        n = self.dll.p3dfft_b200_describe_plan1d(c_int(plan), None, c_size_t(0))
        buf = ctypes.create_string_buffer(n + 8)
        self.dll.p3dfft_b200_describe_plan1d(c_int(plan), buf, c_size_t(n + 8))
        return json.loads(buf.value.decode())

    def enable_timers(self, on=True):
        self.dll.p3dfft_b200_enable_timers(c_int(1 if on else 0))

    def stage_times(self, plan, n=16):
        arr = (c_float * n)()
        k = self.dll.p3dfft_b200_stage_times(c_int(plan), arr, c_int(n))
        return [arr[i] for i in range(min(k, n))]


_cache = {}


def load(emulated=False):
    """the process-wide Library object (built library required; raises otherwise)"""
    if emulated not in _cache:
        _cache[emulated] = Library(emulated=emulated)
    return _cache[emulated]
