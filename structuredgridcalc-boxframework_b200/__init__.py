"""structuredgridcalc-boxframework_b200 — Python binding over the C ABI (include/sgc_b200.h).

The product is the CUDA library `libsgc_b200.so` next to this file; this module only loads it with
ctypes and mirrors the reference's class names (Box / DisjointBoxLayout / LevelData / Copier) so
that tests and bench.py read like the reference's own drivers.  There is no CPU fallback: creating
a LevelData without a usable CUDA device raises SgcError.

(The directory name has a hyphen, so import it through `_loadpkg.load_package()` at the repo root.)
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsgc_b200.so")

TrimCenter, TrimFace, TrimEdge, TrimCorner = 1, 2, 4, 8
PeriodicX, PeriodicY, PeriodicZ = 1, 2, 4
SGC_FMA = 1
SGC_WAVE_DIRSUM = 2

# every symbol include/sgc_b200.h declares (checked by tests/test_cabi_symbols.py)
SYMBOLS = [
    "sgc_init", "sgc_shutdown", "sgc_abi_version", "sgc_last_error", "sgc_device_info", "sgc_sync",
    "sgc_launch_count", "sgc_profile_enable", "sgc_profile_read", "sgc_event_create", "sgc_event_destroy", "sgc_event_record", "sgc_event_elapsed_ms",
    "sgc_host_alloc", "sgc_host_free", "sgc_layout_define", "sgc_copier_define", "sgc_level_create",
    "sgc_level_destroy", "sgc_level_elems", "sgc_level_fab_offset", "sgc_level_device_ptr", "sgc_level_upload",
    "sgc_level_download", "sgc_level_setval", "sgc_level_copy", "sgc_fab_copy", "sgc_fab_linear_out", "sgc_fab_linear_in",
    "sgc_plan_create_exchange", "sgc_plan_create_items", "sgc_plan_destroy", "sgc_plan_num_items",
    "sgc_plan_num_remote_items", "sgc_plan_num_cells", "sgc_exchange", "sgc_exchange_begin", "sgc_exchange_end",
    "sgc_plan_create_bc_zero_gradient", "sgc_heat_sweep", "sgc_heat_sweep_range", "sgc_heat_sweep_ranges", "sgc_laplacian_apply",
    "sgc_wave_step", "sgc_wave_run", "sgc_heat_run", "sgc_step_loop_class", "sgc_step_loop_pillars", "sgc_plan_pillars", "sgc_comm_unique_id", "sgc_comm_init", "sgc_comm_finalize", "sgc_comm_nproc",
    "sgc_comm_rank", "sgc_halo_describe", "sgc_stream_create", "sgc_stream_destroy", "sgc_stream_sync", "sgc_event_record_stream", "sgc_stream_wait_event",
    "sgc_level_upload_async", "sgc_level_download_async", "sgc_copyset_create", "sgc_copyset_run", "sgc_copyset_destroy",
    "sgc_copyset_num_cells", "sgc_level_valid_elems", "sgc_level_download_valid", "sgc_level_download_valid_async", "sgc_lb_collide", "sgc_lb_macroscopic", "sgc_lb_copy_ops", "sgc_lb_run", "sgc_lb_selftest_divconst",
    "sgc_level_edge_strips", "sgc_plan_rebox_factors", "sgc_plan_create_exchange_ex", "sgc_plan_matches_level", "sgc_stencil7_apply", "sgc_level_upload_valid", "sgc_level_upload_valid_async",
]


class SgcError(RuntimeError):
    pass


class MotionItem(C.Structure):
    """sgc_motion_item (one Motion2Way, Copier.H:37-160)."""
    _fields_ = [("recv_box", C.c_int), ("send_box", C.c_int),
                ("recv_lo", C.c_int * 3), ("recv_hi", C.c_int * 3),
                ("src_lo", C.c_int * 3), ("src_hi", C.c_int * 3),
                ("dir", C.c_int * 3), ("recv_proc", C.c_int), ("send_proc", C.c_int),
                ("tag_send", C.c_int), ("tag_recv", C.c_int), ("comp_flags", C.c_uint),
                ("send_lo", C.c_int * 3), ("send_hi", C.c_int * 3)]


COPY_OP_DTYPE = np.dtype([("dst_box", "i4"), ("dst_lo", "i4", 3), ("dst_hi", "i4", 3), ("dst_comp", "i4"),
                          ("src_box", "i4"), ("src_lo", "i4", 3), ("src_comp", "i4"), ("ncomp", "i4"),
                          ("comp_flags", "u4")])     # sgc_copy_op

ITEM_DTYPE = np.dtype([("recv_box", "i4"), ("send_box", "i4"), ("recv_lo", "i4", 3), ("recv_hi", "i4", 3),
                       ("src_lo", "i4", 3), ("src_hi", "i4", 3), ("dir", "i4", 3), ("recv_proc", "i4"),
                       ("send_proc", "i4"), ("tag_send", "i4"), ("tag_recv", "i4"), ("comp_flags", "u4"),
                       ("send_lo", "i4", 3), ("send_hi", "i4", 3)])
assert ITEM_DTYPE.itemsize == C.sizeof(MotionItem)

_i3 = C.c_int * 3
_DT = {8: np.float64, 4: np.int32, 1: np.uint8}


def _v3(v):
    return _i3(int(v[0]), int(v[1]), int(v[2]))


_lib = None


def lib():
    """Load libsgc_b200.so (building it first if the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SgcError("%s is missing: run `python structuredgridcalc-boxframework_b200/build.py` "
                       "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.sgc_last_error.restype = C.c_char_p
    L.sgc_launch_count.restype = C.c_longlong
    L.sgc_level_elems.restype = C.c_longlong
    L.sgc_level_elems.argtypes = [C.c_void_p]
    L.sgc_level_fab_offset.restype = C.c_longlong
    L.sgc_level_fab_offset.argtypes = [C.c_void_p, C.c_int]
    L.sgc_level_device_ptr.restype = C.c_void_p
    L.sgc_level_device_ptr.argtypes = [C.c_void_p, C.c_int]
    L.sgc_level_edge_strips.argtypes = [C.c_void_p]
    L.sgc_plan_rebox_factors.argtypes = [C.c_void_p, _i3]
    L.sgc_plan_num_cells.restype = C.c_longlong
    L.sgc_plan_num_cells.argtypes = [C.c_void_p]
    L.sgc_device_info.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
    L.sgc_layout_define.argtypes = [_i3, _i3, _i3, C.c_int, C.c_void_p, C.c_void_p, C.c_int, _i3]
    L.sgc_copier_define.argtypes = [_i3, _i3, _i3, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.sgc_step_loop_class.argtypes = [_i3, _i3, _i3, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int]
    L.sgc_step_loop_pillars.argtypes = [_i3, _i3, _i3, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int * 2]
    L.sgc_plan_pillars.argtypes = [C.c_void_p, C.c_int * 2]
    L.sgc_halo_describe.argtypes = [_i3, _i3, _i3, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(C.c_int), C.c_void_p, C.c_int]
    L.sgc_level_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.sgc_level_destroy.argtypes = [C.c_void_p]
    L.sgc_level_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.sgc_level_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.sgc_level_setval.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.sgc_level_copy.argtypes = [C.c_void_p, C.c_void_p]
    L.sgc_fab_copy.argtypes = [C.c_void_p, C.c_int, _i3, _i3, C.c_int, C.c_void_p, C.c_int, _i3, _i3, C.c_int, C.c_int,
                               C.c_uint]
    L.sgc_fab_linear_out.argtypes = [C.c_void_p, C.c_int, _i3, _i3, C.c_int, C.c_int, C.c_uint, C.c_void_p]
    L.sgc_fab_linear_in.argtypes = [C.c_void_p, C.c_int, _i3, _i3, C.c_int, C.c_int, C.c_uint, C.c_void_p]
    L.sgc_plan_create_exchange.argtypes = [C.c_void_p, _i3, _i3, _i3, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.POINTER(C.c_void_p)]
    L.sgc_plan_create_exchange_ex.argtypes = [C.c_void_p, _i3, _i3, _i3, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.sgc_plan_matches_level.argtypes = [C.c_void_p, C.c_void_p]
    L.sgc_plan_create_items.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.sgc_plan_create_bc_zero_gradient.argtypes = [C.c_void_p, _i3, _i3, C.POINTER(C.c_void_p)]
    for f in ("sgc_plan_destroy", "sgc_plan_num_items", "sgc_plan_num_remote_items"):
        getattr(L, f).argtypes = [C.c_void_p]
    for f in ("sgc_exchange", "sgc_exchange_begin", "sgc_exchange_end"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_void_p]
    L.sgc_heat_sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint]
    L.sgc_heat_sweep_range.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint, C.c_int, C.c_int]
    L.sgc_heat_sweep_ranges.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int]
    L.sgc_stencil7_apply.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_uint]
    L.sgc_laplacian_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_uint]
    L.sgc_wave_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_uint]
    L.sgc_wave_run.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_double, C.c_uint, C.c_int,
                               C.POINTER(C.c_int)]
    L.sgc_heat_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_uint, C.c_int, C.POINTER(C.c_int)]
    L.sgc_profile_read.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.sgc_event_create.argtypes = [C.POINTER(C.c_void_p)]
    L.sgc_event_destroy.argtypes = [C.c_void_p]
    L.sgc_event_record.argtypes = [C.c_void_p]
    L.sgc_event_elapsed_ms.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    L.sgc_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.sgc_host_free.argtypes = [C.c_void_p]
    L.sgc_stream_create.argtypes = [C.POINTER(C.c_void_p)]
    L.sgc_stream_destroy.argtypes = [C.c_void_p]
    L.sgc_event_record_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.sgc_stream_wait_event.argtypes = [C.c_void_p, C.c_void_p]
    L.sgc_stream_sync.argtypes = [C.c_void_p]
    L.sgc_level_upload_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.sgc_level_download_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sgc_level_valid_elems.argtypes = [C.c_void_p]
    L.sgc_level_valid_elems.restype = C.c_longlong
    L.sgc_level_download_valid.argtypes = [C.c_void_p, C.c_void_p]
    L.sgc_level_upload_valid.argtypes = [C.c_void_p, C.c_void_p]
    L.sgc_level_upload_valid_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.sgc_level_download_valid_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sgc_copyset_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.sgc_copyset_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.sgc_copyset_destroy.argtypes = [C.c_void_p]
    L.sgc_copyset_num_cells.argtypes = [C.c_void_p]
    L.sgc_copyset_num_cells.restype = C.c_longlong
    L.sgc_lb_collide.argtypes = [C.c_void_p, C.c_void_p]
    L.sgc_lb_macroscopic.argtypes = [C.c_void_p, C.c_void_p]
    L.sgc_lb_copy_ops.argtypes = [C.c_void_p, _i3, _i3, C.c_int, C.c_void_p, C.c_int]
    L.sgc_lb_selftest_divconst.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.POINTER(C.c_double)]
    L.sgc_lb_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                             C.POINTER(C.c_int)]
    L.sgc_comm_unique_id.argtypes = [C.c_char_p]
    L.sgc_comm_init.argtypes = [C.c_int, C.c_int, C.c_char_p]
    _lib = L
    return L


def check(status):
    if status < 0:
        raise SgcError("sgc error %d: %s" % (status, lib().sgc_last_error().decode()))
    return status


def init(device=0):
    check(lib().sgc_init(device))


def sync():
    check(lib().sgc_sync())


def launch_count():
    return int(lib().sgc_launch_count())


def profile_enable(on=True):
    check(lib().sgc_profile_enable(1 if on else 0))


def profile_read(klass):
    """(launches, total_ms) of class 0 = sweep kernels, 1 = exchange copy kernels, 2 = two-step sweep passes; clears the log."""
    n, ms = C.c_int(), C.c_double()
    check(lib().sgc_profile_read(klass, C.byref(n), C.byref(ms)))
    return n.value, ms.value


def device_info():
    name = C.create_string_buffer(128)
    sms, mem = C.c_int(), C.c_size_t()
    check(lib().sgc_device_info(name, 128, C.byref(sms), C.byref(mem)))
    return name.value.decode(), sms.value, mem.value


class Event:
    def __init__(self):
        self.h = C.c_void_p()
        check(lib().sgc_event_create(C.byref(self.h)))

    def record(self, stream=None):
        """mark the current end of `stream` (a Stream; None = the compute stream)"""
        if stream is None:
            check(lib().sgc_event_record(self.h))
        else:
            check(lib().sgc_event_record_stream(self.h, stream.h))
        return self

    def elapsed_ms(self, stop):
        ms = C.c_float()
        check(lib().sgc_event_elapsed_ms(self.h, stop.h, C.byref(ms)))
        return ms.value

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.sgc_event_destroy(self.h)
            self.h = None


class Stream:
    """A transfer stream (LevelData::copyToDeviceAsync / copyToHostAsync take a cudaStream_t,
    LevelData.H:734-765).  `COMPUTE` (None) names the library's compute stream in wait()."""
    COMPUTE = None

    def __init__(self):
        self.h = C.c_void_p()
        check(lib().sgc_stream_create(C.byref(self.h)))

    def wait(self, event):
        """work submitted to this stream from now on starts after `event`"""
        check(lib().sgc_stream_wait_event(self.h, event.h))
        return self

    def sync(self):
        check(lib().sgc_stream_sync(self.h))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.sgc_stream_destroy(self.h)
            self.h = None


def compute_wait(event):
    """work submitted to the compute stream from now on starts after `event`"""
    check(lib().sgc_stream_wait_event(None, event.h))


class PinnedBuffer:
    """cudaMallocHost memory with a numpy view (`.array`) — the reference's host mirror under
    USE_GPU (BaseFab.cpp:515-519)."""

    def __init__(self, n, dtype=np.float64):
        self.ptr = C.c_void_p()
        nbytes = max(int(n) * np.dtype(dtype).itemsize, 1)
        check(lib().sgc_host_alloc(nbytes, C.byref(self.ptr)))
        self._buf = (C.c_char * nbytes).from_address(self.ptr.value)
        self.array = np.frombuffer(self._buf, dtype=dtype, count=int(n))

    def free(self):
        if self.ptr and _lib is not None:
            self.array = None
            self._buf = None
            _lib.sgc_host_free(self.ptr)
        self.ptr = None

    def __del__(self):
        self.free()


class Box:
    """Closed index box [lo, hi] (Box.H:56-228)."""

    def __init__(self, lo, hi):
        self.lo = tuple(int(x) for x in lo)
        self.hi = tuple(int(x) for x in hi)

    def grow(self, n):
        return Box([l - n for l in self.lo], [h + n for h in self.hi])

    def size(self):
        return int(np.prod(self.dimensions()))

    def dimensions(self):
        return tuple(h - l + 1 for l, h in zip(self.lo, self.hi))

    def __repr__(self):
        return "Box(%s, %s)" % (self.lo, self.hi)


class DisjointBoxLayout:
    """Uniform decomposition of `domain` into boxes of `maxbox`, block-assigned to `nproc` ranks
    (DisjointBoxLayout.H:38-185).  `rank` is this process' rank (== GPU)."""

    def __init__(self, domain, maxbox, nproc=1, rank=0):
        self.domain = domain
        self.maxbox = tuple(int(m) for m in (maxbox if hasattr(maxbox, "__len__") else (maxbox,) * 3))
        self.nproc, self.rank = int(nproc), int(rank)
        nb = _i3()
        n = check(lib().sgc_layout_define(_v3(domain.lo), _v3(domain.hi), _v3(self.maxbox), self.nproc, None, None, 0, nb))
        self.boxes = np.zeros((n, 6), np.int32)
        self.owner = np.zeros(n, np.int32)
        check(lib().sgc_layout_define(_v3(domain.lo), _v3(domain.hi), _v3(self.maxbox), self.nproc,
                                      self.boxes.ctypes.data, self.owner.ctypes.data, n, nb))
        self.numbox = tuple(nb)
        bpp = n // self.nproc
        self._beg, self._end = self.rank * bpp, (self.rank + 1) * bpp

    def size(self):
        return len(self.boxes)

    def localSize(self):
        return self._end - self._beg

    def localIdxBegin(self):
        return self._beg

    def localIdxEnd(self):
        return self._end

    def local_boxes(self):
        return self.boxes[self._beg:self._end]

    def dimensions(self):
        return self.numbox

    def copier_items(self, nghost, periodic=0, trim=0, rank=None):
        """Copier::defineExchangeDBL on the host (sgc_copier_define) -> structured numpy array."""
        r = self.rank if rank is None else rank
        args = (_v3(self.domain.lo), _v3(self.domain.hi), _v3(self.maxbox), nghost, periodic, trim, self.nproc, r)
        n = check(lib().sgc_copier_define(*args, None, 0))
        items = np.zeros(max(n, 1), ITEM_DTYPE)
        check(lib().sgc_copier_define(*args, items.ctypes.data, n))
        return items[:n]

    def halo_lists(self, nghost, periodic=0, trim=0, rank=None):
        """[(peer, recv_items, send_items)] — aggregated halo messages (sgc_halo_describe)."""
        r = self.rank if rank is None else rank
        base = (_v3(self.domain.lo), _v3(self.domain.hi), _v3(self.maxbox), nghost, periodic, trim, self.nproc, r)
        out, k = [], 0
        while True:
            peer = C.c_int(-1)
            n = lib().sgc_halo_describe(*base, k, 0, C.byref(peer), None, 0)
            if n < 0:
                break
            lists = []
            for kind in (0, 1):
                n = check(lib().sgc_halo_describe(*base, k, kind, C.byref(peer), None, 0))
                items = np.zeros(max(n, 1), ITEM_DTYPE)
                check(lib().sgc_halo_describe(*base, k, kind, C.byref(peer), items.ctypes.data, n))
                lists.append(items[:n])
            out.append((peer.value, lists[0], lists[1]))
            k += 1
        return out


class LevelData:
    """Device-resident LevelData<BaseFab<T>> (LevelData.H:40-155): one fab per local box of `dbl`
    on grow(box, nghost), all in one HBM arena.  Host arrays are the fabs concatenated in box order,
    reference memory order."""

    def __init__(self, dbl, ncomp, nghost, elem_bytes=8, boxes=None, dtype=None):
        """dtype: host element type when it is not the default of the size (np.float32: a level of float,
        the reference's USE_SINGLE_PRECISION Real)"""
        if dtype is not None:
            elem_bytes = np.dtype(dtype).itemsize
        self.dbl, self.ncomp, self.nghost, self.elem_bytes = dbl, int(ncomp), int(nghost), int(elem_bytes)
        self.dtype = np.dtype(dtype).type if dtype is not None else _DT[self.elem_bytes]
        if boxes is None:
            boxes, gidx0 = dbl.local_boxes(), dbl.localIdxBegin()
        else:
            gidx0 = 0
        self.boxes = np.ascontiguousarray(boxes, np.int32)
        self.h = C.c_void_p()
        check(lib().sgc_level_create(self.boxes.ctypes.data, len(self.boxes), self.ncomp, self.nghost, self.elem_bytes,
                                     gidx0, C.byref(self.h)))
        self.elems = int(lib().sgc_level_elems(self.h))

    def size(self):
        return len(self.boxes)

    def fab_offset(self, box):
        return int(lib().sgc_level_fab_offset(self.h, box))

    def fab_elems(self, box):
        b = self.boxes[box]
        return int(np.prod(b[3:] - b[:3] + 1 + 2 * self.nghost)) * self.ncomp

    def upload(self, host, box=-1, wait=True):
        """H2D.  wait=False leaves the copy in flight (pinned source only; keep it alive)."""
        host = np.ascontiguousarray(host, self.dtype)
        assert host.size == (self.elems if box < 0 else self.fab_elems(box))
        check(lib().sgc_level_upload(self.h, box, host.ctypes.data))
        if wait:
            sync()
        return self

    def download(self, box=-1, out=None):
        n = self.elems if box < 0 else self.fab_elems(box)
        if out is None:
            out = np.empty(n, self.dtype)
        check(lib().sgc_level_download(self.h, box, out.ctypes.data))
        return out

    def uploadAsync(self, pinned, stream):
        """whole-level H2D on a transfer stream; `pinned` is a PinnedBuffer that must outlive the copy"""
        assert pinned.array.size == self.elems and pinned.array.dtype == self.dtype
        check(lib().sgc_level_upload_async(self.h, pinned.ptr, stream.h))
        return self

    def downloadAsync(self, pinned, stream, gathered=None):
        """`gathered` (an Event) is recorded once the level has been read: waiting for it is enough
        before the level is overwritten, the D2H copy may still be in flight"""
        assert pinned.array.size == self.elems and pinned.array.dtype == self.dtype
        check(lib().sgc_level_download_async(self.h, pinned.ptr, stream.h, gathered.h if gathered else None))
        return self

    def validElems(self):
        return int(lib().sgc_level_valid_elems(self.h))

    def downloadValid(self, out=None, stream=None, gathered=None):
        """plot gather: [box][comp] valid cells, ghosts stripped (LevelData.cpp:150-193).  With `stream`
        (and `out` a PinnedBuffer) the gather + D2H run on that transfer stream and the call returns."""
        if stream is not None:
            assert out.array.size == self.validElems()
            check(lib().sgc_level_download_valid_async(self.h, out.ptr, stream.h, gathered.h if gathered else None))
            return out
        if out is None:
            out = np.empty(self.validElems(), self.dtype)
        check(lib().sgc_level_download_valid(self.h, out.ctypes.data))
        return out

    def uploadValid(self, host, stream=None):
        """valid cells only, dense [box][comp][k][j][i]; ghost cells keep their contents.  With `stream`, `host` is a
        PinnedBuffer and the copy runs on that transfer stream."""
        if stream is not None:
            assert host.array.size == self.validElems()
            check(lib().sgc_level_upload_valid_async(self.h, host.ptr, stream.h))
            return self
        host = np.ascontiguousarray(host, self.dtype)
        assert host.size == self.validElems()
        check(lib().sgc_level_upload_valid(self.h, host.ctypes.data))
        sync()
        return self

    def edgeStrips(self):
        """-1 none, 0 stale, 1 outer columns current, 2 all four (sgc_level_edge_strips)"""
        return int(lib().sgc_level_edge_strips(self.h))

    def setVal(self, value, comp=-1, box=-1):
        v = np.array([value], self.dtype)
        check(lib().sgc_level_setval(self.h, box, comp, v.ctypes.data))

    def copyFrom(self, src):
        """whole-level device copy (same geometry), ghosts included"""
        check(lib().sgc_level_copy(self.h, src.h))

    def exchange(self, copier):
        check(lib().sgc_exchange(self.h, copier.h))

    def exchangeBegin(self, copier):
        check(lib().sgc_exchange_begin(self.h, copier.h))

    def exchangeEnd(self, copier):
        check(lib().sgc_exchange_end(self.h, copier.h))

    def copy(self, dst_box, dlo, dhi, dcomp, src, src_box, slo, shi, scomp, ncomp, flags=0xFFFFFFFF):
        """BaseFab::copy between device fabs (BaseFab.cpp:296-329)."""
        check(lib().sgc_fab_copy(self.h, dst_box, _v3(dlo), _v3(dhi), dcomp, src.h, src_box, _v3(slo), _v3(shi), scomp,
                                 ncomp, flags))

    def linearOut(self, box, rlo, rhi, c0, c1, flags=0xFFFFFFFF):
        ncopy = sum(1 for c in range(c0, c1) if c >= 32 or (flags >> c) & 1)
        n = int(np.prod(np.array(rhi) - np.array(rlo) + 1)) * ncopy
        buf = np.empty(max(n, 1), self.dtype)
        check(lib().sgc_fab_linear_out(self.h, box, _v3(rlo), _v3(rhi), c0, c1, flags, buf.ctypes.data))
        return buf[:n]

    def linearIn(self, box, rlo, rhi, c0, c1, buffer, flags=0xFFFFFFFF):
        buf = np.ascontiguousarray(buffer, self.dtype)
        check(lib().sgc_fab_linear_in(self.h, box, _v3(rlo), _v3(rhi), c0, c1, flags, buf.ctypes.data))

    def destroy(self):
        if self.h and _lib is not None:
            _lib.sgc_level_destroy(self.h)
            self.h = None

    def __del__(self):
        self.destroy()


class Copier:
    """Device-resident exchange plan (Copier.H:169-280)."""

    def __init__(self):
        self.h = None

    def defineExchangeLD(self, level, periodic=0, trim=0, comp_begin=0, ncomp=None, nghost=0, item_flags=None):
        """nghost: ghost width to fill (0 = the level's); item_flags: one component mask per motion item
        (Motion2Way::setCompRecvFlags), local or remote"""
        self._release()
        dbl = level.dbl
        h = C.c_void_p()
        nc = level.ncomp if ncomp is None else ncomp
        if nghost == 0 and item_flags is None:
            check(lib().sgc_plan_create_exchange(level.h, _v3(dbl.domain.lo), _v3(dbl.domain.hi), _v3(dbl.maxbox), periodic,
                                                 trim, comp_begin, nc, dbl.nproc, dbl.rank, C.byref(h)))
        else:
            fl = None if item_flags is None else np.ascontiguousarray(item_flags, np.uint32)
            check(lib().sgc_plan_create_exchange_ex(level.h, _v3(dbl.domain.lo), _v3(dbl.domain.hi), _v3(dbl.maxbox), nghost,
                                                    periodic, trim, comp_begin, nc, dbl.nproc, dbl.rank,
                                                    None if fl is None else fl.ctypes.data, 0 if fl is None else len(fl),
                                                    C.byref(h)))
        self.h = h
        return self

    def matches(self, level):
        return bool(lib().sgc_plan_matches_level(self.h, level.h))

    def defineItems(self, level, items, comp_begin=0, ncomp=None):
        self._release()
        items = np.ascontiguousarray(items, ITEM_DTYPE)
        h = C.c_void_p()
        check(lib().sgc_plan_create_items(level.h, items.ctypes.data, len(items), comp_begin,
                                          level.ncomp if ncomp is None else ncomp, C.byref(h)))
        self.h = h
        return self

    def defineBCZeroGradient(self, level, domain):
        self._release()
        h = C.c_void_p()
        check(lib().sgc_plan_create_bc_zero_gradient(level.h, _v3(domain.lo), _v3(domain.hi), C.byref(h)))
        self.h = h
        return self

    def numMotionItem(self):
        return int(lib().sgc_plan_num_items(self.h))

    def numRemoteItem(self):
        return int(lib().sgc_plan_num_remote_items(self.h))

    def numCells(self):
        return int(lib().sgc_plan_num_cells(self.h))

    def reboxFactors(self):
        """(decided, (cx, cy, cz)): user boxes per device fab in sgc_heat_run's re-boxed step loop"""
        f = (C.c_int * 3)()
        decided = check(lib().sgc_plan_rebox_factors(self.h, f))
        return bool(decided), tuple(f)

    def pillars(self):
        """(npil, nbz): the two-step sweep marches down npil pillars of nbz boxes stacked in z"""
        p = (C.c_int * 2)()
        check(lib().sgc_plan_pillars(self.h, p))
        return tuple(p)

    def _release(self):
        if self.h and _lib is not None:
            _lib.sgc_plan_destroy(self.h)
        self.h = None

    def __del__(self):
        self._release()


def heat_sweep(unew, u, f, flags=SGC_FMA, box_range=None):
    if box_range is None:
        check(lib().sgc_heat_sweep(unew.h, u.h, f, flags))
    elif len(box_range) == 4:
        check(lib().sgc_heat_sweep_ranges(unew.h, u.h, f, flags, *box_range))
    else:
        check(lib().sgc_heat_sweep_range(unew.h, u.h, f, flags, box_range[0], box_range[1]))


def stencil7_apply(out, u, weights, flags=SGC_FMA):
    """out = w0*u + w1*u[i-1] + w2*u[i+1] + w3*u[j-1] + w4*u[j+1] + w5*u[k-1] + w6*u[k+1] on every component"""
    w = (C.c_double * 7)(*[float(x) for x in weights])
    check(lib().sgc_stencil7_apply(out.h, u.h, w, flags))


def laplacian_apply(B, A, coef=0.5, niter=16, fuse_iters=0, flags=SGC_FMA):
    check(lib().sgc_laplacian_apply(B.h, A.h, coef, niter, fuse_iters, flags))


def wave_step(unp1, un, unm1, factor, flags=SGC_FMA):
    check(lib().sgc_wave_step(unp1.h, un.h, unm1.h, factor, flags))


def wave_run(levels, copier, bc, factor, nstep, flags=SGC_FMA):
    """nstep x {exchange; BC; wave step; rotate}; returns (un, scratch, unm1) levels after the run."""
    hs = (C.c_void_p * 3)(*[l.h for l in levels])
    latest = (C.c_int * 3)()
    check(lib().sgc_wave_run(hs, copier.h if copier else None, bc.h if bc else None, factor, flags, nstep, latest))
    return tuple(levels[i] for i in latest)


LOOP_FUSED, LOOP_TWO_STEP_LOCAL, LOOP_TWO_STEP_SLAB, LOOP_TWO_STEP_BND = 1, 2, 4, 8


def step_loop_class(dlo, dhi, maxbox, nghost=1, periodic=7, trim=0, nproc=1, rank=0):
    """Mask of LOOP_* bits: the device-resident step loops the exchange plan of this layout admits (host only)."""
    mb = (maxbox,) * 3 if np.isscalar(maxbox) else maxbox
    r = lib().sgc_step_loop_class(_i3(*dlo), _i3(*dhi), _i3(*mb), nghost, periodic, trim, nproc, rank)
    if r < 0:
        check(r)
    return r


def step_loop_pillars(dlo, dhi, maxbox, nghost=1, periodic=7, trim=0, nproc=1, rank=0):
    """(npil, nbz) of the two-step sweep's pillars for rank `rank`'s boxes of this layout and exchange (host only)."""
    mb = (maxbox,) * 3 if np.isscalar(maxbox) else maxbox
    p = (C.c_int * 2)()
    r = lib().sgc_step_loop_pillars(_i3(*dlo), _i3(*dhi), _i3(*mb), nghost, periodic, trim, nproc, rank, p)
    if r < 0:
        check(r)
    return tuple(p)


def heat_run(a, b, copier, f, nstep, flags=SGC_FMA):
    """nstep x {exchange; sweep}; returns the level holding the result."""
    r = C.c_int()
    check(lib().sgc_heat_run(a.h, b.h, copier.h, f, flags, nstep, C.byref(r)))
    return b if r.value else a


class CopySet:
    """A recorded list of BaseFab::copy calls (COPY_OP_DTYPE records, local box indices) replayed in
    one launch: LBPatch::stream (LBPatch.H:49-64), the LB bounce-back fill (LBLevel.cpp:65-112)."""

    def __init__(self, dst_like, src_like, ops):
        ops = np.ascontiguousarray(ops, COPY_OP_DTYPE)
        self.h = C.c_void_p()
        check(lib().sgc_copyset_create(dst_like.h, src_like.h, ops.ctypes.data, len(ops), C.byref(self.h)))
        self.nop = len(ops)

    def run(self, dst, src):
        check(lib().sgc_copyset_run(self.h, dst.h, src.h))

    def numCells(self):
        return int(lib().sgc_copyset_num_cells(self.h))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.sgc_copyset_destroy(self.h)
            self.h = None


LB_NVEL, LB_NMACRO = 19, 4
LB_PERIODIC, LB_TRIM = 3, 8         # PeriodicX|PeriodicY, TrimCorner (LBLevel.cpp:13)


def lb_copy_ops(fi, which):
    """which = 0: the stream copies, 1: the bounce-back copies, as COPY_OP_DTYPE records"""
    dom = fi.dbl.domain
    n = check(lib().sgc_lb_copy_ops(fi.h, _v3(dom.lo), _v3(dom.hi), which, None, 0))
    ops = np.zeros(n, COPY_OP_DTYPE)
    check(lib().sgc_lb_copy_ops(fi.h, _v3(dom.lo), _v3(dom.hi), which, ops.ctypes.data, n))
    return ops


def lb_selftest_divconst(which, x):
    """(fast, exact, constant): the collision's 3-instruction x / c against the device's IEEE division"""
    x = np.ascontiguousarray(x, np.float64)
    fast, exact, c = np.empty_like(x), np.empty_like(x), C.c_double()
    check(lib().sgc_lb_selftest_divconst(which, x.ctypes.data, fast.ctypes.data, exact.ctypes.data, x.size, C.byref(c)))
    return fast, exact, c.value


def lb_collide(fi, macro):
    check(lib().sgc_lb_collide(fi.h, macro.h))


def lb_macroscopic(macro, fi):
    check(lib().sgc_lb_macroscopic(macro.h, fi.h))


class LBLevel:
    """The reference's LBLevel (LBLevel.H:12-49): distribution levels m_curr / m_prev (19 comps),
    m_macro_comps (4 comps), initialData, advance — device-resident."""

    def __init__(self, dbl):
        self.dbl = dbl
        self.curr = LevelData(dbl, LB_NVEL, 1)
        self.prev = LevelData(dbl, LB_NVEL, 1)
        self.macro = LevelData(dbl, LB_NMACRO, 1)
        self.copier = Copier().defineExchangeLD(self.curr, LB_PERIODIC, LB_TRIM)
        self.bounce = CopySet(self.curr, self.curr, lb_copy_ops(self.curr, 1))
        self.stream = CopySet(self.prev, self.curr, lb_copy_ops(self.curr, 0))
        self.initialData()

    def initialData(self):
        w = [1. / 3.] + [1. / 18.] * 6 + [1. / 36.] * 12
        for k in range(LB_NVEL):
            self.curr.setVal(w[k], comp=k)
            self.prev.setVal(w[k], comp=k)
        self.macro.setVal(1.0, comp=0)
        for k in range(1, 4):
            self.macro.setVal(0.0, comp=k)

    def advance(self, nstep=1, fuse=True):
        r = C.c_int()
        check(lib().sgc_lb_run(self.curr.h, self.prev.h, self.macro.h, self.copier.h, self.bounce.h, self.stream.h,
                               nstep, 1 if fuse else 0, C.byref(r)))
        if r.value:
            self.curr, self.prev = self.prev, self.curr

    def advance_phases(self):
        """one step as the five separate calls of LBLevel::advance"""
        lb_collide(self.curr, self.macro)
        self.curr.exchange(self.copier)
        self.bounce.run(self.curr, self.curr)
        self.stream.run(self.prev, self.curr)
        self.curr, self.prev = self.prev, self.curr
        lb_macroscopic(self.macro, self.curr)


def comm_unique_id():
    buf = C.create_string_buffer(128)
    check(lib().sgc_comm_unique_id(buf))
    return buf.raw


def comm_init(nproc, rank, unique_id):
    check(lib().sgc_comm_init(nproc, rank, unique_id))


def comm_finalize():
    check(lib().sgc_comm_finalize())
