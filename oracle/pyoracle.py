"""TEST INFRASTRUCTURE — ctypes access to the oracle (oracle/_build/libsgcoracle.so, the plain-C
restatement in oracle/sgc_oracle.c) and, when present, to the REAL reference library compiled from
/root/reference (oracle/_ref/libsgcref*.so, via oracle/ref_shim.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker.  Nothing under structuredgridcalc-boxframework_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ITEM_INTS = 30          # oracle record (sgc_oracle.c SGCO_ITEM_INTS)
REF_ITEM_INTS = 24      # reference-shim record (ref_shim.cpp sgcref_copier_items)

_i3 = C.c_int * 3
_P = np.ctypeslib.ndpointer
_ip = _P(dtype=np.int32, flags="C_CONTIGUOUS")
_dp = _P(dtype=np.float64, flags="C_CONTIGUOUS")


def i3(v):
    return _i3(int(v[0]), int(v[1]), int(v[2]))


def build(force=False):
    """Compile the oracle (and the reference shim when /root/reference is present)."""
    so = os.path.join(HERE, "_build", "libsgcoracle.so")
    src = os.path.join(HERE, "sgc_oracle.c")
    ref_ok = os.path.exists(os.path.join(HERE, "_ref", "libsgcref.so")) or not os.path.isdir(
        os.environ.get("REF_DIR", "/root/reference/Structured_gridcalc"))
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src) or not ref_ok:
        subprocess.run(["bash", os.path.join(HERE, "build_oracle.sh")], check=True)
    return so


class Oracle:
    """The C restatement.  All arrays are numpy; level storage is the concatenation of the fabs."""

    def __init__(self):
        self.lib = L = C.CDLL(build())
        L.sgco_layout.argtypes = [_i3, _i3, _i3, C.c_int, _ip, _ip, C.c_int, _i3]
        L.sgco_copier_items.argtypes = [_i3, _i3, _i3, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int,
                                        C.c_void_p, C.c_int]
        L.sgco_level_offsets.argtypes = [_ip, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.sgco_level_offsets.restype = C.c_long
        L.sgco_level_exchange.argtypes = [_ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, C.c_int,
                                          C.c_int, _dp]
        L.sgco_level_heat_sweep.argtypes = [_ip, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_int]
        L.sgco_level_heat.argtypes = [_i3, _i3, _i3, C.c_int, C.c_uint, C.c_uint, C.c_double, C.c_int,
                                      C.c_int, _dp, _dp]
        L.sgco_level_wave_sweep.argtypes = [_ip, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_int]
        L.sgco_laplacian.argtypes = [C.c_int, C.c_int, _dp, _dp]
        L.sgco_wave.argtypes = [_i3, _i3, C.c_double, C.c_int, C.c_int, _dp, _dp]
        L.sgco_wave_dirsum.argtypes = [_i3, _i3, C.c_double, C.c_int, C.c_int, _dp, _dp]
        L.sgco_bc_zero_gradient.argtypes = [_i3, _i3, C.c_int, _dp]
        L.sgco_level_fill_hash.argtypes = [_ip, C.c_int, C.c_int, C.c_double, _dp]
        L.sgco_level_gather.argtypes = [_i3, _i3, _ip, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp]
        L.sgco_lb_constants.argtypes = [_dp, _ip, _ip, _dp]
        L.sgco_lb_collision.argtypes = [_ip, C.c_int, C.c_int, _dp, _dp]
        L.sgco_lb_macroscopic.argtypes = [_ip, C.c_int, C.c_int, _dp, _dp]
        L.sgco_lb_stream.argtypes = [_ip, C.c_int, C.c_int, _dp, _dp]
        L.sgco_lb_bounce.argtypes = [_i3, _i3, _ip, C.c_int, C.c_int, _dp]
        L.sgco_lb_advance.argtypes = [_i3, _i3, _i3, C.c_int, _dp, _dp, _dp]
        L.sgco_level_valid_out.argtypes = [_ip, C.c_int, C.c_int, C.c_int, _dp, _dp]
        L.sgco_level_valid_out.restype = C.c_long
        L.sgco_fnv1a64.argtypes = [C.c_void_p, C.c_long]
        L.sgco_fnv1a64.restype = C.c_uint64
        L.sgco_adjbox.argtypes = [_i3, _i3, C.c_int, C.c_int, C.c_int, _ip]
        fabargs = [_i3, _i3, C.c_int, _dp]
        L.sgco_fab_setval.argtypes = fabargs + [C.c_int, C.c_double]
        L.sgco_fab_copy.argtypes = fabargs + fabargs + [_i3, _i3, C.c_int, _i3, _i3, C.c_int, C.c_int, C.c_uint]
        L.sgco_fab_copy_same.argtypes = fabargs + [_i3, _i3, _dp, _i3, _i3]
        L.sgco_fab_linear_out.argtypes = fabargs + [_i3, _i3, C.c_int, C.c_int, C.c_uint, _dp]
        L.sgco_fab_linear_out.restype = C.c_long
        L.sgco_fab_linear_in.argtypes = fabargs + [_i3, _i3, C.c_int, C.c_int, C.c_uint, _dp]
        L.sgco_fab_linear_in.restype = C.c_long

    # -- layout / plan
    def layout(self, dlo, dhi, maxbox, nproc=1):
        n = 1
        for d in range(3):
            n *= (dhi[d] - dlo[d] + 1) // maxbox[d]
        boxes = np.zeros((n, 6), np.int32)
        owner = np.zeros(n, np.int32)
        nb = _i3()
        r = self.lib.sgco_layout(i3(dlo), i3(dhi), i3(maxbox), nproc, boxes, owner, n, nb)
        if r < 0:
            raise ValueError("sgco_layout failed: %d" % r)
        return boxes, owner, tuple(nb)

    def copier_items(self, dlo, dhi, maxbox, ng, periodic=0, trim=0, nproc=1, procid=0):
        n = self.lib.sgco_copier_items(i3(dlo), i3(dhi), i3(maxbox), ng, periodic, trim, nproc, procid, None, 0)
        if n < 0:
            raise ValueError("sgco_copier_items failed: %d" % n)
        items = np.zeros((max(n, 1), ITEM_INTS), np.int32)
        self.lib.sgco_copier_items(i3(dlo), i3(dhi), i3(maxbox), ng, periodic, trim, nproc, procid,
                                   items.ctypes.data, n)
        return items[:n]

    def level_size(self, boxes, ng, ncomp):
        boxes = np.ascontiguousarray(boxes, np.int32)
        return int(self.lib.sgco_level_offsets(boxes, len(boxes), ng, ncomp, None))

    def level_offsets(self, boxes, ng, ncomp):
        boxes = np.ascontiguousarray(boxes, np.int32)
        off = np.zeros(len(boxes), np.int64)
        self.lib.sgco_level_offsets(boxes, len(boxes), ng, ncomp, off.ctypes.data)
        return off

    def level_exchange(self, boxes, ng, ncomp, items, fabs, c0=0, nc=None):
        boxes = np.ascontiguousarray(boxes, np.int32)
        items = np.ascontiguousarray(items, np.int32)
        nc = ncomp if nc is None else nc
        return self.lib.sgco_level_exchange(boxes, len(boxes), ng, ncomp, c0, nc, items, len(items),
                                            items.shape[1] if len(items) else ITEM_INTS, fabs)

    def level_heat_sweep(self, boxes, ng, u, unew, f, contract=1):
        boxes = np.ascontiguousarray(boxes, np.int32)
        self.lib.sgco_level_heat_sweep(boxes, len(boxes), ng, u, unew, f, contract)

    def level_wave_sweep(self, boxes, ng, un, unm1, unp1, factor, contract=1):
        boxes = np.ascontiguousarray(boxes, np.int32)
        self.lib.sgco_level_wave_sweep(boxes, len(boxes), ng, un, unm1, unp1, factor, contract)

    def level_heat(self, dlo, dhi, maxbox, ng, periodic, trim, f, nstep, fabs, contract=1):
        t = np.zeros(3)
        n = self.lib.sgco_level_heat(i3(dlo), i3(dhi), i3(maxbox), ng, periodic, trim, f, nstep, contract, fabs, t)
        if n < 0:
            raise ValueError("sgco_level_heat failed: %d" % n)
        return n, t

    def laplacian(self, n, niter, A):
        B = np.zeros((n - 2) ** 3)
        self.lib.sgco_laplacian(n, niter, A, B)
        return B

    def wave(self, dlo, dhi, factor, nstep, u0, u1, contract=1):
        self.lib.sgco_wave(i3(dlo), i3(dhi), factor, nstep, contract, u0, u1)

    def wave_dirsum(self, dlo, dhi, factor, nstep, u0, u1, contract=1):
        """the shipped wave app's pencil form (MD_DIRSUM), WavePatch.cpp:194-206"""
        self.lib.sgco_wave_dirsum(i3(dlo), i3(dhi), factor, nstep, contract, u0, u1)

    def bc_zero_gradient(self, dlo, dhi, ng, u):
        self.lib.sgco_bc_zero_gradient(i3(dlo), i3(dhi), ng, u)

    def fill_hash(self, boxes, ng, ghostval=0.0):
        boxes = np.ascontiguousarray(boxes, np.int32)
        fabs = np.zeros(self.level_size(boxes, ng, 1))
        self.lib.sgco_level_fill_hash(boxes, len(boxes), ng, ghostval, fabs)
        return fabs

    def gather(self, dlo, dhi, boxes, ng, ncomp, comp, fabs):
        boxes = np.ascontiguousarray(boxes, np.int32)
        g = np.zeros(int(np.prod([dhi[d] - dlo[d] + 1 for d in range(3)])))
        self.lib.sgco_level_gather(i3(dlo), i3(dhi), boxes, len(boxes), ng, ncomp, comp, fabs, g)
        return g

    def fnv1a64(self, arr):
        arr = np.ascontiguousarray(arr)
        return int(self.lib.sgco_fnv1a64(arr.ctypes.data, arr.nbytes))

    def adjbox(self, lo, hi, ncell, d, side):
        out = np.zeros(6, np.int32)
        self.lib.sgco_adjbox(i3(lo), i3(hi), ncell, d, side, out)
        return out

    # ---- lattice Boltzmann application (LBPatch.H / LBPhysics.H / LBLevel.cpp restated)
    def lb_collision(self, boxes, fi, macro):
        boxes = np.ascontiguousarray(boxes, np.int32)
        self.lib.sgco_lb_collision(boxes, len(boxes), 1, fi, macro)

    def lb_macroscopic(self, boxes, fi, macro):
        boxes = np.ascontiguousarray(boxes, np.int32)
        self.lib.sgco_lb_macroscopic(boxes, len(boxes), 1, fi, macro)

    def lb_stream(self, boxes, fi, prev):
        boxes = np.ascontiguousarray(boxes, np.int32)
        self.lib.sgco_lb_stream(boxes, len(boxes), 1, fi, prev)

    def lb_bounce(self, dlo, dhi, boxes, fi):
        boxes = np.ascontiguousarray(boxes, np.int32)
        self.lib.sgco_lb_bounce(i3(dlo), i3(dhi), boxes, len(boxes), 1, fi)

    def lb_advance(self, dlo, dhi, maxbox, nstep, fi, prev, macro):
        n = self.lib.sgco_lb_advance(i3(dlo), i3(dhi), i3(maxbox), nstep, fi, prev, macro)
        if n < 0:
            raise ValueError("sgco_lb_advance failed: %d" % n)
        return n

    def lb_initial(self, boxes):
        """LBLevel::initialData (LBLevel.H:82-94): f_q = w_q everywhere, macro = (1, 0, 0, 0)."""
        boxes = np.ascontiguousarray(boxes, np.int32)
        w = np.zeros(19); e = np.zeros(57, np.int32); opp = np.zeros(19, np.int32); c = np.zeros(3)
        self.lib.sgco_lb_constants(w, e, opp, c)
        fi, macro = [], []
        for b in boxes:
            S = int(np.prod(b[3:] - b[:3] + 3))
            fi.append(np.repeat(w, S))
            macro.append(np.repeat(np.array([1.0, 0, 0, 0]), S))
        return np.concatenate(fi), np.concatenate(macro)

    def level_valid_out(self, boxes, ng, ncomp, fabs):
        boxes = np.ascontiguousarray(boxes, np.int32)
        n = int(sum(np.prod(b[3:] - b[:3] + 1) for b in boxes)) * ncomp
        out = np.zeros(n)
        got = self.lib.sgco_level_valid_out(boxes, len(boxes), ng, ncomp, fabs, out)
        assert got == n
        return out


def ref_path(variant=""):
    return os.path.join(HERE, "_ref", "libsgcref%s.so" % variant)


def have_reference(variant=""):
    return os.path.exists(ref_path(variant))


class Reference:
    """The real reference library (compiled from /root/reference) behind oracle/ref_shim.cpp.
    variant: "" (release, g++ default -ffp-contract=fast), "_nofma", "_omp", "_dbg" (asserts live)."""

    def __init__(self, variant=""):
        self.variant = variant
        self.lib = L = C.CDLL(ref_path(variant))
        # "_sp", "_sp_nofma": the reference built with -DUSE_SINGLE_PRECISION (Real = float): float32 buffers
        self.dtype = np.float32 if variant.startswith("_sp") else np.float64
        _dp = _P(dtype=self.dtype, flags="C_CONTIGUOUS")
        _real = C.c_float if self.dtype == np.float32 else C.c_double
        L.sgcref_openmp_threads.restype = C.c_int
        L.sgcref_layout.argtypes = [_i3, _i3, _i3, C.c_int, _ip, _ip, C.c_int]
        L.sgcref_copier_items.argtypes = [_i3, _i3, _i3, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int,
                                          C.c_int, C.c_void_p, C.c_int]
        L.sgcref_level_exchange.argtypes = [_i3, _i3, _i3, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int, _dp]
        L.sgcref_level_heat.argtypes = [_i3, _i3, _i3, C.c_int, C.c_uint, C.c_uint, _real, C.c_int, _dp,
                                        _P(dtype=np.float64, flags="C_CONTIGUOUS")]
        L.sgcref_laplacian.argtypes = [C.c_int, C.c_int, _dp, _dp]
        L.sgcref_laplacian.restype = C.c_double
        L.sgcref_laplacian_boxiter.argtypes = [C.c_int, C.c_int, _dp, _dp]
        L.sgcref_laplacian_boxiter.restype = C.c_double
        L.sgcref_wave.argtypes = [_i3, _i3, _real, C.c_int, _dp, _dp]
        fabargs = [_i3, _i3, C.c_int, _dp]
        L.sgcref_fab_copy.argtypes = fabargs + fabargs + [_i3, _i3, C.c_int, _i3, _i3, C.c_int, C.c_int, C.c_uint]
        L.sgcref_fab_copy_same.argtypes = fabargs + [_i3, _i3, _dp, _i3, _i3]
        L.sgcref_fab_linear_out.argtypes = fabargs + [_i3, _i3, C.c_int, C.c_int, C.c_uint, _dp]
        L.sgcref_fab_linear_in.argtypes = fabargs + [_i3, _i3, C.c_int, C.c_int, C.c_uint, _dp]
        L.sgcref_fab_setval.argtypes = fabargs + [C.c_int, _real]
        L.sgcref_adjbox.argtypes = [_i3, _i3, C.c_int, C.c_int, C.c_int, _ip]

    def threads(self):
        return int(self.lib.sgcref_openmp_threads())

    def set_threads(self, n):
        """OpenMP builds only: use n threads regardless of OMP_NUM_THREADS (torchrun exports 1)"""
        if hasattr(self.lib, "sgcref_set_threads"):
            return int(self.lib.sgcref_set_threads(int(n)))
        return self.threads()

    def layout(self, dlo, dhi, maxbox, nproc=1):
        n = 1
        for d in range(3):
            n *= (dhi[d] - dlo[d] + 1) // maxbox[d]
        boxes = np.zeros((n, 6), np.int32)
        owner = np.zeros(n, np.int32)
        r = self.lib.sgcref_layout(i3(dlo), i3(dhi), i3(maxbox), nproc, boxes, owner, n)
        assert r == n
        return boxes, owner

    def copier_items(self, dlo, dhi, maxbox, ng, ncomp=1, periodic=0, trim=0, nproc=1, procid=0):
        n = self.lib.sgcref_copier_items(i3(dlo), i3(dhi), i3(maxbox), ng, ncomp, periodic, trim, nproc, procid,
                                         None, 0)
        items = np.zeros((max(n, 1), REF_ITEM_INTS), np.int32)
        self.lib.sgcref_copier_items(i3(dlo), i3(dhi), i3(maxbox), ng, ncomp, periodic, trim, nproc, procid,
                                     items.ctypes.data, n)
        return items[:n]

    def level_exchange(self, dlo, dhi, maxbox, ng, ncomp, periodic, trim, fabs, mode=0):
        return self.lib.sgcref_level_exchange(i3(dlo), i3(dhi), i3(maxbox), ng, ncomp, periodic, trim, mode, fabs)

    def level_heat(self, dlo, dhi, maxbox, ng, periodic, trim, f, nstep, fabs):
        t = np.zeros(3)
        n = self.lib.sgcref_level_heat(i3(dlo), i3(dhi), i3(maxbox), ng, periodic, trim, f, nstep, fabs, t)
        return n, t

    def laplacian(self, n, niter, A, boxiter=False):
        B = np.zeros((n - 2) ** 3, self.dtype)
        fn = self.lib.sgcref_laplacian_boxiter if boxiter else self.lib.sgcref_laplacian
        ms = fn(n, niter, A, B)
        return B, ms

    def wave(self, dlo, dhi, factor, nstep, u0, u1):
        self.lib.sgcref_wave(i3(dlo), i3(dhi), factor, nstep, u0, u1)

    def adjbox(self, lo, hi, ncell, d, side):
        out = np.zeros(6, np.int32)
        self.lib.sgcref_adjbox(i3(lo), i3(hi), ncell, d, side, out)
        return out


class ReferenceWave:
    """The reference's shipped WavePatch time stepper (oracle/_ref/libsgcref_wave.so)."""

    def __init__(self):
        self.lib = L = C.CDLL(ref_path("_wave"))
        L.sgcref_wavepatch_run.argtypes = [C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, _dp, _dp,
                                           C.POINTER(C.c_double)]

    def run(self, h, cfl, nstep, u_init=None, um1_init=None):
        n = (h + 2) ** 3
        un, unm1 = np.zeros(n), np.zeros(n)
        f = C.c_double()
        if not os.path.isdir("plot"):
            pass
        self.lib.sgcref_wavepatch_run(h, cfl, nstep,
                                      None if u_init is None else u_init.ctypes.data,
                                      None if um1_init is None else um1_init.ctypes.data, un, unm1, C.byref(f))
        return un, unm1, f.value


class ReferenceLB:
    """The reference's lattice-Boltzmann application (oracle/_ref/libsgcref_lb{,_nofma}.so, built from
    application/latticeBoltzmann + oracle/ref_lb_shim.cpp with the recording CGNS stand-in)."""
    NVEL, NMACRO = 19, 4

    def __init__(self, variant=""):
        self.lib = L = C.CDLL(ref_path("_lb" + variant))
        _i3p = _i3
        L.sgcref_lb_run.restype = C.c_double
        L.sgcref_lb_run.argtypes = [_i3p, _i3p, _i3p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sgcref_lb_phase.argtypes = [_i3p, _i3p, _i3p, C.c_int, _dp, _dp, _dp]
        L.sgcref_lb_plot.argtypes = [_i3p, _i3p, _i3p, C.c_int, C.c_void_p, _P(np.int64, flags="C_CONTIGUOUS"),
                                     C.c_void_p, C.c_void_p, C.c_void_p]

    @staticmethod
    def _p(a):
        return None if a is None else a.ctypes.data

    def run(self, dlo, dhi, maxbox, nstep, nfi, nmacro, fi_init=None, macro_init=None):
        """nstep x LBLevel::advance(); returns (fi, macro, total mass)."""
        fi, macro = np.zeros(nfi), np.zeros(nmacro)
        mass = self.lib.sgcref_lb_run(i3(dlo), i3(dhi), i3(maxbox), nstep, self._p(fi_init), self._p(macro_init),
                                      fi.ctypes.data, macro.ctypes.data)
        return fi, macro, mass

    def phase(self, dlo, dhi, maxbox, phase, fi, macro, prev):
        """0 collision, 1 macroscopic, 2 stream — in place on copies, returns (fi, macro, prev)."""
        fi, macro, prev = fi.copy(), macro.copy(), prev.copy()
        if self.lib.sgcref_lb_phase(i3(dlo), i3(dhi), i3(maxbox), phase, fi, macro, prev):
            raise ValueError("bad phase")
        return fi, macro, prev

    def plot(self, dlo, dhi, maxbox, nstep, macro_init=None):
        """what LBLevel::writePlotFile hands to CGNS: (data, meta, names)"""
        sizes = np.zeros(3, np.int64)
        self.lib.sgcref_lb_plot(i3(dlo), i3(dhi), i3(maxbox), nstep, self._p(macro_init), sizes, None, None, None)
        data, meta = np.zeros(sizes[0]), np.zeros(sizes[1], np.int32)
        names = C.create_string_buffer(int(sizes[2]) + 1)
        err = self.lib.sgcref_lb_plot(i3(dlo), i3(dhi), i3(maxbox), nstep, self._p(macro_init), sizes,
                                      data.ctypes.data, meta.ctypes.data, names)
        if err:
            raise ValueError("writePlotFile failed: %d" % err)
        return data, meta, names.value.decode().split("\n")[:-1]
