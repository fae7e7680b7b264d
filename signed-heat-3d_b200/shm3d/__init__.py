"""
shm3d -- thin ctypes binding of libshm3d_grid.so (include/shm3d_grid.h) plus a Python mirror of the reference's
operator interface for the grid path (SignedHeat3DOptions / SignedHeatGridSolver.computeDistance,
reference include/signed_heat_3d.h:20-28 and include/signed_heat_grid_solver.h:11-22).

Plumbing only: every number is produced by the CUDA library.  There is NO CPU fallback -- creating a solver
without a usable GPU raises Shm3dError(SHM3D_ERR_CUDA).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_PKG), "lib", "libshm3d_grid.so")

OK, ERR_INVALID_ARG, ERR_CUDA, ERR_NONFINITE, ERR_FACTORIZATION, ERR_NO_CONVERGENCE, ERR_NCCL = range(7)
FLAG_FAST, FLAG_SCRUB_NONFINITE, FLAG_VERBOSE, FLAG_NO_MG, FLAG_PLAIN_MG, FLAG_PROFILE = 1, 2, 4, 8, 16, 32
FLAG_NO_TMA = 128         # diagnostics: row-streaming stencil kernels instead of the TMA-staged marching ones
FLAG_TAIL_PROGRAM = 256     # experiment, off by default: coarsest multigrid levels as one program launch (mg_tail.cuh)
FLAG_NO_PDL = 1024        # diagnostics: projector sweep kernels without programmatic dependent launch
FLAG_NO_CYCLIC_SUM = 2048  # diagnostics (slab contexts): no round-robin z-chunks for Steps 1-2
FLAG_NO_GRAPH = 512       # diagnostics: PCG iterations launched kernel by kernel instead of CUDA-graph replays
FLAG_FP64_UNDERFLOW = 64  # reproduce the reference's fp64 underflow in X.norm() at far nodes (include/shm3d_grid.h)


class Params(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("bbox_min", C.c_double * 3),
                ("cell", C.c_double), ("lambda_", C.c_double), ("flags", C.c_uint32), ("cull_tau", C.c_double),
                ("cg_rel_tol", C.c_double), ("cg_max_iters", C.c_int32), ("mg_smooth", C.c_int32),
                ("mg_constrained_from", C.c_int32), ("reserved", C.c_int32)]

    @property
    def N(self):
        return self.nx * self.ny * self.nz


class Stats(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_h2d", C.c_double), ("ms_sum", C.c_double), ("ms_rhs", C.c_double),
                ("ms_constraints", C.c_double), ("ms_pcg", C.c_double), ("ms_shift", C.c_double),
                ("ms_d2h", C.c_double), ("pairs_evaluated", C.c_int64), ("pairs_bruteforce", C.c_int64),
                ("n_clusters", C.c_int32), ("m_constraints", C.c_int32), ("cg_iters", C.c_int32),
                ("cg_rel_residual", C.c_double), ("shift", C.c_double), ("kernel_launches", C.c_int64),
                ("ms_pcg_stencil", C.c_double), ("pcg_stencil_launches", C.c_int64), ("ms_pcg_vcycle", C.c_double),
                ("ms_pcg_projector", C.c_double), ("pcg_projector_applies", C.c_int64), ("ms_pcg_update", C.c_double),
                ("pcg_vcycles", C.c_int64), ("tail_ops", C.c_int32), ("graph_replays", C.c_int32)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class IsoStats(C.Structure):
    _fields_ = [("n_vertices", C.c_int64), ("n_triangles", C.c_int64), ("ms_device", C.c_double),
                ("gpu_launches", C.c_int64)]


FIELD_DEVICE_F32, FIELD_HOST_F64, FIELD_HOST_F32 = 0, 1, 2
ISO_LATTICE = 1


class Shm3dError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"shm3d error {code}: {msg}")
        self.code = code


EXPORTS = ["shm3d_slab_range", "shm3d_ctx_create", "shm3d_ctx_create_dist", "shm3d_nccl_unique_id", "shm3d_ctx_destroy",
           "shm3d_last_error", "shm3d_slab", "shm3d_solve", "shm3d_solve_device", "shm3d_step12", "shm3d_rhs",
           "shm3d_step3", "shm3d_prepare_mesh", "shm3d_prepare_points", "shm3d_debug_constraints",
           "shm3d_debug_factor_solve", "shm3d_version", "shm3d_ctx_stream", "shm3d_host_alloc", "shm3d_host_free", "shm3d_step12_points", "shm3d_point_weights",
           "shm3d_debug_local_ring", "shm3d_debug_tufted_weights", "shm3d_debug_knn_mode", "shm3d_debug_cyclic_plan", "shm3d_isosurface", "shm3d_isosurface_fetch",
           "shm3d_isosurface_device", "shm3d_slice", "shm3d_debug_stencil_op"]

_lib = None


def lib():
    """Load libshm3d_grid.so (fails loudly if it has not been built: run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Shm3dError(ERR_CUDA, f"{LIB_PATH} is missing -- build it with signed-heat-3d_b200/csrc/build.sh")
        L = C.CDLL(LIB_PATH)
        dp, fp, vp = C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_void_p
        i64p, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        PP, SP = C.POINTER(Params), C.POINTER(Stats)
        L.shm3d_ctx_create.argtypes = [C.POINTER(vp), C.c_int]
        L.shm3d_ctx_create_dist.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp]
        L.shm3d_nccl_unique_id.argtypes = [vp]
        L.shm3d_ctx_destroy.argtypes = [vp]
        L.shm3d_ctx_destroy.restype = None
        L.shm3d_last_error.argtypes = [vp]
        L.shm3d_last_error.restype = C.c_char_p
        L.shm3d_slab.argtypes = [vp, C.c_int32, i32p, i32p]
        L.shm3d_slab_range.argtypes = [C.c_int32, C.c_int32, C.c_int32, i32p, i32p]
        L.shm3d_solve.argtypes = [vp, PP, C.c_int64, dp, dp, dp, dp, SP]
        L.shm3d_solve_device.argtypes = [vp, PP, C.c_int64, vp, vp, vp, vp, SP]
        L.shm3d_step12.argtypes = [vp, PP, C.c_int64, dp, dp, dp, fp, SP]
        L.shm3d_rhs.argtypes = [vp, PP, fp, fp]
        L.shm3d_step12_points.argtypes = [vp, C.c_double, C.c_int64, dp, dp, dp, C.c_int64, dp, fp]
        L.shm3d_step3.argtypes = [vp, PP, C.c_int64, dp, dp, fp, dp, SP]
        L.shm3d_prepare_mesh.argtypes = [dp, C.c_int64, i64p, i64p, C.c_int64, C.c_double, C.c_double, C.c_double, PP,
                                         dp, dp, dp, dp]
        L.shm3d_prepare_points.argtypes = [dp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, PP]
        L.shm3d_debug_constraints.argtypes = [PP, C.c_int64, dp, i32p, i64p, dp, i64p, C.c_int64]
        L.shm3d_debug_factor_solve.argtypes = [PP, C.c_int64, dp, C.c_int32, dp, C.c_int32, dp, i32p]
        L.shm3d_version.restype = C.c_char_p
        L.shm3d_point_weights.argtypes = [dp, dp, C.c_int64, C.c_int32, dp, dp, i64p, i64p, dp, dp]
        L.shm3d_debug_local_ring.argtypes = [dp, C.c_int32, i32p, i32p]
        L.shm3d_debug_knn_mode.argtypes = [C.c_int32]
        L.shm3d_debug_knn_mode.restype = None
        L.shm3d_debug_tufted_weights.argtypes = [dp, C.c_int64, i64p, C.c_int64, dp, dp, i64p, dp, dp]
        L.shm3d_isosurface.argtypes = [vp, PP, vp, C.c_int32, C.c_float, fp, fp, C.c_uint32, C.POINTER(IsoStats)]
        L.shm3d_isosurface_fetch.argtypes = [vp, fp, C.POINTER(C.c_uint32)]
        L.shm3d_isosurface_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
        L.shm3d_slice.argtypes = [vp, PP, vp, C.c_int32, dp, dp, dp, C.c_int32, C.c_int32, fp]
        L.shm3d_debug_stencil_op.argtypes = [vp] + [C.c_int32] * 6 + [fp, fp, fp, dp, C.c_int32, fp, fp, dp, C.c_int32, dp]
        L.shm3d_ctx_stream.argtypes = [vp]
        L.shm3d_ctx_stream.restype = vp
        L.shm3d_host_alloc.argtypes = [C.c_size_t]
        L.shm3d_host_alloc.restype = vp
        L.shm3d_host_free.argtypes = [vp]
        L.shm3d_host_free.restype = None
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def flatten_faces(faces):
    """list-of-lists (or [nF,3] array) -> (face_vertices int64[], face_offsets int64[nF+1])"""
    if isinstance(faces, np.ndarray) and faces.ndim == 2:
        nF, d = faces.shape
        return np.ascontiguousarray(faces, dtype=np.int64).ravel(), np.arange(0, (nF + 1) * d, d, dtype=np.int64)
    off = np.zeros(len(faces) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(f) for f in faces])
    return np.asarray([v for f in faces for v in f], dtype=np.int64), off


def prepare_mesh(V, faces, tCoef=1.0, hCoef=0.0, scale=2.0):
    """Host half of computeDistance(VertexPositionGeometry&) -- rows a4-a6.  Returns (Params, pos, nrm, area, h)."""
    V = _c64(V)
    fv, fo = flatten_faces(faces)
    nF = len(fo) - 1
    p = Params()
    pos = np.empty((nF, 3))
    nrm = np.empty((nF, 3))
    area = np.empty(nF)
    h = C.c_double()
    rc = lib().shm3d_prepare_mesh(_dp(V), len(V), fv.ctypes.data_as(C.POINTER(C.c_int64)),
                                  fo.ctypes.data_as(C.POINTER(C.c_int64)), nF, tCoef, hCoef, scale, C.byref(p),
                                  _dp(pos), _dp(nrm), _dp(area), C.byref(h))
    if rc != OK:
        raise Shm3dError(rc, "shm3d_prepare_mesh: invalid mesh")
    return p, pos, nrm, area, h.value


def prepare_points(P, h, tCoef=1.0, hCoef=0.0, scale=2.0):
    P = _c64(P)
    p = Params()
    rc = lib().shm3d_prepare_points(_dp(P), len(P), h, tCoef, hCoef, scale, C.byref(p))
    if rc != OK:
        raise Shm3dError(rc, "shm3d_prepare_points: invalid input")
    return p


def point_weights(P, normals, k=30, diagnostics=False):
    """Per-point areas and mean edge length for the point-cloud overload (row N1: geometry-central's tufted-cover
    pipeline restated).  Returns (areas[nP], h, n_soup_triangles) [+ dict(flips, min_cotan, area_before)]."""
    P, Nn = _c64(P), _c64(normals)
    areas = np.empty(len(P))
    h = C.c_double()
    nt, nf = C.c_int64(), C.c_int64()
    mc, ab = C.c_double(), C.c_double()
    rc = lib().shm3d_point_weights(_dp(P), _dp(Nn), len(P), k, _dp(areas), C.byref(h), C.byref(nt), C.byref(nf),
                                   C.byref(mc), C.byref(ab))
    if rc != OK:
        raise Shm3dError(rc, "shm3d_point_weights: invalid input (need more than k points, finite data, some triangles)")
    if diagnostics:
        return areas, h.value, nt.value, dict(flips=nf.value, min_cotan=mc.value, area_before=ab.value)
    return areas, h.value, nt.value


def debug_tufted_weights(P, tris):
    """Tufted cover + intrinsic Delaunay flips of a given triangle soup -> (areas[nP], h, dict(flips, min_cotan, area_before))."""
    P = _c64(P)
    t = np.ascontiguousarray(tris, dtype=np.int64)
    areas = np.empty(len(P))
    h, mc, ab = C.c_double(), C.c_double(), C.c_double()
    nf = C.c_int64()
    rc = lib().shm3d_debug_tufted_weights(_dp(P), len(P), t.ctypes.data_as(C.POINTER(C.c_int64)), len(t), _dp(areas),
                                          C.byref(h), C.byref(nf), C.byref(mc), C.byref(ab))
    if rc != OK:
        raise Shm3dError(rc, "shm3d_debug_tufted_weights: invalid input")
    return areas, h.value, dict(flips=nf.value, min_cotan=mc.value, area_before=ab.value)


def debug_local_ring(coords2d):
    c = _c64(coords2d)
    ring = np.empty(len(c), dtype=np.int32)
    tri = np.empty(len(c), dtype=np.int32)
    n = lib().shm3d_debug_local_ring(_dp(c), len(c), ring.ctypes.data_as(C.POINTER(C.c_int32)),
                                     tri.ctypes.data_as(C.POINTER(C.c_int32)))
    return ring[:n].copy(), tri[:n].copy()


def debug_constraints(p: Params, pos):
    pos = _c64(pos)
    m = C.c_int32()
    rc = lib().shm3d_debug_constraints(C.byref(p), len(pos), _dp(pos), C.byref(m), None, None, None, 0)
    if rc != OK:
        raise Shm3dError(rc, lib().shm3d_last_error(None).decode())
    node = np.empty((m.value, 8), dtype=np.int64)
    w = np.empty((m.value, 8))
    src = np.empty(m.value, dtype=np.int64)
    rc = lib().shm3d_debug_constraints(C.byref(p), len(pos), _dp(pos), C.byref(m),
                                       node.ctypes.data_as(C.POINTER(C.c_int64)), _dp(w),
                                       src.ctypes.data_as(C.POINTER(C.c_int64)), m.value)
    if rc != OK:
        raise Shm3dError(rc, lib().shm3d_last_error(None).decode())
    return src, node, w


def debug_factor_solve(p: Params, pos, v, uniform=True):
    pos = _c64(pos)
    v = np.array(v, dtype=np.float64)
    mb = C.c_double()
    th = C.c_int32()
    rc = lib().shm3d_debug_factor_solve(C.byref(p), len(pos), _dp(pos), 1 if uniform else 0, _dp(v), len(v),
                                        C.byref(mb), C.byref(th))
    if rc != OK:
        raise Shm3dError(rc, lib().shm3d_last_error(None).decode())
    return v, mb.value, th.value


class PinnedArray:
    """numpy view of a page-locked host buffer (shm3d_host_alloc); freed with the object."""

    def __init__(self, shape, dtype=np.float64):
        self.shape = tuple(np.atleast_1d(shape))
        n = int(np.prod(self.shape)) * np.dtype(dtype).itemsize
        self._p = lib().shm3d_host_alloc(n)
        if not self._p:
            raise Shm3dError(ERR_CUDA, f"cudaHostAlloc of {n} bytes failed")
        buf = (C.c_char * n).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(self.shape)

    def __del__(self):
        try:
            p, self._p = self._p, None
            if p:
                self.array = None
                lib().shm3d_host_free(p)
        except Exception:
            pass


class Context:
    """One GPU context (shm3d_ctx).  rank/world/nccl_id select the slab-partitioned multi-GPU mode."""

    def __init__(self, device=0, rank=0, world=1, nccl_id: bytes | None = None):
        self._h = C.c_void_p()
        L = lib()
        if world > 1:
            _point_at_nccl()
            buf = C.create_string_buffer(nccl_id, 128)
            rc = L.shm3d_ctx_create_dist(C.byref(self._h), device, rank, world, C.cast(buf, C.c_void_p))
        else:
            rc = L.shm3d_ctx_create(C.byref(self._h), device)
        if rc != OK:
            raise Shm3dError(rc, L.shm3d_last_error(None).decode())
        self.rank, self.world = rank, world

    def close(self):
        if self._h:
            lib().shm3d_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise Shm3dError(rc, lib().shm3d_last_error(self._h).decode())

    @property
    def stream(self):
        """cudaStream_t handle (int) of the stream this context launches on."""
        return int(lib().shm3d_ctx_stream(self._h) or 0)

    def slab(self, nz):
        k0, k1 = C.c_int32(), C.c_int32()
        self._check(lib().shm3d_slab(self._h, nz, C.byref(k0), C.byref(k1)))
        return k0.value, k1.value

    def local_n(self, p: Params):
        k0, k1 = self.slab(p.nz)
        return p.nx * p.ny * (k1 - k0)

    def solve(self, p: Params, pos, nrm, area, out=None):
        pos, nrm, area = _c64(pos), _c64(nrm), _c64(area)
        phi = out if out is not None else np.empty(self.local_n(p))
        st = Stats()
        self._check(lib().shm3d_solve(self._h, C.byref(p), len(area), _dp(pos), _dp(nrm), _dp(area), _dp(phi),
                                      C.byref(st)))
        return phi, st

    def solve_device(self, p: Params, d_pos, d_nrm, d_area, d_phi, n_sources):
        """device pointers (ints) -- inputs double, output float32[local N]"""
        st = Stats()
        self._check(lib().shm3d_solve_device(self._h, C.byref(p), n_sources, C.c_void_p(d_pos), C.c_void_p(d_nrm),
                                             C.c_void_p(d_area), C.c_void_p(d_phi), C.byref(st)))
        return st

    def step12(self, p: Params, pos, nrm, area):
        pos, nrm, area = _c64(pos), _c64(nrm), _c64(area)
        Y = np.empty((3, self.local_n(p)), dtype=np.float32)
        st = Stats()
        self._check(lib().shm3d_step12(self._h, C.byref(p), len(area), _dp(pos), _dp(nrm), _dp(area), _fp(Y),
                                       C.byref(st)))
        return Y, st

    def step12_points(self, lambda_, pos, nrm, area, query):
        """Steps 1-2 at arbitrary query points [Q,3] -> unit vectors float32[Q,3] (tet-barycentre queries, row N4)."""
        pos, nrm, area, query = _c64(pos), _c64(nrm), _c64(area), _c64(query)
        Y = np.empty((len(query), 3), dtype=np.float32)
        self._check(lib().shm3d_step12_points(self._h, float(lambda_), len(area), _dp(pos), _dp(nrm), _dp(area),
                                              len(query), _dp(query), _fp(Y)))
        return Y

    def rhs(self, p: Params, Y):
        Y = np.ascontiguousarray(Y, dtype=np.float32)
        b = np.empty(self.local_n(p), dtype=np.float32)
        self._check(lib().shm3d_rhs(self._h, C.byref(p), _fp(Y), _fp(b)))
        return b

    def step3(self, p: Params, pos, area, b):
        pos, area = _c64(pos), _c64(area)
        b = np.ascontiguousarray(b, dtype=np.float32)
        phi = np.empty(self.local_n(p))
        st = Stats()
        self._check(lib().shm3d_step3(self._h, C.byref(p), len(area), _dp(pos), _dp(area), _fp(b), _dp(phi),
                                      C.byref(st)))
        return phi, st

    # ---- row N3: the consumer of phi on the device
    @staticmethod
    def _field(p: Params, phi):
        """numpy float64 / float32 array (host) or an int device pointer to float32[N] -> (pointer, kind, keep-alive)"""
        if isinstance(phi, (int, np.integer)):
            return C.c_void_p(int(phi)), FIELD_DEVICE_F32, None
        a = np.asarray(phi)
        if a.dtype == np.float32:
            a = np.ascontiguousarray(a).ravel()
            kind = FIELD_HOST_F32
        else:
            a = np.ascontiguousarray(a, dtype=np.float64).ravel()
            kind = FIELD_HOST_F64
        if a.size != p.N:
            raise ValueError("field size does not match the grid")
        return C.c_void_p(a.ctypes.data), kind, a

    def isosurface(self, p: Params, phi, isoval=0.0, bound_min=None, bound_max=None, lattice=False, fetch=True):
        """registerIsosurfaceAsMesh on the GPU: (vertices float32[nV,3], triangles uint32[nT,3], IsoStats), numbered and
        ordered like MC::marching_cube's output.  phi: host array (float64 as computeDistance returns it, or float32)
        or a device pointer (int) to float32[N].  fetch=False leaves the mesh on the device (returns the stats only)."""
        ptr, kind, keep = self._field(p, phi)
        st = IsoStats()
        bm = None if bound_min is None else np.ascontiguousarray(bound_min, dtype=np.float32)
        bM = None if bound_max is None else np.ascontiguousarray(bound_max, dtype=np.float32)
        self._check(lib().shm3d_isosurface(self._h, C.byref(p), ptr, kind, float(isoval), None if bm is None else _fp(bm),
                                           None if bM is None else _fp(bM), ISO_LATTICE if lattice else 0, C.byref(st)))
        del keep
        if not fetch:
            return st
        V = np.empty((st.n_vertices, 3), dtype=np.float32)
        T = np.empty((st.n_triangles, 3), dtype=np.uint32)
        self._check(lib().shm3d_isosurface_fetch(self._h, _fp(V), T.ctypes.data_as(C.POINTER(C.c_uint32))))
        return V, T, st

    def isosurface_device(self):
        """device pointers (ints) of the last mesh: (vertices float32[nV,3], triangles uint32[nT,3])"""
        dv, dt = C.c_void_p(), C.c_void_p()
        self._check(lib().shm3d_isosurface_device(self._h, C.byref(dv), C.byref(dt)))
        return dv.value, dt.value

    def slice(self, p: Params, phi, origin, du, dv, nu, nv):
        """float32[nv, nu]: trilinear interpolant of phi at origin + a*du + b*dv (NaN outside the grid)"""
        ptr, kind, keep = self._field(p, phi)
        o, u, v = _c64(origin), _c64(du), _c64(dv)
        out = np.empty((nv, nu), dtype=np.float32)
        self._check(lib().shm3d_slice(self._h, C.byref(p), ptr, kind, _dp(o), _dp(u), _dp(v), nu, nv, _fp(out)))
        del keep
        return out


    def debug_stencil_op(self, op, dims, k0, k1, in0, in1, pw, scal, use_tma, reps=0):
        """One PCG / V-cycle stencil operation on padded float32 vectors [(k1-k0)+2, ny, nx] (include/shm3d_grid.h)."""
        nx, ny, nz = dims
        a = np.ascontiguousarray(in0, dtype=np.float32)
        b = None if in1 is None else np.ascontiguousarray(in1, dtype=np.float32)
        w = None if pw is None else np.ascontiguousarray(pw, dtype=np.float32)
        sc = np.zeros(4)
        sc[:len(scal)] = scal
        o0, o1, red = np.empty_like(a), np.empty_like(a), np.zeros(2)
        ms = C.c_double(0.0)
        self._check(lib().shm3d_debug_stencil_op(self._h, op, nx, ny, nz, k0, k1, _fp(a), None if b is None else _fp(b),
                                                 None if w is None else _fp(w), _dp(sc), 1 if use_tma else 0, _fp(o0), _fp(o1),
                                                 _dp(red), reps, C.byref(ms)))
        return (o0, o1, red, ms.value) if reps else (o0, o1, red)


def slab_range(rank, world, nz):
    k0, k1 = C.c_int32(), C.c_int32()
    rc = lib().shm3d_slab_range(rank, world, nz, C.byref(k0), C.byref(k1))
    if rc != OK:
        raise Shm3dError(rc, "bad rank/world/nz")
    return k0.value, k1.value


def _point_at_nccl():
    """The library binds NCCL at run time: SHM3D_NCCL_LIB, else a libnccl already mapped into the process (torch's), else
    the loader's search path.  When none of that is set up, offer the copy that ships in the `nvidia-nccl` wheel next to
    torch -- found through the import system, not through a hard-coded path."""
    if os.environ.get("SHM3D_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for loc in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(loc, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["SHM3D_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def nccl_unique_id() -> bytes:
    _point_at_nccl()
    buf = C.create_string_buffer(128)
    rc = lib().shm3d_nccl_unique_id(C.cast(buf, C.c_void_p))
    if rc != OK:
        raise Shm3dError(rc, "ncclGetUniqueId failed (NCCL not loadable?)")
    return buf.raw


# ----------------------------------------------------------------------------------------------------
# Mirror of the reference's operator interface
# ----------------------------------------------------------------------------------------------------
@dataclass
class SignedHeat3DOptions:
    """include/signed_heat_3d.h:20-28 (levelSetConstraint / useCrouzeixRaviart are tet-solver options and are
    ignored by the grid solver, src/signed_heat_grid_solver.cpp:75)."""
    tCoef: float = 1.0
    hCoef: float = 0.0
    rebuild: bool = True
    scale: float = 2.0
    fastIntegration: bool = False


class SignedHeatGridSolver:
    """include/signed_heat_grid_solver.h:11-22.  computeDistance(V, faces) is the mesh overload,
    computeDistancePoints(P, N, areas, h) the point-cloud overload (the tufted-triangulation areas and mean edge
    length are the caller's: SURVEY.md section 8f row N1)."""

    def __init__(self, device=0, context: Context | None = None, reuse_output: bool = True):
        self.VERBOSE = False
        self.reference_underflow = True    # SHM3D_FLAG_FP64_UNDERFLOW (include/shm3d_grid.h): follow the reference's X.norm() underflow
        # True: computeDistance returns a view of a solver-owned page-locked buffer that the NEXT call overwrites
        # (fast D2H, no per-call 1 GB allocation); False: every call returns a freshly allocated array it owns.
        self.reuse_output = reuse_output
        self.ctx = context if context is not None else Context(device)
        self.params = None   # grid of the last solve (the reference caches nx, bbox, cellSize)
        self.stats = None
        self._out = None     # page-locked result buffer, reused while the grid size stays the same

    def _finish(self, p, pos, nrm, area, options):
        if self.VERBOSE:
            p.flags |= FLAG_VERBOSE
        if options.fastIntegration:
            p.flags |= FLAG_FAST
        if self.reference_underflow:
            p.flags |= FLAG_FP64_UNDERFLOW
        else:
            p.flags &= ~FLAG_FP64_UNDERFLOW
        n = self.ctx.local_n(p)
        if not self.reuse_output:
            phi, st = self.ctx.solve(p, pos, nrm, area)
            self.params, self.stats = p, st
            return phi
        if self._out is None or self._out.array.size != n:
            self._out = None
            self._out = PinnedArray(n)
        phi, st = self.ctx.solve(p, pos, nrm, area, out=self._out.array)
        self.params, self.stats = p, st
        return phi  # view of the solver-owned pinned buffer: valid until the next computeDistance call

    def computeDistance(self, V, faces, options: SignedHeat3DOptions = SignedHeat3DOptions()):
        if self.params is not None and not options.rebuild:
            # grid cached from the previous call (src/signed_heat_grid_solver.cpp:8): only lambda / sources change
            p_new, pos, nrm, area, _ = prepare_mesh(V, faces, options.tCoef, 0.0, options.scale)
            p = Params.from_buffer_copy(self.params)
            p.lambda_ = p_new.lambda_
        else:
            p, pos, nrm, area, _ = prepare_mesh(V, faces, options.tCoef, options.hCoef, options.scale)
        return self._finish(p, pos, nrm, area, options)

    def isosurface(self, phi, isoval=0.0):
        """What the reference's contour() does with PHI (src/main.cpp:116-128): the level set as an indexed mesh in world
        coordinates, identical to polyscope's registerIsosurfaceAsMesh on the float32-narrowed field."""
        if self.params is None:
            raise Shm3dError(ERR_INVALID_ARG, "isosurface: no grid yet (call computeDistance first)")
        V, T, st = self.ctx.isosurface(self.params, phi, isoval)
        self.iso_stats = st
        return V, T

    def computeDistancePoints(self, P, normals, areas=None, h=None,
                              options: SignedHeat3DOptions = SignedHeat3DOptions()):
        if areas is None or h is None:  # row N1 (partial): local-Delaunay weights instead of the caller's
            a, hh, _ = point_weights(P, normals)
            areas = a if areas is None else areas
            h = hh if h is None else h
        p = prepare_points(P, h, options.tCoef, options.hCoef, options.scale)
        return self._finish(p, _c64(P), _c64(normals), _c64(areas), options)
