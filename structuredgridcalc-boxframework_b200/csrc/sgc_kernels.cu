// sgc_kernels.cu — generic sm_100a kernels: batched region copy (ghost fill / pack / unpack /
// BaseFab::copy / physical BC), fill, and the table-driven 7-point sweeps that work for any fab
// geometry.  The bandwidth-tuned streaming sweep for uniform levels lives in sgc_stream.cu.
#include <cstdint>
#include <cstdlib>

#include "sgc_device.h"
#include "sgc_pipe.cuh"

namespace sgc {

// ---------------------------------------------------------------------------------------------
// Batched region copy: ONE launch runs every item of a table (all motion items of
// an exchange are independent and race-free, so order does not matter and the result is
// bit-identical to the reference's serial loop, LevelData.H:471-486).
//
// Work split: the items' cells are numbered consecutively (prefix sum cell_start[]); CTA c owns
// cells [c*kCopyChunk, (c+1)*kCopyChunk) of its class (16-byte "wide" items first, see CopyTable in
// sgc_device.h).  chunk_first[c] is the first item overlapping the
// chunk, so a thread finds its item with a short binary search over a handful of entries (one
// entry for a 64x64 face, up to kCopyChunk entries when the chunk is all single-cell corners).
// Balanced regardless of the face/edge/corner mix (C3: 851 968 items from 1 to 256 cells).
// ---------------------------------------------------------------------------------------------
// Source loads ask L2 for the smallest DRAM fetch (64 B): an x-face item reads one element per
// row, and with the default 128-byte fetch ncu shows 583 MB read for 104 MB of ghost cells on C2.
template <typename T>
__device__ __forceinline__ T ld_small_fetch(const T* p) { return *p; }
template <>
__device__ __forceinline__ uint64_t ld_small_fetch<uint64_t>(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.global.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// one unit (an element, or 16 bytes of a wide item) per thread and iteration.  Kept deliberately small (32 registers,
// eight CTAs per SM): variants that decode a single-item chunk once and batch its eight loads, or batch the searches
// of multi-item chunks, execute a third of the instructions but run at 72 registers and were never faster on the
// exchange and up to 1.7x slower on tables of small items (profiles/r02_copy_variants.log) — occupancy, not instruction
// count, is what hides this kernel's dependent loads.
template <typename T, bool SMALL_FETCH_OK>
__device__ __forceinline__ void copy_units(const CopyItem* __restrict__ items, const long long* __restrict__ cell_start,
                                           int it_lo, int it_hi, long long base, long long end, T* __restrict__ dst_arena,
                                           const T* __restrict__ src_arena, int small_fetch) {
#pragma unroll 2
  for (int t = threadIdx.x; t < kCopyChunk; t += kCopyThreads) {
    const long long cell = base + t;
    if (cell >= end) break;
    // largest it in [it_lo, it_hi] with cell_start[it] <= cell
    int lo = it_lo, hi = it_hi;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(cell_start + mid) <= cell) lo = mid; else hi = mid - 1;
    }
    const CopyItem* __restrict__ I = items + lo;
    unsigned r = (unsigned)(cell - __ldg(cell_start + lo));
    const unsigned nx = I->nx, ny = I->ny, nz = I->nz;
    const unsigned x = r % nx; r /= nx;
    const unsigned y = r % ny; r /= ny;
    const unsigned z = r % nz;
    const unsigned c = r / nz;
    const unsigned dc = I->c0 + c;
    if (dc < 32u && !((I->flags >> dc) & 1u)) continue;
    const long long d = I->dst + c * I->dcs + (long long)x * I->dsx + (long long)y * I->dsy + (long long)z * I->dsz;
    const long long s = I->src + c * I->scs + (long long)x * I->ssx + (long long)y * I->ssy + (long long)z * I->ssz;
    if (SMALL_FETCH_OK) dst_arena[d] = (small_fetch != 0) ? ld_small_fetch(src_arena + s) : src_arena[s];
    else dst_arena[d] = src_arena[s];
  }
}

// CTAs [0, nwchunk) run the wide items (16-byte units, 8-byte element tables only), the others the rest.
template <typename T>
__global__ void __launch_bounds__(kCopyThreads)
k_copy_items(const CopyItem* __restrict__ items, const long long* __restrict__ cell_start,
             const int* __restrict__ chunk_first, T* __restrict__ dst_arena,
             const T* __restrict__ src_arena, long long ncell, int getenv_small_fetch, int nwchunk, long long nwcell) {
  if (sizeof(T) == 8 && (int)blockIdx.x < nwchunk) {
    const unsigned ch = blockIdx.x;
    const long long base = (long long)ch * kCopyChunk;
    const long long end = base + kCopyChunk < nwcell ? base + kCopyChunk : nwcell;
    copy_units<ulonglong2, false>(items, cell_start, chunk_first[ch], chunk_first[ch + 1], base, end,
                                  reinterpret_cast<ulonglong2*>(dst_arena), reinterpret_cast<const ulonglong2*>(src_arena), 0);
    return;
  }
  const unsigned c = blockIdx.x - nwchunk;
  const long long base = nwcell + (long long)c * kCopyChunk;
  const long long end = base + kCopyChunk < ncell ? base + kCopyChunk : ncell;
  copy_units<T, true>(items, cell_start, chunk_first[nwchunk + c + 1], chunk_first[nwchunk + c + 2], base, end, dst_arena,
                      src_arena, getenv_small_fetch);
}

int launch_copy(const CopyTable& t, void* dst_arena, const void* src_arena, int esz, cudaStream_t s) {
  if (t.nitem == 0 || t.ncell == 0) return 0;
  static const char* sf_env = getenv("SGC_COPY_SMALL_FETCH");
  const int small_fetch = sf_env ? atoi(sf_env) : (t.strided_reads ? 1 : 0);
  if (t.nwchunk > 0 && (esz != 8 || (((uintptr_t)dst_arena | (uintptr_t)src_arena) & 15)))
    return fail(SGC_ERR_ARG, "copy table has 16-byte items but the buffers are not 16-byte aligned 8-byte arrays");
#define SGC_COPY_LAUNCH(T, NW, WC)                                                                                \
  k_copy_items<T><<<t.nchunk, kCopyThreads, 0, s>>>(t.d_items, t.d_cell_start, t.d_chunk_first, (T*)dst_arena,     \
                                                    (const T*)src_arena, t.ncell, small_fetch, NW, WC)
  switch (esz) {
    case 8: SGC_COPY_LAUNCH(uint64_t, t.nwchunk, t.nwcell); break;
    case 4: SGC_COPY_LAUNCH(uint32_t, 0, 0); break;
    case 1: SGC_COPY_LAUNCH(uint8_t, 0, 0); break;
    default:
      return fail(SGC_ERR_UNSUPPORTED, "element size must be 1, 4 or 8 bytes");
  }
#undef SGC_COPY_LAUNCH
  rt().launches++;
  SGC_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Fill (BaseFab::setVal, BaseFab.cpp:219-246)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_fill(T* __restrict__ p, long long n, T v) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

// LevelData::setVal(comp, val) over every fab of a level in one launch: blockIdx.y = fab, the main
// and xg blocks of components [c0, c0+nc) are contiguous ranges of the arena
template <typename T>
__global__ void k_fill_fabs(T* __restrict__ arena, const FabGeom* __restrict__ geom, int c0, int nc, T v) {
  const FabGeom g = geom[blockIdx.y];
  const long long nmain = g.cs * nc, nx = g.xcs * nc;
  T* pm = arena + g.off + c0 * g.cs;
  T* px = arena + g.xoff + c0 * g.xcs;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nmain + nx; i += stride) {
    if (i < nmain) pm[i] = v; else px[i - nmain] = v;
  }
}

int launch_fill_fabs(void* arena, const FabGeom* d_geom, int nbox, long long max_elems, int c0, int nc, int esz,
                     const void* value, cudaStream_t s) {
  if (nbox <= 0) return 0;
  const int threads = 256;
  long long want = (max_elems + threads - 1) / threads;
  const dim3 grid((unsigned)(want < 64 ? (want < 1 ? 1 : want) : 64), (unsigned)nbox, 1);
  switch (esz) {
    case 8: k_fill_fabs<uint64_t><<<grid, threads, 0, s>>>((uint64_t*)arena, d_geom, c0, nc, *(const uint64_t*)value); break;
    case 4: k_fill_fabs<uint32_t><<<grid, threads, 0, s>>>((uint32_t*)arena, d_geom, c0, nc, *(const uint32_t*)value); break;
    case 1: k_fill_fabs<uint8_t><<<grid, threads, 0, s>>>((uint8_t*)arena, d_geom, c0, nc, *(const uint8_t*)value); break;
    default: return fail(SGC_ERR_UNSUPPORTED, "element size must be 1, 4 or 8 bytes");
  }
  rt().launches++;
  SGC_CUDA(cudaGetLastError());
  return 0;
}

int launch_fill(void* p, long long n, int esz, const void* value, cudaStream_t s) {
  if (n <= 0) return 0;
  const int threads = 256;
  long long want = (n + threads - 1) / threads;
  const int blocks = (int)(want < 148LL * 16 ? want : 148LL * 16);
  switch (esz) {
    case 8: k_fill<uint64_t><<<blocks, threads, 0, s>>>((uint64_t*)p, n, *(const uint64_t*)value); break;
    case 4: k_fill<uint32_t><<<blocks, threads, 0, s>>>((uint32_t*)p, n, *(const uint32_t*)value); break;
    case 1: k_fill<uint8_t><<<blocks, threads, 0, s>>>((uint8_t*)p, n, *(const uint8_t*)value); break;
    default: return fail(SGC_ERR_UNSUPPORTED, "element size must be 1, 4 or 8 bytes");
  }
  rt().launches++;
  SGC_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Table-driven 7-point sweep, any geometry.  One CTA per work item (box, y0, z0) of a
// kTileY x kTileZ slab of columns; a thread owns column (x, y) and marches in z with the three
// z-neighbours rolling through registers (the "register-blocked sweep along the slow axis");
// x/y neighbours come from global memory through L1.  The arithmetic is written with explicit
// rounding intrinsics so nvcc cannot contract or re-associate anything.
// ---------------------------------------------------------------------------------------------
constexpr int kTileX = 64, kTileY = 4, kTileZ = 16;

// second difference in the reference's association: (u[+e] - 2u) + u[-e]; 2u is exact.
template <typename T>
__device__ __forceinline__ T d2(T p, T c, T m) {
  return Ar<T>::add(Ar<T>::sub(p, Ar<T>::mul((T)2, c)), m);
}

// The ops see the centre value and its six neighbours (xm, xp, ym, yp, zm, zp) and the auxiliary value.
template <typename T>
struct HeatOp {                       // unew = u + f*S
  T f; int fma;
  __device__ __forceinline__ T operator()(T c, T xm, T xp, T ym, T yp, T zm, T zp, T) const {
    const T S = Ar<T>::add(Ar<T>::add(d2(xp, c, xm), d2(yp, c, ym)), d2(zp, c, zm));
    return fma ? Ar<T>::fma(f, S, c) : Ar<T>::add(c, Ar<T>::mul(f, S));
  }
};
template <typename T>
struct LapOp {                        // niter x { b -= coef*Tx; b -= coef*Ty; b -= coef*Tz }
  T coef; int fma; int niter;
  __device__ __forceinline__ T operator()(T c, T xm, T xp, T ym, T yp, T zm, T zp, T b) const {
    const T tx = d2(xp, c, xm), ty = d2(yp, c, ym), tz = d2(zp, c, zm);
    const T mc = -coef;
    for (int it = 0; it < niter; ++it) {
      if (fma) { b = Ar<T>::fma(mc, tx, b); b = Ar<T>::fma(mc, ty, b); b = Ar<T>::fma(mc, tz, b); }
      else {
        b = Ar<T>::sub(b, Ar<T>::mul(coef, tx)); b = Ar<T>::sub(b, Ar<T>::mul(coef, ty));
        b = Ar<T>::sub(b, Ar<T>::mul(coef, tz));
      }
    }
    return b;
  }
};
template <typename T>
struct WaveOp {                       // r = 2u - um1; r += f*Tx; r += f*Ty; r += f*Tz   (or the MD_DIRSUM form)
  T f; int fma;                       // fma: bit 0 = fused multiply-add, bit 1 = SGC_WAVE_DIRSUM
  __device__ __forceinline__ T operator()(T c, T xm, T xp, T ym, T yp, T zm, T zp, T um1) const {
    const T tx = d2(xp, c, xm), ty = d2(yp, c, ym), tz = d2(zp, c, zm);
    T r = Ar<T>::sub(Ar<T>::mul((T)2, c), um1);         // 2*c is exact: contraction-invariant
    if (fma & 2) {                                      // shipped pencil form: (2u - um1) + f*(Tz + (Ty + Tx))
      const T sum = Ar<T>::add(tz, Ar<T>::add(ty, tx));
      return (fma & 1) ? Ar<T>::fma(f, sum, r) : Ar<T>::add(r, Ar<T>::mul(f, sum));
    }
    if (fma & 1) { r = Ar<T>::fma(f, tx, r); r = Ar<T>::fma(f, ty, r); r = Ar<T>::fma(f, tz, r); return r; }
    r = Ar<T>::add(r, Ar<T>::mul(f, tx)); r = Ar<T>::add(r, Ar<T>::mul(f, ty)); r = Ar<T>::add(r, Ar<T>::mul(f, tz));
    return r;
  }
};
// General weighted 7-point stencil: r = w0*c; r += w1*u[x-1]; r += w2*u[x+1]; r += w3*u[y-1]; r += w4*u[y+1];
// r += w5*u[z-1]; r += w6*u[z+1], in that order (each multiply-add fused when fma).  The loop form of the
// reference's device stencil test (gpuSandbox.cpp:120-151: B = A[-e_z] + A + A[+e_z] on every component).
template <typename T>
struct WeightOp {
  T w[7]; int fma;
  __device__ __forceinline__ T operator()(T c, T xm, T xp, T ym, T yp, T zm, T zp, T) const {
    const T v[6] = {xm, xp, ym, yp, zm, zp};
    T r = Ar<T>::mul(w[0], c);
#pragma unroll
    for (int k = 0; k < 6; ++k) r = fma ? Ar<T>::fma(w[k + 1], v[k], r) : Ar<T>::add(r, Ar<T>::mul(w[k + 1], v[k]));
    return r;
  }
};

// AUX: 0 none, 1 = read old value of out (read-modify-write), 2 = third level with in geometry.
// Fabs are in the split-x device layout: x neighbours of the first / last valid column of the
// input fab come from its xg block, everything else from the main block (fab_index()).
// blockIdx.y = component (the same component of the in, out and aux fabs).
template <typename T, class Op, int AUX>
__global__ void __launch_bounds__(kTileX * kTileY)
k_sweep7(const int* __restrict__ work, const FabGeom* __restrict__ gin, const FabGeom* __restrict__ gout,
         const T* __restrict__ in, T* __restrict__ out, const T* __restrict__ aux, Op op, int tile_z) {
  const int b = work[3 * blockIdx.x], y0 = work[3 * blockIdx.x + 1], z0 = work[3 * blockIdx.x + 2];
  const int comp = blockIdx.y;
  const FabGeom GI = gin[b], GO = gout[b];
  const int y = y0 + threadIdx.y;
  if (y >= GO.vn[1]) return;
  const int zend = min(z0 + tile_z, GO.vn[2]);
  // fab-relative coordinates of valid cell (0, y, z0) of this box in the in / out fabs
  const int ix0 = GO.vlo[0] - GI.lo[0], iy = GO.vlo[1] + y - GI.lo[1], iz0 = GO.vlo[2] + z0 - GI.lo[2];
  const int ox0 = GO.vlo[0] - GO.lo[0], oy = GO.vlo[1] + y - GO.lo[1], oz0 = GO.vlo[2] + z0 - GO.lo[2];
  for (int x = threadIdx.x; x < GO.vn[0]; x += kTileX) {
    const int ix = ix0 + x;
    // the centre column and its y / z neighbours live in one block of the in fab; strides there:
    const bool cmain = (ix >= GI.ngx) && (ix < GI.ngx + GI.mx);
    const long long sy = cmain ? GI.mx : 1, sz = cmain ? (long long)GI.mx * GI.n[1] : GI.n[1];
    const T* __restrict__ p = in + fab_index(GI, ix, iy, iz0, comp);
    const T* __restrict__ pxm = in + fab_index(GI, ix - 1, iy, iz0, comp);
    const T* __restrict__ pxp = in + fab_index(GI, ix + 1, iy, iz0, comp);
    const bool mmain = (ix - 1 >= GI.ngx) && (ix - 1 < GI.ngx + GI.mx), pmain = (ix + 1 >= GI.ngx) && (ix + 1 < GI.ngx + GI.mx);
    const long long szm = mmain ? (long long)GI.mx * GI.n[1] : GI.n[1], szp = pmain ? (long long)GI.mx * GI.n[1] : GI.n[1];
    T* __restrict__ q = out + fab_index(GO, ox0 + x, oy, oz0, comp);
    const bool omain = (ox0 + x >= GO.ngx) && (ox0 + x < GO.ngx + GO.mx);
    const long long osz = omain ? (long long)GO.mx * GO.n[1] : GO.n[1];
    const T* __restrict__ a = (AUX == 2) ? aux + fab_index(GI, ix, iy, iz0, comp) : nullptr;
    T zm = __ldg(p - sz), c = __ldg(p);
    for (int z = z0; z < zend; ++z) {
      const T zp = __ldg(p + sz);
      T av = (T)0;
      if (AUX == 1) av = *q;
      if (AUX == 2) { av = __ldg(a); a += sz; }
      *q = op(c, __ldg(pxm), __ldg(pxp), __ldg(p - sy), __ldg(p + sy), zm, zp, av);
      zm = c; c = zp;
      p += sz; pxm += szm; pxp += szp; q += osz;
    }
  }
}

template <typename T, class Op, int AUX>
static int launch_sweep(const int* d_work, int w0, int nwork, int ncomp, const FabGeom* gin, const FabGeom* gout,
                        const void* in, void* out, const void* aux, Op op, int tile_z, cudaStream_t s) {
  if (nwork <= 0) return 0;
  k_sweep7<T, Op, AUX><<<dim3(nwork, ncomp), dim3(kTileX, kTileY), 0, s>>>(d_work + 3 * (size_t)w0, gin, gout, (const T*)in,
                                                                           (T*)out, (const T*)aux, op, tile_z);
  rt().launches++;
  SGC_CUDA(cudaGetLastError());
  return 0;
}

// esz 8: double, 4: float.  Every component of the levels is swept.
int launch_heat_generic(const int* d_work, int w0, int nwork, int ncomp, int esz, const FabGeom* gin, const FabGeom* gout,
                        const void* in, void* out, double f, int fma, int tile_z, cudaStream_t s) {
  if (esz == 4) return launch_sweep<float, HeatOp<float>, 0>(d_work, w0, nwork, ncomp, gin, gout, in, out, nullptr, HeatOp<float>{(float)f, fma}, tile_z, s);
  return launch_sweep<double, HeatOp<double>, 0>(d_work, w0, nwork, ncomp, gin, gout, in, out, nullptr, HeatOp<double>{f, fma}, tile_z, s);
}
int launch_lap_generic(const int* d_work, int w0, int nwork, int ncomp, int esz, const FabGeom* gin, const FabGeom* gout,
                       const void* in, void* out, double coef, int fma, int niter, int tile_z, cudaStream_t s) {
  if (esz == 4) return launch_sweep<float, LapOp<float>, 1>(d_work, w0, nwork, ncomp, gin, gout, in, out, nullptr, LapOp<float>{(float)coef, fma, niter}, tile_z, s);
  return launch_sweep<double, LapOp<double>, 1>(d_work, w0, nwork, ncomp, gin, gout, in, out, nullptr, LapOp<double>{coef, fma, niter}, tile_z, s);
}
int launch_wave_generic(const int* d_work, int w0, int nwork, int ncomp, int esz, const FabGeom* gin, const FabGeom* gout,
                        const void* in, void* out, const void* um1, double f, int fma, int tile_z, cudaStream_t s) {
  if (esz == 4) return launch_sweep<float, WaveOp<float>, 2>(d_work, w0, nwork, ncomp, gin, gout, in, out, um1, WaveOp<float>{(float)f, fma}, tile_z, s);
  return launch_sweep<double, WaveOp<double>, 2>(d_work, w0, nwork, ncomp, gin, gout, in, out, um1, WaveOp<double>{f, fma}, tile_z, s);
}
int launch_weight_generic(const int* d_work, int w0, int nwork, int ncomp, int esz, const FabGeom* gin, const FabGeom* gout,
                          const void* in, void* out, const double w[7], int fma, int tile_z, cudaStream_t s) {
  if (esz == 4) {
    WeightOp<float> op;
    for (int k = 0; k < 7; ++k) op.w[k] = (float)w[k];
    op.fma = fma;
    return launch_sweep<float, WeightOp<float>, 0>(d_work, w0, nwork, ncomp, gin, gout, in, out, nullptr, op, tile_z, s);
  }
  WeightOp<double> op;
  for (int k = 0; k < 7; ++k) op.w[k] = w[k];
  op.fma = fma;
  return launch_sweep<double, WeightOp<double>, 0>(d_work, w0, nwork, ncomp, gin, gout, in, out, nullptr, op, tile_z, s);
}

int sweep_tile_y() { return kTileY; }
int sweep_tile_z() { return kTileZ; }


// ---------------------------------------------------------------------------------------------
// Whole-level layout conversion for uniform levels of 8-byte elements with one ghost column per side:
// reference order (rows of n0 = mx + 2 elements) <-> split-x arena (dense rows of mx elements + the two
// x-ghost strips).  One warp per row: the row is read once, fully coalesced, and every element goes to its
// place — the generic item kernel reads the two ghost columns as 8-byte elements of 528-byte rows (a 32-byte
// sector each).  Rows are numbered (fab, comp, k, j) in memory order of both layouts.
// ---------------------------------------------------------------------------------------------
template <bool TO_ARENA>
__global__ void __launch_bounds__(256)
k_convert_rows(double* __restrict__ arena, double* __restrict__ stage, long long nrows, int mx, int n1, int n2, int ncomp,
               long long xg_base) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const int n0 = mx + 2;
  const long long rows_per_fab = (long long)ncomp * n2 * n1;
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += warps) {
    const long long fab = r / rows_per_fab;
    const long long q = r - fab * rows_per_fab;            // (comp, k, j) flattened
    const int c = (int)(q / ((long long)n2 * n1));
    const long long kj = q - (long long)c * n2 * n1;       // k * n1 + j
    double* srow = stage + r * n0;                         // reference order: the rows follow each other
    double* mrow = arena + r * mx;                         // main blocks: same order, dense rows
    // xg block of the fab: [comp][side][k][j]
    double* xlo = arena + xg_base + (fab * ncomp + c) * (2LL * n2 * n1) + kj;
    double* xhi = xlo + (long long)n2 * n1;
    if (TO_ARENA) {
      for (int i = lane; i < n0; i += 32) {
        const double v = srow[i];
        if (i == 0) *xlo = v; else if (i == n0 - 1) *xhi = v; else mrow[i - 1] = v;
      }
    } else {
      for (int i = lane; i < n0; i += 32) srow[i] = i == 0 ? *xlo : (i == n0 - 1 ? *xhi : mrow[i - 1]);
    }
  }
}

// 0 = done, 1 = geometry not covered (caller uses the item tables)
int launch_convert_level(void* arena, void* stage, long long nrows, int mx, int n1, int n2, int ncomp, long long xg_base,
                         bool to_arena, cudaStream_t s) {
  const int wpb = 8;
  long long blocks = (nrows + wpb - 1) / wpb;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  if (to_arena) k_convert_rows<true><<<(unsigned)blocks, wpb * 32, 0, s>>>((double*)arena, (double*)stage, nrows, mx, n1, n2, ncomp, xg_base);
  else k_convert_rows<false><<<(unsigned)blocks, wpb * 32, 0, s>>>((double*)arena, (double*)stage, nrows, mx, n1, n2, ncomp, xg_base);
  rt().launches++;
  SGC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sgc
