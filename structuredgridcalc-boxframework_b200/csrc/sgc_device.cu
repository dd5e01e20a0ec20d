// sgc_device.cu — runtime, device-resident level data, exchange plans and the C ABI entry points
// of include/sgc_b200.h (everything except the host-only layout/plan builders in sgc_plan.cpp and
// the multi-rank transport in sgc_comm.cu).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sgc_device.h"

namespace sgc {

Runtime& rt() {
  static Runtime r;
  return r;
}

// kernels / launchers (sgc_kernels.cu, sgc_stream.cu)
int launch_heat_generic(const int*, int, int, int ncomp, int esz, const FabGeom*, const FabGeom*, const void*, void*, double, int,
                        int tile_z, cudaStream_t);
int launch_lap_generic(const int*, int, int, int ncomp, int esz, const FabGeom*, const FabGeom*, const void*, void*, double, int, int,
                       int tile_z, cudaStream_t);
int launch_wave_generic(const int*, int, int, int ncomp, int esz, const FabGeom*, const FabGeom*, const void*, void*, const void*,
                        double, int, int tile_z, cudaStream_t);
int launch_weight_generic(const int*, int, int, int ncomp, int esz, const FabGeom*, const FabGeom*, const void*, void*,
                          const double w[7], int, int tile_z, cudaStream_t);
int sweep_tile_y();
int sweep_tile_z();
// streaming sweep (returns 1 when it does not apply to this geometry, 0 when launched, <0 on error)
int launch_heat_stream(sgc_level_s* unew, sgc_level_s* u, double f, int fma, int box_begin, int box_end,
                       cudaStream_t s);
int launch_heat_stream2(sgc_level_s* unew, sgc_level_s* u, double f, int fma, int a0, int a1, int b0, int b1,
                        cudaStream_t s, const PushArgs* fuse);
// peer-memory halo session (sgc_comm.cu)
int comm_p2p_setup(sgc_plan_s* plan, sgc_level_s* a, sgc_level_s* b, sgc_level_s* third);
void comm_p2p_release(sgc_plan_s* plan);
int comm_p2p_signal(sgc_plan_s* plan, unsigned long long value, cudaStream_t s);
int comm_p2p_check(sgc_plan_s* plan);
int comm_all_min(int v, int* out);
int comm_stream_barrier(cudaStream_t s);
bool comm_p2p_can_pull(sgc_plan_s* p, sgc_level_s* l);
int comm_p2p_pull(sgc_plan_s* p, sgc_level_s* l, unsigned long long budget_ns, cudaStream_t s);
void comm_level_destroyed(sgc_level_s* l);
void comm_drain_deferred_frees();
bool heat_stream_covers(const sgc_level_s* unew, const sgc_level_s* u);
bool heat_t2_covers(const sgc_level_s* unew, const sgc_level_s* u);
int launch_build_xe(sgc_level_s* l, cudaStream_t s);
int launch_heat_t2(sgc_level_s* unew, sgc_level_s* u, double f, int fma, const int* d_nbr, const T2Args* remote,
                   cudaStream_t s);
int launch_wave_stream(sgc_level_s* unp1, sgc_level_s* un, sgc_level_s* unm1, double f, int fma, cudaStream_t s);
// multi-rank transport (sgc_comm.cu)
int comm_plan_setup(sgc_plan_s* plan, sgc_level_s* level, const Grid& g, unsigned periodic, unsigned trim);
void comm_plan_release(sgc_plan_s* plan);
int comm_exchange_begin(sgc_level_s* level, sgc_plan_s* plan);
int comm_exchange_end(sgc_level_s* level, sgc_plan_s* plan);

bool step_graph_ok(int nremote_a, int nremote_b) {
  return !rt().profiling && nremote_a == 0 && nremote_b == 0 && !getenv("SGC_NO_GRAPH");
}
int StepGraph::begin() {
  launches_before = rt().launches;
  SGC_CUDA(cudaStreamBeginCapture(rt().compute, cudaStreamCaptureModeThreadLocal));
  return 0;
}
int StepGraph::end() {
  cudaGraph_t g = nullptr;
  SGC_CUDA(cudaStreamEndCapture(rt().compute, &g));
  unit = rt().launches - launches_before;
  rt().launches = launches_before;                 // the captured launches have not run
  const cudaError_t e = cudaGraphInstantiate(&exec, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return fail(SGC_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
  return 0;
}
int StepGraph::replay(int times) {
  for (int i = 0; i < times; ++i) SGC_CUDA(cudaGraphLaunch(exec, rt().compute));
  rt().launches += unit * times;
  return 0;
}
StepGraph::~StepGraph() {
  if (exec) cudaGraphExecDestroy(exec);
}

ProfScope::ProfScope(int klass_, cudaStream_t s_) : klass(klass_), s(s_), slot(-1) {
  Runtime& r = rt();
  if (!r.profiling) return;
  ProfLog& L = r.prof[klass];
  if (L.used >= 8192) return;
  if (L.used == (int)L.a.size()) {
    cudaEvent_t ea, eb;
    if (cudaEventCreate(&ea) != cudaSuccess || cudaEventCreate(&eb) != cudaSuccess) return;
    L.a.push_back(ea); L.b.push_back(eb);
  }
  slot = L.used++;
  cudaEventRecord(L.a[slot], s);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(rt().prof[klass].b[slot], s);
}

void CopyTable::release() {
  if (d_items) cudaFree(d_items);
  if (d_cell_start) cudaFree(d_cell_start);
  if (d_chunk_first) cudaFree(d_chunk_first);
  d_items = nullptr; d_cell_start = nullptr; d_chunk_first = nullptr;
  nitem = nchunk = nwitem = nwchunk = 0; ncell = nwcell = 0; strided_reads = false;
}

// can the item run with 16-byte accesses?  (8-byte elements; every base pointer the tables are used with
// — level arenas, staging and pack buffers — is 256-byte aligned, checked in launch_copy)
static bool wide_ok(const CopyItem& c) {
  if (c.dsx != 1 || c.ssx != 1 || (c.nx & 1) || (c.dst & 1) || (c.src & 1)) return false;
  if ((c.ny > 1 && ((c.dsy | c.ssy) & 1)) || (c.nz > 1 && ((c.dsz | c.ssz) & 1))) return false;
  return c.nc == 1 || !((c.dcs | c.scs) & 1);
}

int CopyTable::upload(const std::vector<CopyItem>& in, int esz) {
  release();
  nitem = (int)in.size();
  if (nitem == 0) return 0;
  // Two classes, wide (16-byte units) first.  Items are independent, so their order is free:
  // within a class the items smaller than a chunk go FIRST (their chunks hold many items and are the slow
  // ones — each unit's item is found by a dependent binary search: at the front of the grid they overlap
  // with everything else instead of forming its tail; C2 exchange 67.5 -> 63.4 us), then one masked-out
  // padding item up to the next chunk boundary, then the big items in their original order (faces and planes
  // start on chunk boundaries, so a chunk of theirs lies inside one item and its search ends at once).
  auto units = [](const CopyItem& c) { return (long long)c.nx * c.ny * c.nz * c.nc; };
  static const char* nw_env = getenv("SGC_COPY_NO_WIDE");
  static const bool no_wide = nw_env && nw_env[0] == '1';
  static const char* ns_env = getenv("SGC_COPY_NO_SORT");
  static const bool no_sort = ns_env && ns_env[0] == '1';
  std::vector<CopyItem> items;
  items.reserve(in.size() + 2);
  for (int wide = 1; wide >= 0; --wide) {
    std::vector<CopyItem> small, big;
    for (const CopyItem& c : in) {
      const bool w = esz == 8 && !no_wide && wide_ok(c);
      if ((int)w != wide) continue;
      CopyItem u = c;
      if (w) { u.nx /= 2; u.dst /= 2; u.src /= 2; u.dsy /= 2; u.dsz /= 2; u.ssy /= 2; u.ssz /= 2; u.dcs /= 2; u.scs /= 2; }
      (units(u) >= kCopyChunk && !no_sort ? big : small).push_back(u);
    }
    long long n = 0;
    for (const CopyItem& c : small) n += units(c);
    if (!big.empty() && n % kCopyChunk) {
      CopyItem pad;
      memset(&pad, 0, sizeof pad);       // flags = 0, c0 = 0: every unit is masked out, nothing is touched
      pad.nx = (int)(kCopyChunk - n % kCopyChunk); pad.ny = pad.nz = pad.nc = 1;
      small.push_back(pad);
    }
    items.insert(items.end(), small.begin(), small.end());
    items.insert(items.end(), big.begin(), big.end());
    if (wide) nwitem = (int)items.size();
  }
  nitem = (int)items.size();
  std::vector<long long> start(nitem + 1);
  start[0] = 0;
  for (int i = 0; i < nitem; ++i) start[i + 1] = start[i] + units(items[i]);
  ncell = start[nitem];
  nwcell = start[nwitem];
  // cells read one per source row, rows at least 256 bytes apart (for 8-byte elements): worth asking
  // L2 for 64-byte DRAM fetches (measured: C2 exchange 0.119 -> 0.097 ms; no gain for 128-byte rows)
  long long lonely = 0;
  for (int i = nwitem; i < nitem; ++i) {
    const CopyItem& c = items[i];
    if (c.nx == 1 && c.ssy >= 32) lonely += (long long)c.ny * c.nz * c.nc;
  }
  strided_reads = lonely * 5 > ncell - nwcell;
  nwchunk = (int)((nwcell + kCopyChunk - 1) / kCopyChunk);
  const int nnchunk = (int)((ncell - nwcell + kCopyChunk - 1) / kCopyChunk);
  nchunk = nwchunk + nnchunk;
  // chunk c of a class covers units [base + c*kCopyChunk, ...): first item overlapping it, then one closing
  // entry per class (the last item of the class)
  std::vector<int> first(nchunk + 2);
  int it = 0;
  for (int c = 0; c < nwchunk; ++c) {
    const long long cell = (long long)c * kCopyChunk;
    while (it + 1 < nwitem && start[it + 1] <= cell) ++it;
    first[c] = it;
  }
  first[nwchunk] = nwitem > 0 ? nwitem - 1 : 0;
  it = nwitem;
  for (int c = 0; c < nnchunk; ++c) {
    const long long cell = nwcell + (long long)c * kCopyChunk;
    while (it + 1 < nitem && start[it + 1] <= cell) ++it;
    first[nwchunk + 1 + c] = it;
  }
  first[nchunk + 1] = nitem - 1;
  SGC_CUDA(cudaMalloc(&d_items, sizeof(CopyItem) * (size_t)nitem));
  SGC_CUDA(cudaMalloc(&d_cell_start, sizeof(long long) * (size_t)(nitem + 1)));
  SGC_CUDA(cudaMalloc(&d_chunk_first, sizeof(int) * (size_t)(nchunk + 2)));
  SGC_CUDA(cudaMemcpy(d_items, items.data(), sizeof(CopyItem) * (size_t)nitem, cudaMemcpyHostToDevice));
  SGC_CUDA(cudaMemcpy(d_cell_start, start.data(), sizeof(long long) * (size_t)(nitem + 1), cudaMemcpyHostToDevice));
  SGC_CUDA(cudaMemcpy(d_chunk_first, first.data(), sizeof(int) * (size_t)(nchunk + 2), cudaMemcpyHostToDevice));
  return 0;
}

FabGeom level_fab_geom(const IBox& valid, int ng, long long off, long long xoff) {
  FabGeom g;
  const IBox f = valid.grown(ng);
  for (int d = 0; d < 3; ++d) { g.lo[d] = f.lo[d]; g.n[d] = f.ext(d); g.vlo[d] = valid.lo[d]; g.vn[d] = valid.ext(d); }
  g.ngx = ng;
  g.mx = g.n[0] - 2 * ng;
  g.cs = (long long)g.mx * g.n[1] * g.n[2];
  g.xcs = (long long)2 * ng * g.n[1] * g.n[2];
  g.off = off;
  g.xoff = xoff;
  return g;
}

// A dense buffer holding `reg` in wire order comp->k->j->i, starting at element `off`.
FabGeom dense_geom(const IBox& reg, long long off) {
  FabGeom g;
  for (int d = 0; d < 3; ++d) { g.lo[d] = reg.lo[d]; g.n[d] = reg.ext(d); g.vlo[d] = reg.lo[d]; g.vn[d] = reg.ext(d); }
  g.ngx = 0;
  g.mx = g.n[0];
  g.cs = reg.size();
  g.xcs = 0;
  g.off = off;
  g.xoff = 0;
  return g;
}

bool fab_contains(const FabGeom& g, const IBox& r) {
  for (int d = 0; d < 3; ++d)
    if (r.lo[d] < g.lo[d] || r.hi[d] >= g.lo[d] + g.n[d]) return false;
  return true;
}

// Strides of the block that holds fab-relative x index i: main (1, mx, mx*n1) or xg (n2*n1, 1, n1).
static void block_strides(const FabGeom& g, int i, int& sx, int& sy, int& sz, long long& cs) {
  const int vi = i - g.ngx;
  if (vi >= 0 && vi < g.mx) { sx = 1; sy = g.mx; sz = g.mx * g.n[1]; cs = g.cs; }
  else { sx = g.n[2] * g.n[1]; sy = 1; sz = g.n[1]; cs = g.xcs; }
}

void make_items(std::vector<CopyItem>& out, const FabGeom& gd, const IBox& dreg, int dcomp, const FabGeom& gs,
                const IBox& sreg, int scomp, int nc, unsigned flags) {
  // cut the x range where either side changes block (low ghost columns | valid columns | high ghost columns)
  const int nx = dreg.ext(0);
  int cuts[6], ncut = 0;
  cuts[ncut++] = 0;
  const int dx0 = dreg.lo[0] - gd.lo[0], sx0 = sreg.lo[0] - gs.lo[0];   // fab-relative x of the region origin
  const int cand[4] = {gd.ngx - dx0, gd.ngx + gd.mx - dx0, gs.ngx - sx0, gs.ngx + gs.mx - sx0};
  for (int c : cand)
    if (c > 0 && c < nx) cuts[ncut++] = c;
  cuts[ncut++] = nx;
  for (int i = 1; i < ncut; ++i)                 // insertion sort of <= 6 cut positions
    for (int j = i; j > 0 && cuts[j] < cuts[j - 1]; --j) std::swap(cuts[j], cuts[j - 1]);
  for (int p = 0; p + 1 < ncut; ++p) {
    const int x0 = cuts[p], x1 = cuts[p + 1];
    if (x1 <= x0) continue;
    CopyItem c;
    c.nx = x1 - x0; c.ny = dreg.ext(1); c.nz = dreg.ext(2); c.nc = nc;
    block_strides(gd, dx0 + x0, c.dsx, c.dsy, c.dsz, c.dcs);
    block_strides(gs, sx0 + x0, c.ssx, c.ssy, c.ssz, c.scs);
    c.dst = fab_index(gd, dx0 + x0, dreg.lo[1] - gd.lo[1], dreg.lo[2] - gd.lo[2], dcomp);
    c.src = fab_index(gs, sx0 + x0, sreg.lo[1] - gs.lo[1], sreg.lo[2] - gs.lo[2], scomp);
    c.flags = flags;
    c.c0 = dcomp;
    out.push_back(c);
  }
}

static bool same_geometry(const sgc_level_s* a, const sgc_level_s* b) {
  if (a->nbox != b->nbox || a->ncomp != b->ncomp || a->ng != b->ng || a->esz != b->esz) return false;
  for (int i = 0; i < a->nbox; ++i)
    if (memcmp(&a->boxes[i], &b->boxes[i], sizeof(IBox)) != 0) return false;
  return true;
}

static bool plan_matches(const sgc_plan_s* p, const sgc_level_s* l) {
  if (p->nbox != l->nbox || p->ncomp != l->ncomp || p->ng != l->ng || p->esz != l->esz) return false;
  for (int i = 0; i < l->nbox; ++i)
    if (memcmp(&p->boxes[i], &l->boxes[i], sizeof(IBox)) != 0) return false;
  return true;
}

static int ensure_scratch(sgc_level_s* l, size_t bytes) {
  if (l->scratch_bytes >= bytes) return 0;
  if (l->d_scratch) cudaFree(l->d_scratch);
  l->d_scratch = nullptr; l->scratch_bytes = 0;
  SGC_CUDA(cudaMalloc(&l->d_scratch, bytes));
  l->scratch_bytes = bytes;
  return 0;
}

}  // namespace sgc

using namespace sgc;

extern "C" {

// ------------------------------------------------------------------------------- runtime
int sgc_init(int device) {
  Runtime& r = rt();
  if (r.ready && r.device == device) return 0;
  if (r.ready) return fail(SGC_ERR_ARG, "sgc_init: already bound to another device");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(SGC_ERR_CUDA, std::string("sgc_init: no CUDA device (") + cudaGetErrorString(e) +
                                  "); this library has no CPU fallback");
  if (device < 0 || device >= n) return fail(SGC_ERR_ARG, "sgc_init: device index out of range");
  SGC_CUDA(cudaSetDevice(device));
  cudaDeviceProp p;
  SGC_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major < 10)
    return fail(SGC_ERR_UNSUPPORTED, "sgc_init: kernels are built for sm_100a only; device is sm_" +
                                         std::to_string(p.major) + std::to_string(p.minor));
  r.sm_count = p.multiProcessorCount;
  // the comm stream (pack, NCCL send/recv) outranks the compute stream so that its few CTAs are
  // placed as soon as an SM has room, instead of queueing behind a persistent sweep
  int prio_lo = 0, prio_hi = 0;
  SGC_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  SGC_CUDA(cudaStreamCreateWithPriority(&r.compute, cudaStreamNonBlocking, prio_lo));
  SGC_CUDA(cudaStreamCreateWithPriority(&r.comm, cudaStreamNonBlocking, prio_hi));
  r.prio_hi = prio_hi;
  r.device = device;
  r.launches = 0;
  r.ready = true;
  return 0;
}

int sgc_shutdown(void) {
  Runtime& r = rt();
  if (!r.ready) return 0;
  cudaDeviceSynchronize();
  if (r.compute) cudaStreamDestroy(r.compute);
  if (r.comm) cudaStreamDestroy(r.comm);
  if (r.d_stage) cudaFree(r.d_stage);
  comm_drain_deferred_frees();
  for (ProfLog& L : r.prof)
    for (size_t i = 0; i < L.a.size(); ++i) { cudaEventDestroy(L.a[i]); cudaEventDestroy(L.b[i]); }
  r = Runtime();
  return 0;
}

int sgc_device_info(char* name, int cap, int* sm_count, size_t* hbm_bytes) {
  SGC_REQUIRE_RT();
  cudaDeviceProp p;
  SGC_CUDA(cudaGetDeviceProperties(&p, rt().device));
  if (name && cap > 0) { strncpy(name, p.name, cap - 1); name[cap - 1] = 0; }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (hbm_bytes) *hbm_bytes = p.totalGlobalMem;
  return 0;
}

int sgc_sync(void) {
  SGC_REQUIRE_RT();
  SGC_CUDA(cudaStreamSynchronize(rt().compute));
  SGC_CUDA(cudaStreamSynchronize(rt().comm));
  return 0;
}

long long sgc_launch_count(void) { return rt().launches; }

int sgc_profile_enable(int on) {
  SGC_REQUIRE_RT();
  rt().profiling = on != 0;
  return 0;
}

int sgc_profile_read(int klass, int* launches, double* total_ms) {
  SGC_REQUIRE_RT();
  if (klass < 0 || klass > 2) return fail(SGC_ERR_ARG, "sgc_profile_read: class must be 0, 1 or 2");
  ProfLog& L = rt().prof[klass];
  double sum = 0;
  for (int i = 0; i < L.used; ++i) {
    float ms = 0;
    SGC_CUDA(cudaEventSynchronize(L.b[i]));
    SGC_CUDA(cudaEventElapsedTime(&ms, L.a[i], L.b[i]));
    sum += ms;
  }
  if (launches) *launches = L.used;
  if (total_ms) *total_ms = sum;
  L.used = 0;
  return 0;
}

int sgc_event_create(sgc_event_t* ev) {
  SGC_REQUIRE_RT();
  if (!ev) return fail(SGC_ERR_ARG, "null event");
  sgc_event_s* e = new sgc_event_s;
  cudaError_t st = cudaEventCreate(&e->ev);
  if (st != cudaSuccess) { delete e; return fail(SGC_ERR_CUDA, cudaGetErrorString(st)); }
  *ev = e;
  return 0;
}
int sgc_event_destroy(sgc_event_t ev) {
  if (!ev) return 0;
  cudaEventDestroy(ev->ev);
  delete ev;
  return 0;
}
int sgc_event_record(sgc_event_t ev) {
  SGC_REQUIRE_RT();
  if (!ev) return fail(SGC_ERR_ARG, "null event");
  SGC_CUDA(cudaEventRecord(ev->ev, rt().compute));
  return 0;
}
int sgc_event_elapsed_ms(sgc_event_t start, sgc_event_t stop, float* ms) {
  SGC_REQUIRE_RT();
  if (!start || !stop || !ms) return fail(SGC_ERR_ARG, "null event");
  SGC_CUDA(cudaEventSynchronize(stop->ev));
  SGC_CUDA(cudaEventElapsedTime(ms, start->ev, stop->ev));
  return 0;
}

int sgc_host_alloc(size_t bytes, void** ptr) {
  SGC_REQUIRE_RT();
  if (!ptr) return fail(SGC_ERR_ARG, "null pointer");
  SGC_CUDA(cudaMallocHost(ptr, bytes));
  return 0;
}
int sgc_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
  return 0;
}

// ------------------------------------------------------------------------------- level data
int sgc_level_create(const int* boxes6, int nbox, int ncomp, int nghost, int elem_bytes, int global_index0,
                     sgc_level_t* level) {
  SGC_REQUIRE_RT();
  if (!boxes6 || !level || nbox < 1 || ncomp < 1 || nghost < 0)
    return fail(SGC_ERR_ARG, "sgc_level_create: bad argument");
  if (elem_bytes != 8 && elem_bytes != 4 && elem_bytes != 1)
    return fail(SGC_ERR_UNSUPPORTED, "sgc_level_create: elem_bytes must be 8, 4 or 1");
  sgc_level_s* l = new sgc_level_s;
  static unsigned long long next_generation = 0;
  l->generation = ++next_generation;
  l->nbox = nbox; l->ncomp = ncomp; l->ng = nghost; l->esz = elem_bytes; l->gidx0 = global_index0;
  l->boxes.resize(nbox); l->off.resize(nbox); l->geom.resize(nbox);
  long long off = 0, moff = 0, xoff = 0;
  for (int b = 0; b < nbox; ++b) {
    const IBox v = IBox::from6(boxes6 + 6 * b);
    if (v.empty()) { delete l; return fail(SGC_ERR_ARG, "sgc_level_create: empty box"); }
    l->boxes[b] = v;
    l->off[b] = off;
    FabGeom& g = l->geom[b];
    g = level_fab_geom(v, nghost, moff, xoff);           // xoff is rebased below
    const long long fab_elems = (long long)g.n[0] * g.n[1] * g.n[2];
    if (fab_elems * ncomp > 0x7fffffffLL) {   // the reference indexes with int; we keep its limit per fab
      delete l;
      return fail(SGC_ERR_UNSUPPORTED, "sgc_level_create: a single fab exceeds 2^31 elements");
    }
    if (b && memcmp(l->geom[0].n, g.n, sizeof g.n) != 0) l->uniform = false;
    off += fab_elems * ncomp;
    moff += g.cs * ncomp;
    xoff += g.xcs * ncomp;
  }
  l->elems = off;
  // Arena (split-x layout, sgc_device.h): the main blocks of all fabs back to back in box order —
  // for a uniform level one homogeneous sequence of z-planes, which is what the streaming sweep
  // walks — then, 512-byte aligned, the xg blocks of all fabs.  32 elements of padding in front and
  // 64 behind keep bulk copies 16-byte aligned and give the kernels slack at both ends.
  l->lead = 32;
  l->main_elems = moff;
  l->xg_base = (moff + 63) / 64 * 64;
  l->arena_elems = l->xg_base + xoff;
  for (int b = 0; b < nbox; ++b) l->geom[b].xoff += l->xg_base;
  // edge strips for the two-step sweep (sgc_stream2.cu), behind the arena proper
  l->xe_base = (l->arena_elems + 63) / 64 * 64;
  {
    const FabGeom& g0 = l->geom[0];
    if (l->uniform && ncomp == 1 && nghost == 1 && elem_bytes == 8 && ((g0.vn[0] == 64 && g0.vn[1] == 64) || (g0.vn[0] == 32 && g0.vn[1] == 32) ||
         (g0.vn[0] == 16 && g0.vn[1] == 16 && getenv("SGC_T2_16"))) && g0.vn[2] >= 4)
      l->xe_elems = ((long long)nbox * g0.vn[2] + 1) * 4 * g0.vn[1] + (long long)g0.vn[0] * g0.vn[1];   // + a spare strip plane + a scratch plane
  }
  const size_t bytes = (size_t)(l->lead + l->xe_base + l->xe_elems + 64) * elem_bytes;
  cudaError_t e = cudaMalloc(&l->d_base, bytes);
  if (e != cudaSuccess) { delete l; return fail(SGC_ERR_CUDA, std::string("cudaMalloc level arena: ") + cudaGetErrorString(e)); }
  cudaMemsetAsync(l->d_base, 0, bytes, rt().compute);
  l->d_arena = (char*)l->d_base + l->lead * elem_bytes;
  {
    // conversion tables: reference-order staging (fab b at element off[b]) <-> arena
    std::vector<CopyItem> cin, cout;
    for (int b = 0; b < nbox; ++b) {
      const IBox f = l->boxes[b].grown(nghost);
      const FabGeom dg = dense_geom(f, l->off[b]);
      make_items(cin, l->geom[b], f, 0, dg, f, 0, ncomp, ~0u);
      make_items(cout, dg, f, 0, l->geom[b], f, 0, ncomp, ~0u);
    }
    int st = l->conv_in.upload(cin, elem_bytes);
    if (!st) st = l->conv_out.upload(cout, elem_bytes);
    if (st) { sgc_level_destroy(l); return st; }
  }
  e = cudaMalloc(&l->d_geom, sizeof(FabGeom) * (size_t)nbox);
  if (e == cudaSuccess) e = cudaMemcpy(l->d_geom, l->geom.data(), sizeof(FabGeom) * (size_t)nbox, cudaMemcpyHostToDevice);
  // sweep work table for the generic kernel: (box, y0, z0) tiles of the valid region
  std::vector<int> work;
  l->work_begin.resize(nbox + 1);
  // z extent of a work tile: 16 planes, halved while the level would give the table-driven kernel fewer than four
  // CTAs per SM (the reference driver's single 128^3 box: 256 tiles of 16 planes leave 40 SMs half idle)
  const int ty = sweep_tile_y();
  int tz = sweep_tile_z();
  {
    auto count = [&](int z) { long long n = 0; for (int b = 0; b < nbox; ++b) n += (long long)((l->geom[b].vn[2] + z - 1) / z) * ((l->geom[b].vn[1] + ty - 1) / ty); return n * ncomp; };
    while (tz > 2 && count(tz) < 4LL * rt().sm_count) tz /= 2;
  }
  l->tile_z = tz;
  for (int b = 0; b < nbox; ++b) {
    l->work_begin[b] = (int)(work.size() / 3);
    for (int z = 0; z < l->geom[b].vn[2]; z += tz)
      for (int y = 0; y < l->geom[b].vn[1]; y += ty) { work.push_back(b); work.push_back(y); work.push_back(z); }
  }
  l->nwork = (int)(work.size() / 3);
  l->work_begin[nbox] = l->nwork;
  if (e == cudaSuccess) e = cudaMalloc(&l->d_work, sizeof(int) * work.size());
  if (e == cudaSuccess) e = cudaMemcpy(l->d_work, work.data(), sizeof(int) * work.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    sgc_level_destroy(l);
    return fail(SGC_ERR_CUDA, std::string("sgc_level_create: ") + cudaGetErrorString(e));
  }
  *level = l;
  return 0;
}

int sgc_level_destroy(sgc_level_t l) {
  if (!l) return 0;
  if (rt().ready) cudaStreamSynchronize(rt().compute);
  comm_level_destroyed(l);             // peer-memory sessions that name this level end here
  l->conv_in.release();
  l->conv_out.release();
  l->conv_valid.release();
  l->conv_valid_in.release();
  // an arena other GPUs may still have mapped is freed at the next point every rank passes (sgc_comm_finalize)
  if (l->d_base) { if (l->exported && rt().ready) rt().deferred_free.push_back(l->d_base); else cudaFree(l->d_base); }
  if (l->d_geom) cudaFree(l->d_geom);
  if (l->d_work) cudaFree(l->d_work);
  if (l->d_scratch) cudaFree(l->d_scratch);
  delete l;
  return 0;
}

long long sgc_level_elems(sgc_level_t l) { return l ? l->elems : -1; }
int sgc_level_edge_strips(sgc_level_t l) { return (!l || !l->xe_elems) ? -1 : l->xe_valid; }
long long sgc_level_fab_offset(sgc_level_t l, int box) {
  if (!l || box < 0 || box >= l->nbox) return -1;
  return l->off[box];
}
void* sgc_level_device_ptr(sgc_level_t l, int box) {
  if (!l || box < 0 || box >= l->nbox) return nullptr;
  l->xe_extern = true;           // the caller may write valid cells behind the edge strips' back from now on
  l->xe_valid = 0;
  return (char*)l->d_arena + l->geom[box].off * l->esz;     // main block of the fab (split-x layout)
}

// whole level between a reference-order staging buffer and the arena: the row kernel for uniform levels of
// 8-byte elements with one ghost column per side, the item tables otherwise
static int convert_level(sgc_level_s* l, void* stage, bool to_arena, cudaStream_t s) {
  const FabGeom& g = l->geom[0];
  if (l->uniform && l->esz == 8 && l->ng >= 1 && g.ngx == 1 && g.mx >= 1)
    return launch_convert_level(l->d_arena, stage, (long long)l->nbox * l->ncomp * g.n[2] * g.n[1], g.mx, g.n[1], g.n[2],
                                l->ncomp, l->xg_base, to_arena, s);
  return to_arena ? launch_copy(l->conv_in, l->d_arena, stage, l->esz, s) : launch_copy(l->conv_out, stage, l->d_arena, l->esz, s);
}

static int ensure_stage(size_t bytes) {
  Runtime& r = rt();
  if (r.stage_bytes >= bytes) return 0;
  SGC_CUDA(cudaStreamSynchronize(r.compute));
  if (r.d_stage) cudaFree(r.d_stage);
  r.d_stage = nullptr; r.stage_bytes = 0;
  SGC_CUDA(cudaMalloc(&r.d_stage, bytes));
  r.stage_bytes = bytes;
  return 0;
}

// one fab <-> dense reference-order buffer at element 0 of `buf`
static int convert_box(sgc_level_s* l, int box, void* buf, bool to_device) {
  const IBox f = l->boxes[box].grown(l->ng);
  const FabGeom dg = dense_geom(f, 0);
  std::vector<CopyItem> v;
  if (to_device) make_items(v, l->geom[box], f, 0, dg, f, 0, l->ncomp, ~0u);
  else make_items(v, dg, f, 0, l->geom[box], f, 0, l->ncomp, ~0u);
  CopyTable t;
  int st = t.upload(v, l->esz);
  if (to_device) level_written(l);
  if (!st) st = to_device ? launch_copy(t, l->d_arena, buf, l->esz, rt().compute)
                          : launch_copy(t, buf, l->d_arena, l->esz, rt().compute);
  if (!st) cudaStreamSynchronize(rt().compute);
  t.release();
  return st;
}

int sgc_level_upload(sgc_level_t l, int box, const void* host) {
  SGC_REQUIRE_RT();
  if (!l || !host || box < -1 || box >= l->nbox) return fail(SGC_ERR_ARG, "sgc_level_upload: bad argument");
  const long long n = box < 0 ? l->elems : (long long)l->geom[box].n[0] * l->geom[box].n[1] * l->geom[box].n[2] * l->ncomp;
  int st = ensure_stage((size_t)n * l->esz);
  if (st) return st;
  // H2D in the reference's memory order, then one batched kernel into the split-x arena
  SGC_CUDA(cudaMemcpyAsync(rt().d_stage, host, (size_t)n * l->esz, cudaMemcpyHostToDevice, rt().compute));
  level_written(l);
  if (box < 0) return convert_level(l, rt().d_stage, true, rt().compute);
  return convert_box(l, box, rt().d_stage, true);
}

int sgc_level_download(sgc_level_t l, int box, void* host) {
  SGC_REQUIRE_RT();
  if (!l || !host || box < -1 || box >= l->nbox) return fail(SGC_ERR_ARG, "sgc_level_download: bad argument");
  const long long n = box < 0 ? l->elems : (long long)l->geom[box].n[0] * l->geom[box].n[1] * l->geom[box].n[2] * l->ncomp;
  int st = ensure_stage((size_t)n * l->esz);
  if (st) return st;
  st = box < 0 ? convert_level(l, rt().d_stage, false, rt().compute)
               : convert_box(l, box, rt().d_stage, false);
  if (st) return st;
  SGC_CUDA(cudaMemcpyAsync(host, rt().d_stage, (size_t)n * l->esz, cudaMemcpyDeviceToHost, rt().compute));
  SGC_CUDA(cudaStreamSynchronize(rt().compute));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Transfer streams (LevelData::copyToDeviceAsync / copyToHostAsync, LevelData.H:734-765)
// ---------------------------------------------------------------------------------------------
int sgc_stream_create(sgc_stream_t* stream) {
  SGC_REQUIRE_RT();
  if (!stream) return fail(SGC_ERR_ARG, "null stream");
  sgc_stream_s* x = new sgc_stream_s;
  // like the comm stream: its conversion kernels should slip in between persistent sweeps
  cudaError_t e = cudaStreamCreateWithPriority(&x->s, cudaStreamNonBlocking, rt().prio_hi);
  if (e != cudaSuccess) { sgc_stream_destroy(x); return fail(SGC_ERR_CUDA, cudaGetErrorString(e)); }
  *stream = x;
  return 0;
}

int sgc_stream_destroy(sgc_stream_t x) {
  if (!x) return 0;
  if (x->s) { cudaStreamSynchronize(x->s); cudaStreamDestroy(x->s); }
  if (x->d_stage) cudaFree(x->d_stage);
  delete x;
  return 0;
}

int sgc_event_record_stream(sgc_event_t ev, sgc_stream_t x) {
  SGC_REQUIRE_RT();
  if (!ev) return fail(SGC_ERR_ARG, "null event");
  SGC_CUDA(cudaEventRecord(ev->ev, x ? x->s : rt().compute));
  return 0;
}

int sgc_stream_wait_event(sgc_stream_t x, sgc_event_t ev) {
  SGC_REQUIRE_RT();
  if (!ev) return fail(SGC_ERR_ARG, "null event");
  SGC_CUDA(cudaStreamWaitEvent(x ? x->s : rt().compute, ev->ev, 0));
  return 0;
}

int sgc_stream_sync(sgc_stream_t x) {
  SGC_REQUIRE_RT();
  SGC_CUDA(cudaStreamSynchronize(x ? x->s : rt().compute));
  return 0;
}

static int stream_stage(sgc_stream_s* x, size_t bytes) {
  if (x->stage_bytes >= bytes) return 0;
  SGC_CUDA(cudaStreamSynchronize(x->s));
  if (x->d_stage) cudaFree(x->d_stage);
  x->d_stage = nullptr; x->stage_bytes = 0;
  SGC_CUDA(cudaMalloc(&x->d_stage, bytes));
  x->stage_bytes = bytes;
  return 0;
}

int sgc_level_upload_async(sgc_level_t l, const void* host, sgc_stream_t x) {
  SGC_REQUIRE_RT();
  if (!l || !host || !x) return fail(SGC_ERR_ARG, "sgc_level_upload_async: bad argument");
  const size_t bytes = (size_t)l->elems * l->esz;
  int st = stream_stage(x, bytes);
  if (st) return st;
  SGC_CUDA(cudaMemcpyAsync(x->d_stage, host, bytes, cudaMemcpyHostToDevice, x->s));
  level_written(l);
  return convert_level(l, x->d_stage, true, x->s);
}

int sgc_level_download_async(sgc_level_t l, void* host, sgc_stream_t x, sgc_event_t gathered) {
  SGC_REQUIRE_RT();
  if (!l || !host || !x) return fail(SGC_ERR_ARG, "sgc_level_download_async: bad argument");
  const size_t bytes = (size_t)l->elems * l->esz;
  int st = stream_stage(x, bytes);
  if (!st) st = convert_level(l, x->d_stage, false, x->s);
  if (st) return st;
  if (gathered) SGC_CUDA(cudaEventRecord(gathered->ev, x->s));
  SGC_CUDA(cudaMemcpyAsync(host, x->d_stage, bytes, cudaMemcpyDeviceToHost, x->s));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Plot gather: valid cells only (LevelData::writeCGNSSolData, LevelData.cpp:150-193)
// ---------------------------------------------------------------------------------------------
long long sgc_level_valid_elems(sgc_level_t l) {
  if (!l) return -1;
  long long n = 0;
  for (const IBox& b : l->boxes) n += b.size() * l->ncomp;
  return n;
}

static int ensure_valid_table(sgc_level_s* l) {
  if (l->conv_valid.d_items || l->nbox == 0) return 0;
  std::vector<CopyItem> v, w;
  long long off = 0;
  for (int b = 0; b < l->nbox; ++b) {
    const IBox& vb = l->boxes[b];
    make_items(v, dense_geom(vb, off), vb, 0, l->geom[b], vb, 0, l->ncomp, ~0u);
    make_items(w, l->geom[b], vb, 0, dense_geom(vb, off), vb, 0, l->ncomp, ~0u);
    off += vb.size() * l->ncomp;
  }
  l->valid_elems = off;
  const int st = l->conv_valid.upload(v, l->esz);
  return st ? st : l->conv_valid_in.upload(w, l->esz);
}

// Valid cells of a UNIFORM level straight between a host buffer [box][comp][k][j][i] and the arena, by the copy
// engine alone: in the split-x layout the valid rows of one fab plane are ONE contiguous block (vny rows of mx
// elements) at the plane pitch mx*n1, planes repeat at that pitch, (box, comp) blocks at cs — exactly a pitched 3-D
// copy (x = the plane's valid block in bytes, y = the valid planes, z = box*ncomp + comp).  No staging buffer, no
// conversion kernel, nothing that competes with the sweeps for SMs or HBM beyond the bytes themselves.  (The e2e pipeline
// of bench.py did not get faster with it — 27.1 ms per C2 batch either way, against 23.6 ms for the two copies alone and
// 24.4 ms for the sweeps alone: with the sweeps running the copy engines get ~39 instead of 45 GB/s per direction, and
// that, not a conversion kernel, is what the pipeline waits for.)  Returns 1 when the level is not uniform (the item
// tables serve it).
static int valid_cells_dma(sgc_level_s* l, void* host, bool to_device, cudaStream_t s) {
  if (!l->uniform || l->nbox == 0 || getenv("SGC_NO_DMA3D")) return 1;
  const FabGeom& g = l->geom[0];
  const size_t row = (size_t)g.mx * g.vn[1] * l->esz;                 // one plane's valid block
  const size_t pitch = (size_t)g.mx * g.n[1] * l->esz;                // plane pitch in the arena
  cudaMemcpy3DParms p;
  memset(&p, 0, sizeof p);
  cudaPitchedPtr hp = make_cudaPitchedPtr(host, row, row, (size_t)g.vn[2]);
  cudaPitchedPtr dp = make_cudaPitchedPtr(l->d_arena, pitch, pitch, (size_t)g.n[2]);
  const cudaPos hpos = make_cudaPos(0, 0, 0);
  const cudaPos dpos = make_cudaPos((size_t)l->ng * g.mx * l->esz, (size_t)l->ng, 0);     // skip the ghost rows / planes
  p.srcPtr = to_device ? hp : dp;
  p.dstPtr = to_device ? dp : hp;
  p.srcPos = to_device ? hpos : dpos;
  p.dstPos = to_device ? dpos : hpos;
  p.extent = make_cudaExtent(row, (size_t)g.vn[2], (size_t)l->nbox * l->ncomp);
  p.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  SGC_CUDA(cudaMemcpy3DAsync(&p, s));
  return 0;
}

// The reverse of the plot gather: the valid cells of every fab from a dense host buffer [box][comp][k][j][i];
// ghost cells keep their contents (the next exchange / BC fill writes them).  9-27 % fewer bytes over PCIe than
// the whole-fab copyToDevice of BaseFab.cpp:427-483 (66^3 -> 64^3: 8.8 %; 18^3 -> 16^3: 30 %).
int sgc_level_upload_valid(sgc_level_t l, const void* host) {
  SGC_REQUIRE_RT();
  if (!l || !host) return fail(SGC_ERR_ARG, "sgc_level_upload_valid: bad argument");
  const size_t bytes = (size_t)sgc_level_valid_elems(l) * l->esz;
  level_written(l);
  int st = valid_cells_dma(l, const_cast<void*>(host), true, rt().compute);
  if (st <= 0) return st;
  st = ensure_valid_table(l);
  if (!st) st = ensure_stage(bytes);
  if (st) return st;
  SGC_CUDA(cudaMemcpyAsync(rt().d_stage, host, bytes, cudaMemcpyHostToDevice, rt().compute));
  return launch_copy(l->conv_valid_in, l->d_arena, rt().d_stage, l->esz, rt().compute);
}

int sgc_level_upload_valid_async(sgc_level_t l, const void* host, sgc_stream_t x) {
  SGC_REQUIRE_RT();
  if (!l || !host || !x) return fail(SGC_ERR_ARG, "sgc_level_upload_valid_async: bad argument");
  const size_t bytes = (size_t)sgc_level_valid_elems(l) * l->esz;
  level_written(l);
  int st = valid_cells_dma(l, const_cast<void*>(host), true, x->s);
  if (st <= 0) return st;
  st = ensure_valid_table(l);
  if (!st) st = stream_stage(x, bytes);
  if (st) return st;
  SGC_CUDA(cudaMemcpyAsync(x->d_stage, host, bytes, cudaMemcpyHostToDevice, x->s));
  return launch_copy(l->conv_valid_in, l->d_arena, x->d_stage, l->esz, x->s);
}

int sgc_level_download_valid(sgc_level_t l, void* host) {
  SGC_REQUIRE_RT();
  if (!l || !host) return fail(SGC_ERR_ARG, "sgc_level_download_valid: bad argument");
  int st = valid_cells_dma(l, host, false, rt().compute);
  if (st < 0) return st;
  if (st == 0) { SGC_CUDA(cudaStreamSynchronize(rt().compute)); return 0; }
  st = ensure_valid_table(l);
  if (!st) st = ensure_stage((size_t)sgc_level_valid_elems(l) * l->esz);
  if (!st) st = launch_copy(l->conv_valid, rt().d_stage, l->d_arena, l->esz, rt().compute);
  if (st) return st;
  SGC_CUDA(cudaMemcpyAsync(host, rt().d_stage, (size_t)sgc_level_valid_elems(l) * l->esz, cudaMemcpyDeviceToHost, rt().compute));
  SGC_CUDA(cudaStreamSynchronize(rt().compute));
  return 0;
}

int sgc_level_download_valid_async(sgc_level_t l, void* host, sgc_stream_t x, sgc_event_t gathered) {
  SGC_REQUIRE_RT();
  if (!l || !host || !x) return fail(SGC_ERR_ARG, "sgc_level_download_valid_async: bad argument");
  const size_t bytes = (size_t)sgc_level_valid_elems(l) * l->esz;
  int st = valid_cells_dma(l, host, false, x->s);
  if (st < 0) return st;
  if (st == 0) {           // the copy engine reads the level itself: "gathered" = the copy has finished
    if (gathered) SGC_CUDA(cudaEventRecord(gathered->ev, x->s));
    return 0;
  }
  st = ensure_valid_table(l);
  if (!st) st = stream_stage(x, bytes);
  if (!st) st = launch_copy(l->conv_valid, x->d_stage, l->d_arena, l->esz, x->s);
  if (st) return st;
  if (gathered) SGC_CUDA(cudaEventRecord(gathered->ev, x->s));
  SGC_CUDA(cudaMemcpyAsync(host, x->d_stage, bytes, cudaMemcpyDeviceToHost, x->s));
  return 0;
}

int sgc_level_setval(sgc_level_t l, int box, int comp, const void* value) {
  SGC_REQUIRE_RT();
  if (!l || !value || box < -1 || box >= l->nbox || comp < -1 || comp >= l->ncomp)
    return fail(SGC_ERR_ARG, "sgc_level_setval: bad argument");
  const int b0 = box < 0 ? 0 : box, b1 = box < 0 ? l->nbox : box + 1;
  level_written(l);
  if (comp < 0 && box < 0) {
    int st = launch_fill(l->d_arena, l->main_elems, l->esz, value, rt().compute);
    if (!st && l->arena_elems > l->xg_base)
      st = launch_fill((char*)l->d_arena + l->xg_base * l->esz, l->arena_elems - l->xg_base, l->esz, value, rt().compute);
    return st;
  }
  if (box < 0) {
    // one component of every fab: one launch over the device box table
    long long mx = 0;
    for (const FabGeom& g : l->geom) mx = std::max(mx, g.cs + g.xcs);
    return launch_fill_fabs(l->d_arena, l->d_geom, l->nbox, mx, comp, 1, l->esz, value, rt().compute);
  }
  for (int b = b0; b < b1; ++b) {
    const FabGeom& g = l->geom[b];
    const int c0 = comp < 0 ? 0 : comp, nc = comp < 0 ? l->ncomp : 1;
    int st = launch_fill((char*)l->d_arena + (g.off + c0 * g.cs) * l->esz, g.cs * nc, l->esz, value, rt().compute);
    if (!st && g.xcs) st = launch_fill((char*)l->d_arena + (g.xoff + c0 * g.xcs) * l->esz, g.xcs * nc, l->esz, value, rt().compute);
    if (st) return st;
  }
  return 0;
}

int sgc_level_copy(sgc_level_t dst, sgc_level_t src) {
  SGC_REQUIRE_RT();
  if (!dst || !src) return fail(SGC_ERR_ARG, "sgc_level_copy: null level");
  if (!same_geometry(dst, src)) return fail(SGC_ERR_GEOMETRY, "sgc_level_copy: levels must share geometry");
  if (dst == src) return 0;
  level_written(dst);
  SGC_CUDA(cudaMemcpyAsync(dst->d_arena, src->d_arena, (size_t)dst->arena_elems * dst->esz, cudaMemcpyDeviceToDevice,
                           rt().compute));
  return 0;
}

int sgc_fab_copy(sgc_level_t dst, int dst_box, const int dlo[3], const int dhi[3], int dst_comp, sgc_level_t src,
                 int src_box, const int slo[3], const int shi[3], int src_comp, int ncomp, unsigned comp_flags) {
  SGC_REQUIRE_RT();
  if (!dst || !src || !dlo || !dhi || !slo || !shi) return fail(SGC_ERR_ARG, "sgc_fab_copy: null argument");
  if (dst_box < 0 || dst_box >= dst->nbox || src_box < 0 || src_box >= src->nbox)
    return fail(SGC_ERR_ARG, "sgc_fab_copy: box index out of range");
  if (dst->esz != src->esz) return fail(SGC_ERR_ARG, "sgc_fab_copy: element types differ");
  IBox dr, sr;
  for (int d = 0; d < 3; ++d) { dr.lo[d] = dlo[d]; dr.hi[d] = dhi[d]; sr.lo[d] = slo[d]; sr.hi[d] = shi[d]; }
  const FabGeom& gd = dst->geom[dst_box];
  const FabGeom& gs = src->geom[src_box];
  // the reference asserts these (BaseFab.cpp:306-312)
  if (dr.empty() || dr.ext(0) != sr.ext(0) || dr.ext(1) != sr.ext(1) || dr.ext(2) != sr.ext(2))
    return fail(SGC_ERR_ARG, "sgc_fab_copy: regions must be non-empty and of equal extents");
  if (!fab_contains(gd, dr) || !fab_contains(gs, sr)) return fail(SGC_ERR_ARG, "sgc_fab_copy: region outside fab");
  if (dst_comp < 0 || src_comp < 0 || ncomp < 1 || dst_comp + ncomp > dst->ncomp || src_comp + ncomp > src->ncomp)
    return fail(SGC_ERR_ARG, "sgc_fab_copy: component range");
  std::vector<CopyItem> v;
  make_items(v, gd, dr, dst_comp, gs, sr, src_comp, ncomp, comp_flags);
  CopyTable t;
  int st = t.upload(v, dst->esz);
  level_written(dst);
  if (!st) st = launch_copy(t, dst->d_arena, src->d_arena, dst->esz, rt().compute);
  if (!st) cudaStreamSynchronize(rt().compute);
  t.release();
  return st;
}

// ---------------------------------------------------------------------------------------------
// Copy sets: recorded BaseFab::copy calls (LBPatch::stream, the LB bounce-back fill), one launch
// ---------------------------------------------------------------------------------------------
static bool set_matches(const sgc_copyset_s* c, int which, const sgc_level_s* l) {
  return c->nbox[which] == l->nbox && c->ncomp[which] == l->ncomp && c->ng[which] == l->ng && c->esz == l->esz &&
         memcmp(c->boxes[which].data(), l->boxes.data(), sizeof(IBox) * (size_t)l->nbox) == 0;
}

int sgc_copyset_create(sgc_level_t dst, sgc_level_t src, const sgc_copy_op* ops, int nop, sgc_copyset_t* set) {
  SGC_REQUIRE_RT();
  if (!dst || !src || (!ops && nop > 0) || nop < 0 || !set) return fail(SGC_ERR_ARG, "sgc_copyset_create: bad argument");
  if (dst->esz != src->esz) return fail(SGC_ERR_ARG, "sgc_copyset_create: element types differ");
  std::vector<CopyItem> v;
  long long ncell = 0;
  for (int n = 0; n < nop; ++n) {
    const sgc_copy_op& o = ops[n];
    if (o.dst_box < 0 || o.dst_box >= dst->nbox || o.src_box < 0 || o.src_box >= src->nbox)
      return fail(SGC_ERR_ARG, "sgc_copyset_create: box index out of range");
    IBox dr, sr;
    for (int d = 0; d < 3; ++d) {
      dr.lo[d] = o.dst_lo[d]; dr.hi[d] = o.dst_hi[d];
      sr.lo[d] = o.src_lo[d]; sr.hi[d] = o.src_lo[d] + (o.dst_hi[d] - o.dst_lo[d]);
    }
    if (dr.empty()) return fail(SGC_ERR_ARG, "sgc_copyset_create: empty region");
    if (!fab_contains(dst->geom[o.dst_box], dr) || !fab_contains(src->geom[o.src_box], sr))
      return fail(SGC_ERR_ARG, "sgc_copyset_create: region outside fab");
    if (o.dst_comp < 0 || o.src_comp < 0 || o.ncomp < 1 || o.dst_comp + o.ncomp > dst->ncomp ||
        o.src_comp + o.ncomp > src->ncomp)
      return fail(SGC_ERR_ARG, "sgc_copyset_create: component range");
    make_items(v, dst->geom[o.dst_box], dr, o.dst_comp, src->geom[o.src_box], sr, o.src_comp, o.ncomp, o.comp_flags);
    ncell += (long long)dr.ext(0) * dr.ext(1) * dr.ext(2) * o.ncomp;
  }
  sgc_copyset_s* c = new sgc_copyset_s;
  const sgc_level_s* both[2] = {dst, src};
  for (int w = 0; w < 2; ++w) {
    c->nbox[w] = both[w]->nbox; c->ncomp[w] = both[w]->ncomp; c->ng[w] = both[w]->ng;
    c->boxes[w] = both[w]->boxes;
  }
  c->esz = dst->esz;
  c->ncell = ncell;
  const int st = c->table.upload(v, dst->esz);
  if (st) { delete c; return st; }
  *set = c;
  return 0;
}

int sgc_copyset_run(sgc_copyset_t c, sgc_level_t dst, sgc_level_t src) {
  SGC_REQUIRE_RT();
  if (!c || !dst || !src) return fail(SGC_ERR_ARG, "sgc_copyset_run: null argument");
  if (!set_matches(c, 0, dst) || !set_matches(c, 1, src))
    return fail(SGC_ERR_GEOMETRY, "sgc_copyset_run: levels differ from the ones the set was recorded on");
  level_written(dst);
  return launch_copy(c->table, dst->d_arena, src->d_arena, c->esz, rt().compute);
}

int sgc_copyset_destroy(sgc_copyset_t c) {
  if (!c) return 0;
  if (rt().ready) cudaStreamSynchronize(rt().compute);
  c->table.release();
  delete c;
  return 0;
}

long long sgc_copyset_num_cells(sgc_copyset_t c) { return c ? c->ncell : -1; }

static int linear_io(sgc_level_t l, int box, const int rlo[3], const int rhi[3], int c0, int c1, unsigned flags,
                     void* buffer, bool out) {
  SGC_REQUIRE_RT();
  if (!l || !rlo || !rhi || !buffer || box < 0 || box >= l->nbox) return fail(SGC_ERR_ARG, "linearOut/In: bad argument");
  if (c0 < 0 || c1 < c0 || c1 > l->ncomp) return fail(SGC_ERR_ARG, "linearOut/In: component range");
  IBox r;
  for (int d = 0; d < 3; ++d) { r.lo[d] = rlo[d]; r.hi[d] = rhi[d]; }
  const FabGeom& g = l->geom[box];
  if (r.empty() || !fab_contains(g, r)) return fail(SGC_ERR_ARG, "linearOut/In: region outside fab");
  // wire format: only the selected components are present, packed back to back (BaseFab.cpp:365-373)
  std::vector<CopyItem> v;
  long long pos = 0;
  for (int c = c0; c < c1; ++c) {
    if (!(c >= 32 || ((flags >> c) & 1u))) continue;
    const FabGeom bg = dense_geom(r, pos);
    if (out) make_items(v, bg, r, 0, g, r, c, 1, ~0u); else make_items(v, g, r, c, bg, r, 0, 1, ~0u);
    pos += r.size();
  }
  if (pos == 0) return 0;
  int st = ensure_scratch(l, (size_t)pos * l->esz);
  if (st) return st;
  CopyTable t;
  st = t.upload(v, l->esz);
  if (st) return st;
  if (out) {
    st = launch_copy(t, l->d_scratch, l->d_arena, l->esz, rt().compute);
    if (!st) {
      cudaError_t e = cudaMemcpyAsync(buffer, l->d_scratch, (size_t)pos * l->esz, cudaMemcpyDeviceToHost, rt().compute);
      if (e == cudaSuccess) e = cudaStreamSynchronize(rt().compute);
      if (e != cudaSuccess) st = fail(SGC_ERR_CUDA, cudaGetErrorString(e));
    }
  } else {
    cudaError_t e = cudaMemcpyAsync(l->d_scratch, buffer, (size_t)pos * l->esz, cudaMemcpyHostToDevice, rt().compute);
    if (e != cudaSuccess) st = fail(SGC_ERR_CUDA, cudaGetErrorString(e));
    level_written(l);
    if (!st) st = launch_copy(t, l->d_arena, l->d_scratch, l->esz, rt().compute);
    if (!st) cudaStreamSynchronize(rt().compute);
  }
  t.release();
  return st;
}

int sgc_fab_linear_out(sgc_level_t l, int box, const int rlo[3], const int rhi[3], int c0, int c1, unsigned flags,
                       void* buffer) {
  return linear_io(l, box, rlo, rhi, c0, c1, flags, buffer, true);
}
int sgc_fab_linear_in(sgc_level_t l, int box, const int rlo[3], const int rhi[3], int c0, int c1, unsigned flags,
                      const void* buffer) {
  return linear_io(l, box, rlo, rhi, c0, c1, flags, const_cast<void*>(buffer), false);
}

// ------------------------------------------------------------------------------- plans
static sgc_plan_s* new_plan_for(sgc_level_s* l, int c0, int nc) {
  sgc_plan_s* p = new sgc_plan_s;
  p->nbox = l->nbox; p->ncomp = l->ncomp; p->ng = l->ng; p->esz = l->esz; p->gidx0 = l->gidx0;
  p->boxes = l->boxes;
  p->c0 = c0; p->nc = nc;
  p->box_is_boundary.assign(l->nbox, 0);
  return p;
}

static int upload_local_items(sgc_plan_s* p, sgc_level_s* l) {
  std::vector<CopyItem> v, vx;
  v.reserve(p->items.size());
  // a second table whose x-face / x-edge / corner sources (one column of the neighbour: nghost = 1) are the
  // level's edge strips xe[box][kz][4][vny] instead of one 8-byte element per 8*vnx-byte row
  const bool strips = l->xe_elems > 0 && p->c0 == 0 && p->nc == 1;
  if (strips) vx.reserve(p->items.size());
  p->nremote = 0;
  for (const sgc_motion_item& it : p->items) {
    const int ld = it.recv_box - l->gidx0;
    if (ld < 0 || ld >= l->nbox) return fail(SGC_ERR_ARG, "plan: receiving box is not local to this level");
    if (it.recv_proc != it.send_proc) { ++p->nremote; p->box_is_boundary[ld] = 1; continue; }
    const int ls = it.send_box - l->gidx0;
    if (ls < 0 || ls >= l->nbox) return fail(SGC_ERR_ARG, "plan: local item refers to a box outside this level");
    IBox dr, sr;
    for (int d = 0; d < 3; ++d) { dr.lo[d] = it.recv_lo[d]; dr.hi[d] = it.recv_hi[d]; sr.lo[d] = it.src_lo[d]; sr.hi[d] = it.src_hi[d]; }
    if (dr.empty()) continue;
    if (!fab_contains(l->geom[ld], dr) || !fab_contains(l->geom[ls], sr))
      return fail(SGC_ERR_ARG, "plan: motion item region outside its fab (nghost of the level too small?)");
    const size_t n0 = v.size();
    make_items(v, l->geom[ld], dr, p->c0, l->geom[ls], sr, p->c0, p->nc, it.comp_flags);
    if (strips) {
      const FabGeom& gs = l->geom[ls];
      const int vhi = gs.vlo[0] + gs.vn[0] - 1;
      bool valid_src = sr.ext(0) == 1 && (sr.lo[0] == gs.vlo[0] || sr.lo[0] == vhi) && v.size() == n0 + 1;
      for (int d = 1; d < 3; ++d) valid_src = valid_src && sr.lo[d] >= gs.vlo[d] && sr.hi[d] < gs.vlo[d] + gs.vn[d];
      for (size_t k = n0; k < v.size(); ++k) {
        CopyItem c = v[k];
        if (valid_src) {
          const int strip = sr.lo[0] == gs.vlo[0] ? 0 : 3;
          c.src = l->xe_base + (((long long)ls * gs.vn[2] + (sr.lo[2] - gs.vlo[2])) * 4 + strip) * gs.vn[1] + (sr.lo[1] - gs.vlo[1]);
          c.ssx = 0; c.ssy = 1; c.ssz = 4 * gs.vn[1];
        }
        vx.push_back(c);
      }
    }
  }
  p->n_boundary_boxes = 0;
  for (char c : p->box_is_boundary) p->n_boundary_boxes += c;
  if (strips) {
    const int st = p->local_xe.upload(vx, l->esz);
    if (st) return st;
  }
  return p->local.upload(v, l->esz);
}

int sgc_plan_create_exchange(sgc_level_t l, const int dlo[3], const int dhi[3], const int maxbox[3], unsigned periodic,
                             unsigned trim, int comp_begin, int ncomp, int nproc, int rank, sgc_plan_t* plan) {
  return sgc_plan_create_exchange_ex(l, dlo, dhi, maxbox, 0, periodic, trim, comp_begin, ncomp, nproc, rank, nullptr, 0, plan);
}

int sgc_plan_create_exchange_ex(sgc_level_t l, const int dlo[3], const int dhi[3], const int maxbox[3], int nghost_exchange,
                                unsigned periodic, unsigned trim, int comp_begin, int ncomp, int nproc, int rank,
                                const unsigned* item_flags, int nitem, sgc_plan_t* plan) {
  SGC_REQUIRE_RT();
  if (!l || !dlo || !dhi || !maxbox || !plan) return fail(SGC_ERR_ARG, "sgc_plan_create_exchange: null argument");
  const int ngx = nghost_exchange > 0 ? nghost_exchange : l->ng;
  if (ngx > l->ng) return fail(SGC_ERR_ARG, "sgc_plan_create_exchange: exchange width exceeds the level's ghost width");
  if (comp_begin < 0 || ncomp < 1 || comp_begin + ncomp > l->ncomp)
    return fail(SGC_ERR_ARG, "sgc_plan_create_exchange: component range");
  Grid g(dlo, dhi, maxbox, nproc);
  if (g.status) return fail(SGC_ERR_ARG, "sgc_plan_create_exchange: boxes must tile the domain and divide among ranks");
  if (g.bpp != l->nbox || l->gidx0 != rank * g.bpp)
    return fail(SGC_ERR_GEOMETRY, "sgc_plan_create_exchange: level does not hold this rank's boxes of the layout");
  for (int b = 0; b < l->nbox; ++b) {
    const IBox gb = g.box(l->gidx0 + b);
    if (memcmp(&gb, &l->boxes[b], sizeof(IBox)) != 0)
      return fail(SGC_ERR_GEOMETRY, "sgc_plan_create_exchange: level boxes differ from the layout");
  }
  sgc_plan_s* p = new_plan_for(l, comp_begin, ncomp);
  p->nproc = nproc; p->rank = rank;
  p->ghost_only = true;          // motion items of an exchange write ghost cells only (Copier.H:521-659)
  p->has_grid = true;
  for (int d = 0; d < 3; ++d) { p->dlo[d] = dlo[d]; p->dhi[d] = dhi[d]; p->maxbox[d] = maxbox[d]; }
  p->periodic = periodic; p->trim = trim;
  p->ng_exchange = ngx;
  int st = build_exchange_items(g, ngx, periodic, trim, rank, p->items);
  if (!st && item_flags) {
    // Motion2Way::setCompRecvFlags (Copier.H:430-444) on some items, local or remote: same items, edited masks
    if (nitem != (int)p->items.size()) st = fail(SGC_ERR_ARG, "sgc_plan_create_exchange_ex: one flag word per motion item expected");
    else for (int i = 0; i < nitem; ++i) p->items[i].comp_flags = item_flags[i];
  }
  if (!st) st = upload_local_items(p, l);
  if (!st && p->nremote > 0) st = comm_plan_setup(p, l, g, periodic, trim);
  if (!st && ngx == 1 && !item_flags && l->ng == 1 && l->ncomp == 1 && comp_begin == 0 && ncomp == 1 && l->uniform && l->esz == 8) {
    // neighbour table of the fused / two-step sweeps and the step loops the plan admits (sgc_plan.cpp)
    std::vector<int> nbr;
    const int loops = classify_step_loops(p->items, l->nbox, l->gidx0, nbr);
    cudaError_t e = st ? cudaErrorUnknown : cudaMalloc(&p->d_push, sizeof(int) * nbr.size());
    if (e == cudaSuccess) e = cudaMemcpy(p->d_push, nbr.data(), sizeof(int) * nbr.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess && !st) st = fail(SGC_ERR_CUDA, std::string("push table: ") + cudaGetErrorString(e));
    p->fusable = (st == 0);
    p->t2_local = p->fusable && (loops & SGC_LOOP_TWO_STEP_LOCAL);
    p->t2_slab = p->fusable && (loops & SGC_LOOP_TWO_STEP_SLAB);
    p->t2_bnd = p->fusable && (loops & SGC_LOOP_TWO_STEP_BND);
    pillar_structure(l->boxes, nbr, p->t2_npil, p->t2_nbz);
  }
  if (st) { sgc_plan_destroy(p); return st; }
  *plan = p;
  return 0;
}

int sgc_plan_create_items(sgc_level_t l, const sgc_motion_item* items, int nitem, int comp_begin, int ncomp,
                          sgc_plan_t* plan) {
  SGC_REQUIRE_RT();
  if (!l || (!items && nitem > 0) || nitem < 0 || !plan) return fail(SGC_ERR_ARG, "sgc_plan_create_items: bad argument");
  if (comp_begin < 0 || ncomp < 1 || comp_begin + ncomp > l->ncomp)
    return fail(SGC_ERR_ARG, "sgc_plan_create_items: component range");
  sgc_plan_s* p = new_plan_for(l, comp_begin, ncomp);
  p->items.assign(items, items + nitem);
  for (const sgc_motion_item& it : p->items)
    if (it.recv_proc != it.send_proc) {
      sgc_plan_destroy(p);
      return fail(SGC_ERR_ARG, "sgc_plan_create_items: only local items are accepted");
    }
  const int st = upload_local_items(p, l);
  if (st) { sgc_plan_destroy(p); return st; }
  *plan = p;
  return 0;
}

int sgc_plan_create_bc_zero_gradient(sgc_level_t l, const int dlo[3], const int dhi[3], sgc_plan_t* plan) {
  SGC_REQUIRE_RT();
  if (!l || !dlo || !dhi || !plan) return fail(SGC_ERR_ARG, "sgc_plan_create_bc_zero_gradient: null argument");
  if (l->ng < 1) return fail(SGC_ERR_ARG, "sgc_plan_create_bc_zero_gradient: level has no ghost cells");
  std::vector<sgc_motion_item> items;
  for (int b = 0; b < l->nbox; ++b) {
    const IBox& v = l->boxes[b];
    for (int dir = 0; dir < 3; ++dir)
      for (int side = -1; side <= 1; side += 2) {
        const bool touches = side < 0 ? v.lo[dir] == dlo[dir] : v.hi[dir] == dhi[dir];
        if (!touches) continue;
        IBox src = v;                                   // boundary layer of valid cells
        if (side < 0) src.hi[dir] = src.lo[dir]; else src.lo[dir] = src.hi[dir];
        IBox dstb = src;
        dstb.lo[dir] += side; dstb.hi[dir] += side;   // the ghost face next to it
        sgc_motion_item it;
        memset(&it, 0, sizeof it);
        it.recv_box = it.send_box = l->gidx0 + b;
        for (int d = 0; d < 3; ++d) {
          it.recv_lo[d] = dstb.lo[d]; it.recv_hi[d] = dstb.hi[d];
          it.src_lo[d] = src.lo[d];   it.src_hi[d] = src.hi[d];
          it.send_lo[d] = src.lo[d];  it.send_hi[d] = src.hi[d];
        }
        it.dir[dir] = side;
        it.comp_flags = ~0u;
        items.push_back(it);
      }
  }
  const int st = sgc_plan_create_items(l, items.data(), (int)items.size(), 0, l->ncomp, plan);
  if (!st) (*plan)->ghost_only = true;
  return st;
}

static void rebox_release(sgc_plan_s* p);

int sgc_plan_destroy(sgc_plan_t p) {
  if (!p) return 0;
  if (rt().ready) { cudaStreamSynchronize(rt().compute); cudaStreamSynchronize(rt().comm); }
  p->local.release();
  p->local_xe.release();
  if (p->d_push) cudaFree(p->d_push);
  comm_p2p_release(p);
  comm_plan_release(p);
  rebox_release(p);
  if (p->t2_scratch) sgc_level_destroy(p->t2_scratch);
  delete p;
  return 0;
}

int sgc_plan_num_items(sgc_plan_t p) { return p ? (int)p->items.size() : -1; }
int sgc_plan_matches_level(sgc_plan_t p, sgc_level_t l) { return (p && l && plan_matches(p, l)) ? 1 : 0; }
int sgc_plan_num_remote_items(sgc_plan_t p) { return p ? p->nremote : -1; }
long long sgc_plan_num_cells(sgc_plan_t p) {
  if (!p) return -1;
  long long n = 0;
  for (const sgc_motion_item& it : p->items) {
    long long c = 1;
    for (int d = 0; d < 3; ++d) c *= (it.recv_hi[d] - it.recv_lo[d] + 1);
    n += c * p->nc;
  }
  return n;
}

int sgc_exchange_begin(sgc_level_t l, sgc_plan_t p) {
  SGC_REQUIRE_RT();
  if (!l || !p) return fail(SGC_ERR_ARG, "sgc_exchange: null handle");
  if (!plan_matches(p, l)) return fail(SGC_ERR_GEOMETRY, "sgc_exchange: plan was built for a different layout");
  if (p->in_flight) return fail(SGC_ERR_ARG, "sgc_exchange_begin: previous exchange not ended");
  if (p->nremote > 0) {
    const int st = comm_exchange_begin(l, p);
    if (st) return st;
  }
  int st;
  {
    ProfScope prof(1, rt().compute);
    // x-face sources: the level's edge strips (dense) when they mirror its valid cells, else the main blocks
    const bool strips = p->local_xe.nitem > 0 && l->xe_valid >= 1 && !getenv("SGC_NO_XE");
    if (!p->ghost_only) level_written(l);
    st = launch_copy(strips ? p->local_xe : p->local, l->d_arena, l->d_arena, l->esz, rt().compute);
  }
  if (st) return st;
  p->in_flight = true;
  return 0;
}

int sgc_exchange_end(sgc_level_t l, sgc_plan_t p) {
  SGC_REQUIRE_RT();
  if (!l || !p) return fail(SGC_ERR_ARG, "sgc_exchange: null handle");
  if (!p->in_flight) return fail(SGC_ERR_ARG, "sgc_exchange_end without sgc_exchange_begin");
  p->in_flight = false;
  if (p->nremote > 0) return comm_exchange_end(l, p);
  return 0;
}

int sgc_exchange(sgc_level_t l, sgc_plan_t p) {
  const int st = sgc_exchange_begin(l, p);
  if (st) return st;
  return sgc_exchange_end(l, p);
}

// ------------------------------------------------------------------------------- stencil apply
static int check_sweep(sgc_level_s* out, sgc_level_s* in, const char* who) {
  if (!out || !in) return fail(SGC_ERR_ARG, std::string(who) + ": null level");
  // Real is double, or float as in the reference's USE_SINGLE_PRECISION build (Parameters.H:17-23)
  if ((out->esz != 8 && out->esz != 4) || in->esz != out->esz)
    return fail(SGC_ERR_UNSUPPORTED, std::string(who) + ": levels must both hold double or both hold float");
  if (out->ncomp != in->ncomp) return fail(SGC_ERR_GEOMETRY, std::string(who) + ": levels have different component counts");
  if (out == in) return fail(SGC_ERR_ARG, std::string(who) + ": in-place sweep is not defined");
  // The swept region of box b is out's valid box; the two levels need not share a layout (the PR1
  // driver sweeps B on (2..n-1)^3 from A on (1..n)^3, laplacian.cpp:13-33) but must pair up box by box.
  if (out->nbox != in->nbox) return fail(SGC_ERR_GEOMETRY, std::string(who) + ": levels have different box counts");
  for (int b = 0; b < out->nbox; ++b)
    if (!fab_contains(in->geom[b], out->boxes[b].grown(1)))
      return fail(SGC_ERR_ARG, std::string(who) + ": input fab must contain grow(box,1)");
  return 0;
}

int sgc_heat_sweep_range(sgc_level_t unew, sgc_level_t u, double f, unsigned flags, int box_begin, int box_end) {
  SGC_REQUIRE_RT();
  int st = check_sweep(unew, u, "sgc_heat_sweep");
  if (st) return st;
  if (box_begin < 0 || box_end > unew->nbox || box_begin > box_end) return fail(SGC_ERR_ARG, "sgc_heat_sweep: box range");
  if (box_begin == box_end) return 0;
  const int fma = (flags & SGC_FMA) ? 1 : 0;
  ProfScope prof(0, rt().compute);
  st = launch_heat_stream(unew, u, f, fma, box_begin, box_end, rt().compute);
  if (st <= 0) return st;
  const int w0 = unew->work_begin[box_begin], w1 = unew->work_begin[box_end];
  level_written(unew);
  return launch_heat_generic(unew->d_work, w0, w1 - w0, unew->ncomp, unew->esz, u->d_geom, unew->d_geom, u->d_arena,
                             unew->d_arena, f, fma, unew->tile_z, rt().compute);
}

int sgc_heat_sweep(sgc_level_t unew, sgc_level_t u, double f, unsigned flags) {
  SGC_REQUIRE_RT();
  if (!unew) return fail(SGC_ERR_ARG, "sgc_heat_sweep: null level");
  return sgc_heat_sweep_range(unew, u, f, flags, 0, unew->nbox);
}

int sgc_laplacian_apply(sgc_level_t B, sgc_level_t A, double coef, int niter, int fuse_iters, unsigned flags) {
  SGC_REQUIRE_RT();
  int st = check_sweep(B, A, "sgc_laplacian_apply");
  if (st) return st;
  if (niter < 0) return fail(SGC_ERR_ARG, "sgc_laplacian_apply: niter");
  const int launches = fuse_iters ? 1 : niter;
  const int per = fuse_iters ? niter : 1;
  level_written(B);
  for (int i = 0; i < launches && !st; ++i)
    st = launch_lap_generic(B->d_work, 0, B->nwork, B->ncomp, B->esz, A->d_geom, B->d_geom, A->d_arena,
                            B->d_arena, coef, (flags & SGC_FMA) ? 1 : 0, per, B->tile_z, rt().compute);
  return st;
}

int sgc_wave_step(sgc_level_t unp1, sgc_level_t un, sgc_level_t unm1, double factor, unsigned flags) {
  SGC_REQUIRE_RT();
  int st = check_sweep(unp1, un, "sgc_wave_step");
  if (st) return st;
  if (!unm1 || !same_geometry(un, unm1)) return fail(SGC_ERR_GEOMETRY, "sgc_wave_step: un and unm1 must share geometry");
  if (unm1->esz != un->esz) return fail(SGC_ERR_UNSUPPORTED, "sgc_wave_step: element types differ");
  if (unm1 == unp1) return fail(SGC_ERR_ARG, "sgc_wave_step: unp1 must not alias unm1");
  ProfScope prof(0, rt().compute);
  level_written(unp1);
  const int wf = (int)(flags & (SGC_FMA | SGC_WAVE_DIRSUM));      // bit 0 fma, bit 1 dirsum form
  st = launch_wave_stream(unp1, un, unm1, factor, wf, rt().compute);
  if (st <= 0) return st;
  return launch_wave_generic(unp1->d_work, 0, unp1->nwork, unp1->ncomp, unp1->esz, un->d_geom, unp1->d_geom, un->d_arena,
                             unp1->d_arena, unm1->d_arena, factor, wf, unp1->tile_z, rt().compute);
}

int sgc_stencil7_apply(sgc_level_t out, sgc_level_t in, const double weights[7], unsigned flags) {
  SGC_REQUIRE_RT();
  int st = check_sweep(out, in, "sgc_stencil7_apply");
  if (st) return st;
  if (!weights) return fail(SGC_ERR_ARG, "sgc_stencil7_apply: null weights");
  ProfScope prof(0, rt().compute);
  level_written(out);
  return launch_weight_generic(out->d_work, 0, out->nwork, out->ncomp, out->esz, in->d_geom, out->d_geom, in->d_arena,
                               out->d_arena, weights, (flags & SGC_FMA) ? 1 : 0, out->tile_z, rt().compute);
}

int sgc_heat_sweep_ranges(sgc_level_t nxt, sgc_level_t cur, double f, unsigned flags, int a0, int a1, int b0, int b1) {
  SGC_REQUIRE_RT();
  int st = check_sweep(nxt, cur, "sgc_heat_sweep_ranges");
  if (st) return st;
  if (a0 < 0 || a0 > a1 || a1 > b0 || b0 > b1 || b1 > nxt->nbox) return fail(SGC_ERR_ARG, "sgc_heat_sweep_ranges: box ranges");
  {
    ProfScope prof(0, rt().compute);
    st = launch_heat_stream2(nxt, cur, f, (flags & SGC_FMA) ? 1 : 0, a0, a1, b0, b1, rt().compute, nullptr);
  }
  if (st <= 0) return st;
  st = sgc_heat_sweep_range(nxt, cur, f, flags, a0, a1);
  if (!st) st = sgc_heat_sweep_range(nxt, cur, f, flags, b0, b1);
  return st;
}

int sgc_wave_run(sgc_level_t levels[3], sgc_plan_t exchange_plan, sgc_plan_t bc_plan, double factor, unsigned flags,
                 int nstep, int latest[3]) {
  SGC_REQUIRE_RT();
  if (!levels || !levels[0] || !levels[1] || !levels[2] || nstep < 0) return fail(SGC_ERR_ARG, "sgc_wave_run: bad argument");
  int n = 0, np1 = 1, nm1 = 2;
  auto step = [&]() -> int {
    int st = 0;
    if (exchange_plan) st = sgc_exchange(levels[n], exchange_plan);
    if (!st && bc_plan) st = sgc_exchange(levels[n], bc_plan);
    if (!st) st = sgc_wave_step(levels[np1], levels[n], levels[nm1], factor, flags);
    const int t = nm1; nm1 = n; n = np1; np1 = t;      // unp1 -> un, un -> unm1 (WavePatch::advanceStepIndex)
    return st;
  };
  int s = 0;
  // three steps rotate the levels back to their roles: run three, then replay them as one graph
  if (nstep >= 12 && step_graph_ok(exchange_plan ? exchange_plan->nremote : 0, bc_plan ? bc_plan->nremote : 0)) {
    for (; s < 3; ++s) { const int st = step(); if (st) return st; }
    StepGraph g;
    int st = g.begin();
    for (int k = 0; k < 3 && !st; ++k) st = step();
    const int st2 = g.end();
    if (st || st2) return st ? st : st2;
    const int times = (nstep - s) / 3;
    st = g.replay(times);
    if (st) return st;
    s += 3 * times;
  }
  for (; s < nstep; ++s) { const int st = step(); if (st) return st; }
  if (latest) { latest[0] = n; latest[1] = np1; latest[2] = nm1; }
  return 0;
}

// streaming sweep of boxes [a0,a1) + [b0,b1) that also fills the ghosts of nxt's neighbours
static int fused_sweep(sgc_level_t nxt, sgc_level_t cur, double f, unsigned flags, int a0, int a1, int b0, int b1,
                       const int* d_push, int mask, const PushArgs* remote = nullptr) {
  ProfScope prof(0, rt().compute);
  PushArgs pa;
  memset(&pa, 0, sizeof pa);
  if (remote) pa = *remote;
  pa.nbr = d_push;
  pa.mask = mask;
  const int st = launch_heat_stream2(nxt, cur, f, (flags & SGC_FMA) ? 1 : 0, a0, a1, b0, b1, rt().compute, &pa);
  return st > 0 ? fail(SGC_ERR_UNSUPPORTED, "fused sweep: no streaming kernel for this geometry") : st;
}

// how long a kernel may wait for a neighbouring GPU's progress counter before it reports a time-out
static unsigned long long p2p_wait_budget_ns() {
  static const char* e = getenv("SGC_P2P_TIMEOUT_MS");
  const double ms = e ? atof(e) : 30000.0;
  return (unsigned long long)((ms > 1.0 ? ms : 1.0) * 1e6);
}

// `npass` two-step passes (sgc_stream2.cu) starting from level `cur` of the pair (a, b); a is the level whose
// physical-boundary ghost cells belong to the even steps.  p2p: z tiles come out of the neighbouring GPUs' memory.
static int run_t2_passes(sgc_level_t a, sgc_level_t b, sgc_plan_t plan, double f, unsigned flags, int npass, bool p2p,
                         sgc_level_t& cur, sgc_level_t& nxt, sgc_level_t scratch = nullptr) {
  sgc_plan_s::P2P& S = plan->p2p;
  int st = 0;
  auto other = [&](sgc_level_t x) { return x == a ? b : a; };
  if (cur->xe_valid != 2) st = launch_build_xe(cur, rt().compute);
  if (!st && p2p) st = comm_p2p_signal(plan, ++S.counter, rt().compute);     // our strips are readable
  for (int k = 0; k < npass && !st; ++k) {
    T2Args ra;
    memset(&ra, 0, sizeof ra);
    ra.npil = plan->t2_npil;
    ra.nbz = plan->t2_nbz;
    if (p2p) {
      const int li = (cur == S.lev[0]) ? 0 : (cur == S.lev[1] ? 1 : 2);       // 2: the plan's third level
      ra.rz = S.d_rz;
      for (int q = 0; q < S.npeer; ++q) {
        ra.peer_in[q] = (const double*)((char*)S.peer_base[q][li] + cur->lead * cur->esz);
        ra.peer_flag[q] = (const unsigned long long*)S.peer_flag_base[q];
      }
      ra.wait_value = S.counter;
      ra.err = S.d_err;
      ra.wait_budget_ns = p2p_wait_budget_ns();
    }
    if (plan->t2_bnd) {
      // faces without a neighbour: step t of the run reads level a's ghost cells, step t+1 level b's —
      // whichever level currently holds the valid cells (passes start on even steps)
      ra.ghost[0] = (const double*)a->d_arena;
      ra.ghost[1] = (const double*)b->d_arena;
      if (p2p) {
        const int ia = (a == S.lev[0]) ? 0 : 1;
        for (int q = 0; q < S.npeer; ++q) {
          ra.peer_ghost[q][0] = (const double*)((char*)S.peer_base[q][ia] + a->lead * a->esz);
          ra.peer_ghost[q][1] = (const double*)((char*)S.peer_base[q][ia ^ 1] + b->lead * b->esz);
        }
      }
    }
    // an odd number of passes ends a -> b -> scratch -> a instead of ping-ponging
    const bool via_scratch = scratch && k == npass - 2;
    sgc_level_t out = via_scratch ? scratch : nxt;
    {
      ProfScope prof(2, rt().compute);
      st = launch_heat_t2(out, cur, f, (flags & SGC_FMA) ? 1 : 0, plan->d_push, &ra, rt().compute);
    }
    if (st > 0) st = fail(SGC_ERR_UNSUPPORTED, "two-step sweep: geometry not covered");
    if (!st && p2p) st = comm_p2p_signal(plan, ++S.counter, rt().compute);
    sgc_level_t prev = cur;
    cur = out;
    nxt = via_scratch ? other(prev) : ((scratch && k == npass - 1) ? other(cur) : prev);
  }
  return st;           // a timed-out wait turns the later launches into no-ops; the caller checks once, at the end
}

// ---------------------------------------------------------------------------------------------
// Re-boxed step loop for levels of many small boxes (BASELINE config "32 768 boxes of 16^3").
// A 16^3 fab carries 18^3 cells, its two-step tile would carry 20^2 per plane: the halo overhead, not
// HBM, bounds such levels (one-step sweep: 273 Gcell-updates/s at 0.86 of the copy bandwidth).  Valid
// cells do not depend on how the domain is cut into boxes (every cell is updated with the same operations
// in the same order from the same neighbours), so inside sgc_heat_run the level is advanced on a
// library-owned pair of levels whose fabs are UNIONS of cx x cy x cz of the user's boxes (64^2 footprint,
// up to 64 planes): one batched copy gathers the valid cells (and, on non-periodic sides, the physical
// ghost faces of both levels, which no exchange ever writes), the two-step passes run on the unions,
// one batched copy scatters the result back, and the user's levels run the last two (three) steps
// themselves, which restores every ghost cell of both levels exactly as the step-by-step schedule leaves it.
// ---------------------------------------------------------------------------------------------
static void rebox_release(sgc_plan_s* p) {
  sgc_plan_s::Rebox& R = p->rebox;
  R.f2c.release(); R.f2c_ghost.release(); R.c2f.release();
  if (R.plan) sgc_plan_destroy(R.plan);
  for (int k = 0; k < 2; ++k) if (R.lev[k]) sgc_level_destroy(R.lev[k]);
  R = sgc_plan_s::Rebox();
}

static int rebox_setup(sgc_plan_s* p, sgc_level_s* a) {
  sgc_plan_s::Rebox& R = p->rebox;
  if (R.tried) return 0;
  R.tried = true;
  if (!p->has_grid || p->nproc != 1 || !a->uniform || a->ncomp != 1 || a->ng != 1 || a->esz != 8 || (p->trim & SGC_TRIM_FACE) ||
      getenv("SGC_NO_REBOX"))
    return 0;
  Grid g(p->dlo, p->dhi, p->maxbox, 1);
  if (g.status) return 0;
  // union factors: 64^2 (else 32^2) footprint, as many planes as divide the grid up to 64
  int c[3] = {1, 1, 1};
  bool found = false;
  for (int T = 64; T >= 32 && !found; T /= 2) {
    if (T % p->maxbox[0] || T % p->maxbox[1]) continue;
    c[0] = T / p->maxbox[0]; c[1] = T / p->maxbox[1];
    if (g.nb[0] % c[0] || g.nb[1] % c[1]) continue;
    c[2] = 1;
    for (int k = 64 / p->maxbox[2]; k >= 1; --k)
      if (k * p->maxbox[2] <= 64 && g.nb[2] % k == 0) { c[2] = k; break; }
    const int vz = c[2] * p->maxbox[2];
    found = (c[0] > 1 || c[1] > 1) && vz >= 4 && vz % 2 == 0;
  }
  if (!found) return 0;
  int cmb[3];
  for (int d = 0; d < 3; ++d) { cmb[d] = p->maxbox[d] * c[d]; R.factor[d] = c[d]; }
  Grid cg(p->dlo, p->dhi, cmb, 1);
  if (cg.status) return 0;
  std::vector<int> boxes6((size_t)cg.nbox * 6);
  for (int i = 0; i < cg.nbox; ++i) {
    const IBox bx = cg.box(i);
    for (int d = 0; d < 3; ++d) { boxes6[6 * i + d] = bx.lo[d]; boxes6[6 * i + 3 + d] = bx.hi[d]; }
  }
  int st = sgc_level_create(boxes6.data(), cg.nbox, 1, 1, 8, 0, &R.lev[0]);
  if (!st) st = sgc_level_create(boxes6.data(), cg.nbox, 1, 1, 8, 0, &R.lev[1]);
  if (!st) st = sgc_plan_create_exchange(R.lev[0], p->dlo, p->dhi, cmb, p->periodic, 0, 0, 1, 1, 0, &R.plan);
  if (st || !heat_t2_covers(R.lev[1], R.lev[0]) || !(R.plan->t2_local || R.plan->t2_bnd)) { rebox_release(p); R.tried = true; return st; }
  std::vector<CopyItem> in, gh, out;
  for (int i = 0; i < g.nbox; ++i) {
    const int ci = (i % g.nb[0]) / c[0] + cg.nb[0] * (((i / g.nb[0]) % g.nb[1]) / c[1] + cg.nb[1] * ((i / (g.nb[0] * g.nb[1])) / c[2]));
    const IBox& vb = a->boxes[i];
    const FabGeom& fg = a->geom[i];
    const FabGeom& ug = R.lev[0]->geom[ci];
    make_items(in, ug, vb, 0, fg, vb, 0, 1, ~0u);
    make_items(out, fg, vb, 0, ug, vb, 0, 1, ~0u);
    for (int d = 0; d < 3; ++d) {
      if (p->periodic & (1u << d)) continue;
      for (int side = 0; side < 2; ++side) {
        if (side == 0 ? vb.lo[d] != p->dlo[d] : vb.hi[d] != p->dhi[d]) continue;
        IBox face = vb;
        face.lo[d] = face.hi[d] = side == 0 ? vb.lo[d] - 1 : vb.hi[d] + 1;
        make_items(gh, ug, face, 0, fg, face, 0, 1, ~0u);
      }
    }
  }
  st = R.f2c.upload(in, 8);
  if (!st) st = R.c2f.upload(out, 8);
  if (!st) st = R.f2c_ghost.upload(gh, 8);
  if (st) { rebox_release(p); R.tried = true; return st; }
  R.ok = true;
  return 0;
}

int sgc_plan_rebox_factors(sgc_plan_t p, int factor[3]) {
  if (!p || !factor) return fail(SGC_ERR_ARG, "sgc_plan_rebox_factors: null argument");
  for (int d = 0; d < 3; ++d) factor[d] = p->rebox.ok ? p->rebox.factor[d] : 1;
  return p->rebox.tried ? 1 : 0;
}

int sgc_plan_pillars(sgc_plan_t p, int pillars[2]) {
  if (!p || !pillars) return fail(SGC_ERR_ARG, "sgc_plan_pillars: null argument");
  const bool have = p->t2_npil > 0 && p->t2_nbz > 0;
  pillars[0] = have ? p->t2_npil : p->nbox;
  pillars[1] = have ? p->t2_nbz : 1;
  return 0;
}

int sgc_heat_run(sgc_level_t a, sgc_level_t b, sgc_plan_t plan, double f, unsigned flags, int nstep, int* result_in_b) {
  SGC_REQUIRE_RT();
  if (!a || !b || !plan || nstep < 0) return fail(SGC_ERR_ARG, "sgc_heat_run: bad argument");
  if (!same_geometry(a, b)) return fail(SGC_ERR_GEOMETRY, "sgc_heat_run: the two levels must share geometry");
  if (!plan_matches(plan, a)) return fail(SGC_ERR_GEOMETRY, "sgc_heat_run: plan was built for a different layout");
  // Boxes with remote motion items ("boundary" boxes; for a z-slab partition the first and last box
  // layers) versus the rest.
  int b0 = 0, b1 = a->nbox;
  bool split = false;
  if (plan->nremote > 0 && plan->n_boundary_boxes < a->nbox) {
    while (b0 < b1 && plan->box_is_boundary[b0]) ++b0;
    while (b1 > b0 && plan->box_is_boundary[b1 - 1]) --b1;
    split = true;
    for (int i = b0; i < b1; ++i) if (plan->box_is_boundary[i]) split = false;   // boundary boxes must be the two ends
    if (const char* e = getenv("SGC_OVERLAP")) split = split && e[0] != '0';
  }
  sgc_level_t cur = a, nxt = b;
  int st = 0;
  int s0 = 0;

  // ---- re-boxed passes (levels of many small boxes, single rank): all but the last two / three steps run as
  // two-step passes on the union fabs; the plain schedule at the bottom finishes on the user's levels
  if (nstep >= 8 && !plan->rebox.tried) { st = rebox_setup(plan, a); if (st) return st; }
  const bool reboxed = nstep >= 8 && plan->rebox.ok;
  if (reboxed) {
    sgc_plan_s::Rebox& R = plan->rebox;
    const int sc = (nstep - 2) & ~1;
    {
      ProfScope prof(1, rt().compute);
      level_written(R.lev[0]);
      st = launch_copy(R.f2c, R.lev[0]->d_arena, a->d_arena, 8, rt().compute);
      if (!st) st = launch_copy(R.f2c_ghost, R.lev[0]->d_arena, a->d_arena, 8, rt().compute);
      if (!st) st = launch_copy(R.f2c_ghost, R.lev[1]->d_arena, b->d_arena, 8, rt().compute);
    }
    sgc_level_t ccur = R.lev[0], cnxt = R.lev[1];
    if (!st) st = run_t2_passes(R.lev[0], R.lev[1], R.plan, f, flags, sc / 2, false, ccur, cnxt);
    if (!st) {
      ProfScope prof(1, rt().compute);
      level_written(a);
      st = launch_copy(R.c2f, a->d_arena, ccur->d_arena, 8, rt().compute);
    }
    if (st) return st;
    s0 = sc;                     // even: the step-by-step schedule has this state in level a, too
  }

  // ---- fused schedule ("pull y/z, push x", sgc_stream.cu): no exchange launch inside the loop.
  // One full exchange first (x strips and the ghosts facing physical boundaries / other GPUs);
  // then every step but the last two sweeps fused — with remote items the boundary box layers
  // first, their halo packed and sent on the comm stream while the interior boxes sweep; the last
  // two steps run unfused, which leaves BOTH levels bit-identical, ghost cells included, to
  // nstep x { exchange ; sweep } (the fused steps do not maintain edge / corner / y / z ghosts).
  // Boxes at least 32 cells wide: x faces pushed by the consumers, y/z pulled by the producer.
  // Narrower boxes: everything pulled — z planes by the producer, x strips and y rows by the gather
  // warp (sgc_stream.cu) — because on 16^3 boxes every 8th cell lies on an x face.
  const bool fused = !reboxed && plan->fusable && nstep >= 3 && heat_stream_covers(b, a) && !getenv("SGC_NO_FUSE");
  const int fmask = 7;                                 // bit 0 push x, bit 1 pull z, bit 2 pull y
  const char* cs = getenv("SGC_COMM_SMS");
  const int comm_sms = cs ? atoi(cs) : 8;              // SMs left to pack / NCCL / unpack during the interior sweep
  if (fused) {
    // Multi-GPU, z-slab partitions: the z ghost planes facing another GPU are pulled by the sweep's
    // producer straight out of that GPU's memory (cudaIpc-mapped arenas over NVLink), gated by a
    // progress counter the neighbour bumps after each of its sweeps — no pack / NCCL / unpack and
    // ONE launch per step.  SGC_P2P=0 keeps the NCCL schedule below.
    // the library-owned third level of the two-step passes (see below): on a multi-rank run it has to exist before the
    // peer-memory session, which maps it on the neighbours
    auto third_level = [&]() -> sgc_level_t {
      if (!plan->t2_scratch && !getenv("SGC_NO_T2_SCRATCH")) {
        std::vector<int> boxes6((size_t)a->nbox * 6);
        for (int i = 0; i < a->nbox; ++i)
          for (int d = 0; d < 3; ++d) { boxes6[6 * i + d] = a->boxes[i].lo[d]; boxes6[6 * i + 3 + d] = a->boxes[i].hi[d]; }
        if (sgc_level_create(boxes6.data(), a->nbox, 1, 1, 8, a->gidx0, &plan->t2_scratch)) plan->t2_scratch = nullptr;
      }
      return plan->t2_scratch;
    };
    const char* pe = getenv("SGC_P2P");
    const bool want_p2p = plan->nremote > 0 && !(pe && pe[0] == '0');
    sgc_level_t third = (want_p2p && heat_t2_covers(b, a)) ? third_level() : nullptr;
    const bool p2p = want_p2p && comm_p2p_setup(plan, a, b, third) == 0 && plan->p2p.ready;
    // every rank has reached this run before any kernel that waits on a neighbour's counter starts
    if (p2p) { st = comm_stream_barrier(rt().compute); if (st) return st; }

    // ---- two steps per pass (sgc_stream2.cu): level t is read once and level t+2 written once, the
    // intermediate level stays on the SM.  It reads VALID cells only (the neighbours' rows, planes and
    // edge strips), so no ghost cell is touched until the phase ends; an even number of passes keeps
    // the result in the level the step-by-step schedule would leave it in.
    int npass = (nstep - 2) / 2;
    const bool t2_remote_ok = plan->nremote == 0 || p2p;          // remote z faces need the peer-memory session
    bool t2 = heat_t2_covers(b, a) && (plan->t2_local || ((plan->t2_slab || plan->t2_bnd) && t2_remote_ok));
    // an odd number of passes needs the library-owned third level (with a peer-memory session: the one the
    // neighbours have mapped); otherwise one pass fewer
    sgc_level_t scratch = nullptr;
    if (npass & 1) {
      if (t2 && npass >= 3) scratch = plan->nremote == 0 ? third_level() : (p2p ? plan->p2p.lev[2] : nullptr);
      if (!scratch) npass &= ~1;
    }
    if (plan->nremote > 0) {
      // ranks may classify differently (the end ranks of a non-periodic slab partition see physical faces,
      // the others do not); the passes signal each other, so all ranks run them or none does
      if (plan->t2_agreed < 0) {
        int all = 0;
        st = comm_all_min(t2 ? 1 : 0, &all);
        if (st) return st;
        plan->t2_agreed = all;
      }
      t2 = t2 && plan->t2_agreed == 1;
    }
    if (npass >= 2 && t2) {
      st = run_t2_passes(a, b, plan, f, flags, npass, p2p, cur, nxt, scratch);
      if (st) return st;
      s0 += 2 * npass;
    }
    // Peer-memory session: the rest of the run needs no message either — every exchange reads its remote regions out
    // of the neighbours' arenas (comm_p2p_pull), gated by the same progress counters as the sweeps.
    // Opt-in (SGC_PULL=1): measured at N = 2 on the bench level it is 0.2 ms per 100 steps SLOWER than the NCCL
    // schedule of the last steps, which hides its 8 MB message behind the interior boxes (26.45 against 26.22 ms).
    const char* pl = getenv("SGC_PULL");
    const bool pull = p2p && pl && pl[0] == '1' && comm_p2p_can_pull(plan, a);
    auto exchange_pull = [&](sgc_level_t lev) -> int {
      ProfScope prof(1, rt().compute);
      const bool strips = plan->local_xe.nitem > 0 && lev->xe_valid >= 1 && !getenv("SGC_NO_XE");
      int e = launch_copy(strips ? plan->local_xe : plan->local, lev->d_arena, lev->d_arena, lev->esz, rt().compute);
      if (!e) e = comm_p2p_pull(plan, lev, p2p_wait_budget_ns(), rt().compute);
      return e;
    };
    // otherwise the exchange below also orders the phases on a multi-GPU run: its receives complete only after
    // the neighbours have left their last pass, so their reads of our levels are over
    // (with no fused step left the plain steps below start with this very exchange; on a multi-GPU run their
    // receives order the phases the same way: a boundary box is swept only after the neighbours' messages, which
    // they send after their last pass, have arrived, and interior boxes are never read by a neighbour)
    if (s0 < nstep - 2) st = pull ? exchange_pull(cur) : sgc_exchange(cur, plan);
    if (p2p && !st) {
      sgc_plan_s::P2P& S = plan->p2p;
      for (; s0 < nstep - 2 && !st; ++s0) {
        PushArgs ra;
        memset(&ra, 0, sizeof ra);
        const int li = (cur == S.lev[0]) ? 0 : 1;               // which of the session's levels is the input
        ra.rz = S.d_rz;
        for (int k = 0; k < S.npeer; ++k) {
          ra.peer_in[k] = (const double*)((char*)S.peer_base[k][li] + cur->lead * cur->esz);
          ra.peer_flag[k] = (const unsigned long long*)S.peer_flag_base[k];
        }
        ra.wait_value = S.counter;                              // peers have finished their previous sweep
        ra.err = S.d_err;
        ra.wait_budget_ns = p2p_wait_budget_ns();
        st = fused_sweep(nxt, cur, f, flags, 0, cur->nbox, cur->nbox, cur->nbox, plan->d_push, fmask, &ra);
        if (!st) st = comm_p2p_signal(plan, ++S.counter, rt().compute);
        std::swap(cur, nxt);
      }
      if (pull) {
        // the last two steps { exchange ; sweep } as they are, the exchange's remote half pulled: restores every
        // ghost cell of both levels exactly as the step-by-step schedule leaves it
        for (; s0 < nstep && !st; ++s0) {
          st = exchange_pull(cur);
          if (!st) st = sgc_heat_sweep(nxt, cur, f, flags);
          if (!st) st = comm_p2p_signal(plan, ++S.counter, rt().compute);
          std::swap(cur, nxt);
        }
      }
      if (!st) st = comm_p2p_check(plan);
      if (st) return st;
    }
    for (; s0 < nstep - 2 && !st; ++s0) {
      if (plan->nremote > 0) {
        if (split) st = fused_sweep(nxt, cur, f, flags, 0, b0, b1, cur->nbox, plan->d_push, fmask);
        else st = fused_sweep(nxt, cur, f, flags, 0, cur->nbox, cur->nbox, cur->nbox, plan->d_push, fmask);
        if (!st) st = comm_exchange_begin(nxt, plan);            // pack nxt's boundary cells, send/recv
        // the interior sweep is persistent (one CTA per SM, 209 KB of shared memory each): leave a few
        // SMs free, otherwise NCCL's CTAs cannot be placed until it ends and the halo is not overlapped
        rt().sweep_grid_cap = rt().sm_count - comm_sms;
        if (!st && split) st = fused_sweep(nxt, cur, f, flags, b0, b1, b1, b1, plan->d_push, fmask);
        rt().sweep_grid_cap = 0;
        if (!st) st = comm_exchange_end(nxt, plan);              // unpack into nxt's remote-facing ghosts
      } else {
        st = fused_sweep(nxt, cur, f, flags, 0, cur->nbox, cur->nbox, cur->nbox, plan->d_push, fmask);
      }
      std::swap(cur, nxt);
    }
    if (st) return st;
  }

  // ---- plain schedule.  "split": halo in flight while the boxes without remote items sweep, then
  // the boundary layers (two box ranges, one launch); otherwise exchange, then one sweep.
  // (after a fused phase only the last two or three steps are left: there one whole-level sweep behind a complete
  // exchange beats the split — 25.55 against 25.85 ms per 100 steps at N = 2 — so the split is for runs that are
  // plain from the first step)
  const bool split_here = split && !fused;
  for (int s = s0; s < nstep; ++s) {
    if (split_here) {
      st = sgc_exchange_begin(cur, plan);
      if (!st) st = sgc_heat_sweep_range(nxt, cur, f, flags, b0, b1);
      if (!st) st = sgc_exchange_end(cur, plan);
      if (!st) st = sgc_heat_sweep_ranges(nxt, cur, f, flags, 0, b0, b1, cur->nbox);
    } else {
      st = sgc_exchange(cur, plan);
      if (!st) st = sgc_heat_sweep(nxt, cur, f, flags);
    }
    if (st) return st;
    std::swap(cur, nxt);
  }
  if (result_in_b) *result_in_b = (cur == b) ? 1 : 0;
  return 0;
}

}  // extern "C"
