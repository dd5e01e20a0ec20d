/* sgc_b200.h — C ABI of the B200-native stencil-sweep + ghost-exchange path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.
 * The reference has no FFI of its own; what it has is (1) the C++ class API of
 * Box / BaseFab / DisjointBoxLayout / LevelData / Copier and (2) a host<->CUDA seam of free
 * functions taking PODs and opaque `AccelPointer` (void*) handles
 * (application/wave/WavePatch_Cuda.H:50-82).  Every entry point below names the reference
 * interface it replaces; citations are relative to /root/reference/Structured_gridcalc/.
 * The C++ facade in include/BoxFramework/ (same class names as the reference) forwards here;
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success and a negative sgc_status otherwise;
 *     sgc_last_error() gives the message (reference convention: no exceptions, CU_SAFE_CALL
 *     prints cudaGetErrorString — lib/src/BoxFramework/CudaSupport.H:21-29).
 *   - boxes cross the boundary as int[6] = {lo0,lo1,lo2,hi0,hi1,hi2}, closed index ranges
 *     (Box.H:56-228; the reference exposes the same int[6] view on device, Box.H:598-616).
 *   - fab memory order is the reference's: data[c*S + (i-lo0) + (j-lo1)*n0 + (k-lo2)*n0*n1]
 *     on the box grown by nghost (BaseFab.H:277-313, LevelData.H:266-271).  Host buffers passed
 *     in are borrowed for the call only.
 *   - all device work is enqueued on the library's compute stream; calls are asynchronous
 *     unless documented otherwise (downloads and sgc_sync block).  Not thread-safe per handle
 *     (reference: single host thread, everything on stream 0 — wave.cpp:261).
 *   - there is NO CPU fallback: every data-path entry point fails with SGC_ERR_CUDA when no
 *     CUDA device is usable.
 */
#ifndef SGC_B200_H
#define SGC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGC_ABI_VERSION 1

typedef enum {
  SGC_OK = 0,
  SGC_ERR_ARG = -1,        /* bad argument (the reference would CH_assert)            */
  SGC_ERR_CUDA = -2,       /* CUDA runtime error / no device                           */
  SGC_ERR_GEOMETRY = -3,   /* levels/plan built on different layouts (tag mismatch,    */
                           /* LevelData.H:285, LayoutIterator.H:802)                   */
  SGC_ERR_COMM = -4,       /* NCCL / multi-rank setup error                            */
  SGC_ERR_UNSUPPORTED = -5
} sgc_status;

/* TrimCenter/Face/Edge/Corner and PeriodicX/Y/Z — LayoutIterator.H:27-41 */
enum { SGC_TRIM_CENTER = 1, SGC_TRIM_FACE = 2, SGC_TRIM_EDGE = 4, SGC_TRIM_CORNER = 8 };
enum { SGC_PERIODIC_X = 1, SGC_PERIODIC_Y = 2, SGC_PERIODIC_Z = 4 };

/* Arithmetic flags for the stencil entry points.
 * SGC_FMA: evaluate `u + f*S` (and `r + f*T`) as one fused multiply-add, which is what the
 * reference's g++ build emits (-ffp-contract=fast); without it the product and the
 * sum round separately (= -ffp-contract=off).  Either way the association order is the
 * reference loop's: ((u[+e]-2u)+u[-e]) per direction, x then y then z.                    */
enum { SGC_FMA = 1,
       /* sgc_wave_step only: evaluate the update in the form of the shipped wave app's pencil path,
        * unp1 = (2 un - unm1) + factor*(Tz + (Ty + Tx))  (MD_DIRSUM, WavePatch.cpp:194-206), instead
        * of the scalar form with three separate accumulations (WavePatch.cpp:226-244).           */
       SGC_WAVE_DIRSUM = 2 };

typedef struct sgc_level_s* sgc_level_t;   /* device-resident LevelData<BaseFab<T>>      */
typedef struct sgc_plan_s*  sgc_plan_t;    /* device-resident Copier (exchange plan)      */
typedef struct sgc_event_s* sgc_event_t;   /* cudaEvent on the compute stream             */
typedef struct sgc_stream_s* sgc_stream_t; /* a transfer stream (cudaStream_t + its staging area) */
typedef struct sgc_copyset_s* sgc_copyset_t; /* a recorded list of BaseFab::copy calls, replayed in one launch */

/* One BaseFab::copy(dstBox, dstComp, src, srcBox, srcComp, numComp, flags) call (BaseFab.cpp:296-329)
 * between fab dst_box of one level and fab src_box of another (or the same) level; boxes are LOCAL
 * indices, the source region is [src_lo, src_lo + (dst_hi - dst_lo)].                         */
typedef struct {
  int dst_box, dst_lo[3], dst_hi[3], dst_comp;
  int src_box, src_lo[3], src_comp;
  int ncomp;
  unsigned comp_flags;
} sgc_copy_op;

/* One Motion2Way (Copier.H:37-160) as a POD. */
typedef struct {
  int recv_box;        /* global index of the local (receiving) box   m_bidxLocal          */
  int send_box;        /* global index of the remote (sending) box    m_bidxRemote         */
  int recv_lo[3], recv_hi[3];   /* regionRecv: ghost region of recv_box to fill            */
  int src_lo[3],  src_hi[3];    /* regionSendRemote: region of send_box to read            */
  int dir[3];          /* sendDir (Copier.H:117)                                           */
  int recv_proc, send_proc;     /* owners (rank == GPU)                                    */
  int tag_send, tag_recv;       /* uniqueTag = 27*globalIdx + dir code (Copier.H:388-395)  */
  unsigned comp_flags; /* m_compRecvFlags, default all ones                                */
  int send_lo[3], send_hi[3];   /* m_regionSend = local ∩ grow(remote,ng) (Copier.H:581)   */
} sgc_motion_item;

/* ------------------------------------------------------------------ runtime ------------- */

/* Bind the calling process to CUDA device `device` and create the library's streams.
 * Replaces: implicit device 0 / cudaFuncSetCacheConfig in WavePatch_Cuda::construct
 * (WavePatch_Cuda.cu:224-225).  Idempotent for the same device.                            */
int sgc_init(int device);
int sgc_shutdown(void);
int sgc_abi_version(void);
const char* sgc_last_error(void);
/* name (up to cap bytes), SM count, total HBM bytes of the bound device */
int sgc_device_info(char* name, int cap, int* sm_count, size_t* hbm_bytes);
/* Block until everything enqueued so far has finished (cudaDeviceSynchronize, wave.cpp:261). */
int sgc_sync(void);
/* Number of kernels this library has launched since sgc_init (for bench.py gpu_launches). */
long long sgc_launch_count(void);

/* Per-kernel device timing for the roofline report: while enabled, every sweep launch (class 0)
 * every batched copy launch of an exchange (class 1) and every two-step sweep launch (class 2,
 * sgc_heat_run's temporal-blocking passes) is bracketed by a cudaEvent pair on its
 * own stream.  sgc_profile_read synchronises and returns the launch count and the summed
 * duration per class, then clears the log.  At most 8192 launches are logged per class.      */
int sgc_profile_enable(int on);
int sgc_profile_read(int klass, int* launches, double* total_ms);

/* cudaEvent pairs around a device-resident step loop (wave.cpp:216-274,
 * WavePatch_Cuda.cu:363-377).  Recorded on the compute stream.                             */
int sgc_event_create(sgc_event_t* ev);
int sgc_event_destroy(sgc_event_t ev);
int sgc_event_record(sgc_event_t ev);
int sgc_event_elapsed_ms(sgc_event_t start, sgc_event_t stop, float* ms); /* syncs on stop */

/* Pinned host memory for H2D/D2H staging (BaseFab::allocate under USE_GPU uses
 * cudaMallocHost, BaseFab.cpp:515-519). */
int sgc_host_alloc(size_t bytes, void** ptr);
int sgc_host_free(void* ptr);

/* ------------------------------------------------------------------ layout (host only) --- */

/* DisjointBoxLayout::define (DisjointBoxLayout.cpp:93-168): uniform decomposition, boxes in
 * x-fastest order, rank k owns [k*n/nproc,(k+1)*n/nproc).  Writes up to cap boxes; returns the
 * number of boxes (>=0) or a negative status.  Needs no GPU.                               */
int sgc_layout_define(const int dlo[3], const int dhi[3], const int maxbox[3], int nproc,
                      int* boxes6, int* owner, int cap, int numbox[3]);

/* Copier::defineExchangeDBL (Copier.H:521-659) as seen by rank `rank` of `nproc`: motion
 * items in the reference's order (per local box: interior neighbours in x-fastest [-1,1]^3
 * order, then periodic images).  Writes up to cap items; returns the item count.  Needs no GPU. */
/* Which device-resident step loops sgc_heat_run may use for rank `rank`'s part of this layout and exchange
 * (host only; the level's box shape decides on top of it): a mask of               */
#define SGC_LOOP_FUSED 1            /* ghost fill fused into the one-step sweep                             */
#define SGC_LOOP_TWO_STEP_LOCAL 2   /* two steps per pass; every neighbour the kernel reads is on this rank  */
#define SGC_LOOP_TWO_STEP_SLAB 4    /* ... z neighbours may be on another rank (read over NVLink)            */
#define SGC_LOOP_TWO_STEP_BND 8     /* ... some faces have no motion item (physical boundary / trimmed)      */
int sgc_step_loop_class(const int dlo[3], const int dhi[3], const int maxbox[3], int nghost,
                        unsigned periodic, unsigned trim, int nproc, int rank);
/* The PILLARS the two-step sweep marches down for the same layout, exchange and rank (host only): pillars[0]
 * pillars of pillars[1] boxes stacked in z, each box the LOCAL z neighbour of the one below it under this
 * exchange (DisjointBoxLayout.cpp:104-139 orders boxes x-fastest, so box p + k * pillars[0] sits on top of
 * box p + (k-1) * pillars[0]) and with the same in-plane neighbours present; pillars[1] = 1 (every box its own
 * pillar) when the rank's boxes have no such structure.  Returns the step-loop mask of sgc_step_loop_class.  */
int sgc_step_loop_pillars(const int dlo[3], const int dhi[3], const int maxbox[3], int nghost,
                          unsigned periodic, unsigned trim, int nproc, int rank, int pillars[2]);

int sgc_copier_define(const int dlo[3], const int dhi[3], const int maxbox[3], int nghost,
                      unsigned periodic, unsigned trim, int nproc, int rank,
                      sgc_motion_item* items, int cap);

/* ------------------------------------------------------------------ level data ---------- */

/* LevelData<BaseFab<T>>::define (LevelData.H:256-272) for the nbox boxes local to this rank:
 * one device arena holding every fab on grow(box,nghost), reference memory order, fabs
 * contiguous in box order.  elem_bytes = sizeof(T): 8 (Real), 4 (int/unsigned/float) or
 * 1 (bool/char) — BaseFab.cpp:560-564; the stencil entry points need 8.
 * global_index0 = DisjointBoxLayout::localIdxBegin() (global index of local box 0).        */
int sgc_level_create(const int* boxes6, int nbox, int ncomp, int nghost, int elem_bytes,
                     int global_index0, sgc_level_t* level);
int sgc_level_destroy(sgc_level_t level);
/* total elements (all fabs, all comps) and element offset of fab `box` in host-side concatenation */
long long sgc_level_elems(sgc_level_t level);
long long sgc_level_fab_offset(sgc_level_t level, int box);
/* device address of fab `box` (BaseFab::m_dataSymbol.device, BaseFab.H:213-216) */
void* sgc_level_device_ptr(sgc_level_t level, int box);
/* Introspection (tests, bench): state of the level's edge strips — dense copies of the first / last valid
 * x columns of every fab plane that the streaming sweeps keep current so that the next
 * LevelData::exchange (LevelData.H:455-566) reads its x faces densely instead of one element per row:
 * -1 the level has none, 0 stale, 1 the outer columns are current, 2 all four.                      */
int sgc_level_edge_strips(sgc_level_t level);

/* LevelData::copyToDevice / copyToHost (LevelData.H:711-765) and BaseFab::copyToDevice/Host
 * (BaseFab.cpp:427-483).  box = -1: the whole level (host buffer = fabs concatenated in box
 * order); otherwise one fab.  Upload is asynchronous when `host` is pinned; download blocks.  */
int sgc_level_upload(sgc_level_t level, int box, const void* host);
int sgc_level_download(sgc_level_t level, int box, void* host);

/* LevelData::copyToDeviceAsync / copyToHostAsync(cudaStream_t) (LevelData.H:734-765,
 * BaseFab.H:175-181) and the overlap of plot-data D2H with compute in wave.cpp:230-275.
 * A transfer stream owns its own reference-order staging area, so whole-level transfers on it
 * run beside the compute stream (kernels of all other entry points) and beside each other:
 * batch i+1 uploads and batch i-1 downloads while batch i sweeps.  Both calls return after
 * enqueueing; `host` must be pinned (sgc_host_alloc) and stay valid until the stream passed it.
 * A download first gathers the level into the stream's staging area and then copies that to the
 * host: `gathered` (may be NULL) is recorded between the two, so a compute stream that waits for it
 * may overwrite the level while the D2H copy is still in flight.
 * Ordering between streams is explicit, through events: sgc_event_record_stream marks a point
 * of a stream, sgc_stream_wait_event makes work submitted LATER to `stream` start after that
 * point.  A NULL stream argument names the library's compute stream.                         */
int sgc_stream_create(sgc_stream_t* stream);
int sgc_stream_destroy(sgc_stream_t stream);
int sgc_stream_sync(sgc_stream_t stream);
int sgc_event_record_stream(sgc_event_t ev, sgc_stream_t stream);
int sgc_stream_wait_event(sgc_stream_t stream, sgc_event_t ev);
int sgc_level_upload_async(sgc_level_t level, const void* host, sgc_stream_t stream);
int sgc_level_download_async(sgc_level_t level, void* host, sgc_stream_t stream, sgc_event_t gathered);

/* Plot output from device data: what LevelData::writeCGNSSolData hands to CGNS
 * (LevelData.cpp:150-193: per box, per component, the core grid only — ghosts stripped through
 * memrmin/memrmax) gathered on the device into one dense buffer [box][comp][k][j][i] and copied
 * to the host; the async form runs on a transfer stream beside the compute stream, the overlap of
 * wave.cpp:230-275.  sgc_level_valid_elems = elements of that buffer.                          */
long long sgc_level_valid_elems(sgc_level_t level);
int sgc_level_download_valid(sgc_level_t level, void* host);
int sgc_level_download_valid_async(sgc_level_t level, void* host, sgc_stream_t stream, sgc_event_t gathered);
/* The reverse: the valid cells of every fab from a dense buffer [box][comp][k][j][i]; ghost cells keep their
 * contents (the next exchange / BC fill writes them).  Moves (n/(n+2g))^3 of the bytes of the whole-fab
 * copyToDevice (BaseFab.cpp:427-483, LevelData.H:711-765).                                              */
int sgc_level_upload_valid(sgc_level_t level, const void* host);
int sgc_level_upload_valid_async(sgc_level_t level, const void* host, sgc_stream_t stream);
/* LevelData::setVal / BaseFab::setVal (LevelData.H:352-383, BaseFab.cpp:219-246);
 * comp = -1: all components.  `value` points at one element of elem_bytes bytes.           */
int sgc_level_setval(sgc_level_t level, int box, int comp, const void* value);

/* Whole-level device copy dst <- src (all fabs, all components, ghosts included); both levels
 * must share geometry.  The device form of `for (dit) dst[dit].copy(src[dit].box(), src[dit])`
 * (BaseFab::copy(box,src), BaseFab.cpp:254-271), as WavePatch::initialData does for unm1
 * (WavePatch.cpp:129).  One cudaMemcpyAsync device-to-device on the compute stream.            */
int sgc_level_copy(sgc_level_t dst, sgc_level_t src);

/* BaseFab::copy(dstBox,dstComp,src,srcBox,srcComp,numComp,compFlags) (BaseFab.cpp:296-329)
 * between two device fabs; BaseFab::copy(box,src) (:254-271) is the same call with equal
 * regions, comps 0..ncomp.                                                                 */
int sgc_fab_copy(sgc_level_t dst, int dst_box, const int dlo[3], const int dhi[3], int dst_comp,
                 sgc_level_t src, int src_box, const int slo[3], const int shi[3], int src_comp,
                 int ncomp, unsigned comp_flags);
/* BaseFab::linearOut / linearIn (BaseFab.cpp:350-419): pack / unpack a region, wire order
 * comp -> k -> j -> i, no header.  `buffer` is HOST memory (packed on the device, then copied). */
int sgc_fab_linear_out(sgc_level_t level, int box, const int rlo[3], const int rhi[3],
                       int comp_begin, int comp_end, unsigned comp_flags, void* buffer);
int sgc_fab_linear_in(sgc_level_t level, int box, const int rlo[3], const int rhi[3],
                      int comp_begin, int comp_end, unsigned comp_flags, const void* buffer);

/* ------------------------------------------------------------------ exchange plan -------- */

/* Copier::defineExchangeLD / defineExchangeDBL (Copier.H:481-494,521-659) for the layout
 * `level` was created on: builds the motion items on the host (sgc_copier_define), uploads the
 * device item table once, and — when nproc > 1 — the per-peer pack/unpack tables.
 * comp_begin/ncomp: component range (m_startComp, numComp).                                */
int sgc_plan_create_exchange(sgc_level_t level, const int dlo[3], const int dhi[3],
                             const int maxbox[3], unsigned periodic, unsigned trim,
                             int comp_begin, int ncomp, int nproc, int rank, sgc_plan_t* plan);
/* The general form.  nghost_exchange: ghost width the items fill (Copier::defineExchangeDBL's a_numGhost,
 * Copier.H:521-528; 0 = the level's own width, which must not be smaller).  item_flags (or NULL): one
 * component mask per motion item, in sgc_copier_define's order, replacing the default ~0u — the state of a
 * Copier after Motion2Way::setCompRecvFlags (Copier.H:430-444) on some of its items, local OR remote (the
 * masks of remote items are applied when their message is unpacked).                                  */
int sgc_plan_create_exchange_ex(sgc_level_t level, const int dlo[3], const int dhi[3], const int maxbox[3],
                                int nghost_exchange, unsigned periodic, unsigned trim, int comp_begin, int ncomp,
                                int nproc, int rank, const unsigned* item_flags, int nitem, sgc_plan_t* plan);
/* 1 when `plan` was built for a level of this geometry (boxes, components, ghost width, element size): a
 * Copier may serve several LevelData on its layout (LevelData.H:455: any level whose tag matches).      */
int sgc_plan_matches_level(sgc_plan_t plan, sgc_level_t level);
/* Same from caller-supplied items (all local), e.g. after Motion2Way::setCompRecvFlags
 * (Copier.H:430-444) or for physical-BC copies.                                            */
int sgc_plan_create_items(sgc_level_t level, const sgc_motion_item* items, int nitem,
                          int comp_begin, int ncomp, sgc_plan_t* plan);
int sgc_plan_destroy(sgc_plan_t plan);
int sgc_plan_num_items(sgc_plan_t plan);             /* Copier::numMotionItem               */
int sgc_plan_num_remote_items(sgc_plan_t plan);      /* items with !isLocal() (Copier.H:369) */
long long sgc_plan_num_cells(sgc_plan_t plan);       /* ghost cells filled per exchange      */

/* LevelData::exchange (LevelData.H:455-566): ONE batched launch over the device item table for
 * the local items; remote items go through per-peer staging buffers and NCCL send/recv.     */
int sgc_exchange(sgc_level_t level, sgc_plan_t plan);
/* LevelData::exchangeBegin / exchangeEnd (LevelData.H:577-677): Begin enqueues local copies
 * on the compute stream and pack+send/recv on the comm stream; End makes the compute stream
 * wait for the halo and unpacks.  Work enqueued between them overlaps the transfer.          */
int sgc_exchange_begin(sgc_level_t level, sgc_plan_t plan);
int sgc_exchange_end(sgc_level_t level, sgc_plan_t plan);

/* Zero-gradient physical BC on domain faces: WavePatch::advance BC copies
 * (WavePatch.cpp:156-166) == kernelBC (WavePatch_Cuda.cu:252-264, driverBC :270-283):
 * for every local box touching a domain face, ghost face <- adjacent interior face.         */
int sgc_plan_create_bc_zero_gradient(sgc_level_t level, const int dlo[3], const int dhi[3],
                                     sgc_plan_t* plan);

/* ------------------------------------------------------------------ batched fab copies --- */

/* The application-side copy loops of the reference — LBPatch::stream (LBPatch.H:49-64: one shifted
 * BaseFab::copy per box and component, dst level != src level) and the bounce-back ghost fill of
 * LBLevel::advance (LBLevel.cpp:65-112: shifted copies with component remap inside one fab) — are
 * lists of BaseFab::copy calls.  A copy set records such a list once (geometry of `dst_like` /
 * `src_like`) and sgc_copyset_run replays it in ONE launch on any two levels of those geometries.
 * The calls must be independent (no op reads a cell another op writes), which makes the batched
 * result identical to the reference's sequential loop.                                        */
int sgc_copyset_create(sgc_level_t dst_like, sgc_level_t src_like, const sgc_copy_op* ops, int nop,
                       sgc_copyset_t* set);
int sgc_copyset_run(sgc_copyset_t set, sgc_level_t dst, sgc_level_t src);
int sgc_copyset_destroy(sgc_copyset_t set);
long long sgc_copyset_num_cells(sgc_copyset_t set);

/* ------------------------------------------------------------------ stencil apply -------- */

/* Heat / Jacobi 7-point sweep over every box of the level, one batched launch over the device
 * box table (fills the slot of the empty kernelRHS stub, WavePatch_Cuda.cu:292-301; loop form
 * laplacian.cpp:90-110):
 *   unew = u + f*( ((u[i+1]-2u)+u[i-1]) + ((u[j+1]-2u)+u[j-1]) + ((u[k+1]-2u)+u[k-1]) )
 * on the valid cells of each box; u needs nghost >= 1, filled by sgc_exchange.              */
int sgc_heat_sweep(sgc_level_t unew, sgc_level_t u, double f, unsigned flags);
/* All sweeps take levels of double or of float (the reference's USE_SINGLE_PRECISION build,
 * Parameters.H:17-23) and update every component.                                            */
/* General weighted 7-point stencil on every component:
 *   out = w[0]*u; out += w[1]*u[i-1]; out += w[2]*u[i+1]; out += w[3]*u[j-1]; out += w[4]*u[j+1];
 *   out += w[5]*u[k-1]; out += w[6]*u[k+1]          (in this order; fused multiply-adds with SGC_FMA)
 * — the loop form of the reference's device stencil test (gpuSandbox.cpp:120-151: B = A[-e_z] + A + A[+e_z]
 * per component, on the moving SlabFab window of CudaFab.H:539-905).                          */
int sgc_stencil7_apply(sgc_level_t out, sgc_level_t in, const double weights[7], unsigned flags);
/* Only boxes [box_begin, box_end) — used to overlap interior boxes with the halo transfer.   */
int sgc_heat_sweep_range(sgc_level_t unew, sgc_level_t u, double f, unsigned flags,
                         int box_begin, int box_end);
/* Boxes [a0,a1) and [b0,b1) (a1 <= b0) in ONE launch — the two boundary box layers of a z-slab. */
int sgc_heat_sweep_ranges(sgc_level_t unew, sgc_level_t u, double f, unsigned flags,
                          int a0, int a1, int b0, int b1);

/* The PR1 driver's update (laplacian.cpp:94-108): for iter < niter, for dir in x,y,z:
 *   B -= coef*((A[i+e_dir]-2A[i])+A[i-e_dir])
 * on B's boxes; A's fabs must contain grow(box,1).  Each cell's 3*niter updates are applied in
 * the reference's order in registers (bit-identical to the 3 separate passes).  fuse_iters = 0:
 * one launch per iteration (B read+written every iteration, 24 B/cell-iter); 1: one launch.   */
int sgc_laplacian_apply(sgc_level_t B, sgc_level_t A, double coef, int niter, int fuse_iters,
                        unsigned flags);

/* Leap-frog wave update, scalar form (WavePatch.cpp:226-244; driverRHS WavePatch_Cuda.cu:305-330):
 *   unp1 = 2*un - unm1; for dir: unp1 += factor*((un[+e]-2un)+un[-e])                         */
int sgc_wave_step(sgc_level_t unp1, sgc_level_t un, sgc_level_t unm1, double factor, unsigned flags);

/* nstep x { fill ghosts ; wave step ; rotate } on three device-resident levels — the body of
 * WavePatch::advance repeated without leaving the device, i.e. driverAdvanceIterGroup
 * (WavePatch_Cuda.cu:352-378: kernelBC; kernelRHS; swap indices).  levels[0] = un, [1] = the level
 * to write (unp1), [2] = unm1 on entry; each step applies `exchange_plan` (interior / periodic
 * ghosts; may be NULL) and `bc_plan` (physical-boundary ghosts; may be NULL) to un, then
 * sgc_wave_step.  On return latest[0..2] = indices into levels[] of un, unp1-scratch, unm1.     */
int sgc_wave_run(sgc_level_t levels[3], sgc_plan_t exchange_plan, sgc_plan_t bc_plan, double factor,
                 unsigned flags, int nstep, int latest[3]);

/* nstep x { exchange(plan) ; heat sweep } with ping-pong between a and b, device resident
 * (the shape of driverAdvanceIterGroup, WavePatch_Cuda.cu:352-378).  *result_in_b = 1 when the
 * last-written level is b (always nstep & 1).  With nproc > 1 each step overlaps the halo with
 * interior boxes.  The library is free to restructure the loop — fused ghost fill inside the sweep,
 * two time steps per pass over memory (uniform levels of 64x64xN boxes; periodic or physical boundaries;
 * an odd number of passes goes through a library-owned third level), levels of many small boxes advanced
 * on library-owned levels whose device fabs are 64x64 unions of the caller's boxes (DisjointBoxLayout with
 * maxBoxSize 8/16/32: sgc_plan_rebox_factors), z halos read straight out of the neighbouring GPU's memory
 * — but on return BOTH levels hold, ghost cells included, exactly the bits nstep x { exchange ; sweep }
 * would have left (tests pin this).  The plan keeps up to three extra level arenas alive for this.        */
int sgc_heat_run(sgc_level_t a, sgc_level_t b, sgc_plan_t plan, double f, unsigned flags,
                 int nstep, int* result_in_b);
/* Introspection: how sgc_heat_run re-boxes this plan's level for its step loop (levels of many small
 * boxes, e.g. DisjointBoxLayout(domain, 16): DisjointBoxLayout.cpp:93-168): factor[d] user boxes per
 * device fab along d, all 1 when the level runs as it is.  Decided on the first sgc_heat_run of >= 8 steps;
 * returns 1 once decided, 0 before.                                                                   */
int sgc_plan_rebox_factors(sgc_plan_t plan, int factor[3]);
/* Introspection: the pillar structure (see sgc_step_loop_pillars) the two-step passes of this plan's level use:
 * pillars[0] pillars of pillars[1] boxes; {nbox, 1} when the level is swept box by box.  Returns 0.            */
int sgc_plan_pillars(sgc_plan_t plan, int pillars[2]);

/* ------------------------------------------------------------------ lattice Boltzmann ---- */

/* The per-cell loops of the reference's multi-box application (application/latticeBoltzmann):
 * a 19-component D3Q19 distribution level `fi` and a 4-component macroscopic level `macro`
 * (density, velocity) on the same boxes, nghost = 1.  Arithmetic = the source expressions without
 * contraction: bit-identical to the reference built with -ffp-contract=off, within rounding of
 * its default build.
 * sgc_lb_collide     : LBPatch::collision / LBPhysics::collision (LBPatch.H:27-46, LBPhysics.H:5-14)
 * sgc_lb_macroscopic : LBPatch::macroscopic / LBPhysics::macroscopic (LBPatch.H:9-23, LBPhysics.H:17-35)
 * sgc_lb_copy_ops    : the copy lists of which = 0 LBPatch::stream (LBPatch.H:49-64) and
 *                      which = 1 the bounce-back ghost fill (LBLevel.cpp:65-112) for sgc_copyset_create;
 *                      returns the number of ops (ops == NULL: count only).
 * sgc_lb_run         : nstep x LBLevel::advance (LBLevel.cpp:7-127) device-resident: collision,
 *                      sgc_exchange(exchange) [the caller's PeriodicX|PeriodicY, TrimCorner plan],
 *                      bounce, stream into `prev` + swap, macroscopic.  fuse != 0 replaces
 *                      stream + macroscopic + the NEXT step's collision by one kernel (same bits,
 *                      ghost cells included; `stream` may then be NULL).  *result_in_prev tells
 *                      which level is the reference's m_curr afterwards.                      */
int sgc_lb_collide(sgc_level_t fi, sgc_level_t macro);
int sgc_lb_macroscopic(sgc_level_t macro, sgc_level_t fi);
int sgc_lb_copy_ops(sgc_level_t fi, const int dlo[3], const int dhi[3], int which, sgc_copy_op* ops, int cap);
int sgc_lb_run(sgc_level_t fi, sgc_level_t prev, sgc_level_t macro, sgc_plan_t exchange, sgc_copyset_t bounce,
               sgc_copyset_t stream, int nstep, int fuse, int* result_in_prev);
/* Device self-test of the collision's three-instruction division by the constants
 * which = 0 cs2, 1 2*cs2, 2 2*cs2*cs2, 3 tau (sgc_lb.cu, tests/test_divconst_proof.py): for n
 * HOST values x writes the fast quotient and the IEEE x / c computed on the device.            */
int sgc_lb_selftest_divconst(int which, const double* x, double* fast, double* exact, long long n, double* constant);

/* ------------------------------------------------------------------ multi-GPU ------------ */

/* One process per GPU.  Rank 0 calls sgc_comm_unique_id, the 128 bytes are distributed by the
 * launcher (torch.distributed / a file), every rank calls sgc_comm_init.  Replaces
 * DisjointBoxLayout::initMPI / finalizeMPI / numProc / procID (DisjointBoxLayout.cpp:391-416). */
int sgc_comm_unique_id(char id128[128]);
int sgc_comm_init(int nproc, int rank, const char id128[128]);
int sgc_comm_finalize(void);
int sgc_comm_nproc(void);
int sgc_comm_rank(void);

/* Host-only (no GPU, no NCCL) description of the aggregated halo messages of rank `rank`: for
 * peer number `peer_index` (peers sorted by rank) writes the motion items in wire order.
 * kind 0: what this rank RECEIVES from the peer (its own remote items, own order — the order
 * LevelData.H:549-561 unpacks in); kind 1: what it SENDS (the peer's items that read from this
 * rank, in the peer's order).  Each region travels in linearOut order comp->k->j->i
 * (BaseFab.cpp:365-373).  Returns the item count; SGC_ERR_ARG when peer_index is past the end. */
int sgc_halo_describe(const int dlo[3], const int dhi[3], const int maxbox[3], int nghost,
                      unsigned periodic, unsigned trim, int nproc, int rank, int peer_index,
                      int kind, int* peer_rank, sgc_motion_item* items, int cap);

#ifdef __cplusplus
}
#endif
#endif /* SGC_B200_H */
