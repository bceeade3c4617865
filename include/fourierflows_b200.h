/* libfourierflows_b200.so -- C ABI of the B200-native pseudospectral time-stepping hot path.
 *
 * Drop-in boundary for FourierFlows.jl v0.10.7 (reference paths are relative to /root/reference).
 * The reference has no FFI: its seams are Julia multiple dispatch.  Every entry point
 * below names the reference interface it replaces; INTEGRATION.md shows the `ccall` stub for each.
 *
 * Conventions
 *   - every function returns int: FFB_OK (0) or a negative ffb_status; message via ffb_last_error()
 *   - all arrays are dense, column-major (x / kx fastest), optional trailing field dimension, exactly the
 *     buffers Julia's `Array`/`CuArray` hold; complex = interleaved (re, im) like Complex{T}
 *   - pointers are DEVICE pointers unless the parameter name starts with `host_`
 *   - work is enqueued asynchronously on the library stream (ffb_set_stream); only ffb_d2h and ffb_sync block
 *   - no CPU fallback anywhere: without a CUDA device every compute entry point returns FFB_ECUDA
 */
#ifndef FOURIERFLOWS_B200_H
#define FOURIERFLOWS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  FFB_OK = 0,
  FFB_EINVAL = -1,       /* bad argument */
  FFB_EDOMAIN = -2,      /* odd grid size: Julia DomainError, src/domains.jl:66,179,316 */
  FFB_ENOMEM = -3,
  FFB_ECUDA = -4,
  FFB_ENCCL = -5,
  FFB_EUNSUPPORTED = -6,
  FFB_ESTEPPER = -7      /* step_until! with (Filtered)ETDRK4: src/timesteppers.jl:736-737 */
} ffb_status;

typedef enum { FFB_F32 = 0, FFB_F64 = 1 } ffb_dtype;
typedef enum { FFB_R2C = 0, FFB_C2C = 1 } ffb_kind;

/* Kind of a linear-operator / coefficient operand (`equation.L`, ETD coefficients): src/problem.jl:11-34,
 * src/timesteppers.jl:689-721 (scalar L gives 0-dim coefficients; real L gives real coefficients). */
typedef enum { FFB_COEF_SCALAR = 0, FFB_COEF_REAL = 1, FFB_COEF_COMPLEX = 2 } ffb_coef_kind;

typedef struct {
  const void* ptr; /* dense device array of `dtype` reals (REAL) or complex pairs (COMPLEX); NULL for SCALAR */
  int kind;        /* ffb_coef_kind */
  int dtype;       /* ffb_dtype of the stored coefficient (Float64 ETD coefficients for Float32 problems allowed) */
  double re, im;   /* value when kind == FFB_COEF_SCALAR */
} ffb_coef;

/* Shape of a spectral (or physical) state array plus the grid's alias index ranges
 * (`grid.kalias/kralias/lalias/malias`, src/domains.jl:408-421; 1-based inclusive, lo = 0 means "none"). */
typedef struct {
  int ndim;          /* 1, 2 or 3 grid dimensions */
  int64_t dims[4];   /* extents of the array: (n0, n1, n2, nfields); unused entries = 1 */
  int dtype;         /* ffb_dtype of the real scalar type T; arrays are Complex{T} unless stated */
  int32_t alias_lo[3], alias_hi[3];
} ffb_desc;

typedef struct ffb_plan ffb_plan;
typedef struct ffb_problem ffb_problem;
typedef struct ffb_dist ffb_dist;
typedef struct ffb_snapshot ffb_snapshot;

/* ---------------------------------------------------------------- runtime */
const char* ffb_last_error(void);
int ffb_version(void);
int ffb_device_count(int* n);
int ffb_set_device(int dev);
int ffb_set_stream(void* cuda_stream);          /* NULL = the library's own non-blocking stream */
int ffb_get_stream(void** cuda_stream);
int ffb_sync(void);
int ffb_launch_count(uint64_t* n);              /* number of kernels this library has launched so far */
/* measurement aid (no reference counterpart; SURVEY 5 "tracing / profiling: none"): brackets every kernel launch with
 * CUDA events on the launching stream; the report is a JSON array of {"name","launches","ms","bytes"} per kernel class,
 * `bytes` being the algorithmic HBM traffic */
int ffb_prof_enable(int on);
int ffb_prof_report(char* buf, size_t buflen);

/* B1 array/device seam: `zeros(GPU(), T, dims)` src/utils.jl:80; `device_array(GPU())` src/utils.jl:330;
 * upload `device_array(dev){T}(host)` src/domains.jl:77; download `Array(dev_array)` src/output.jl:79. */
int ffb_malloc(void** dev_ptr, size_t bytes);
int ffb_free(void* dev_ptr);
int ffb_memset_zero(void* dev_ptr, size_t bytes);
int ffb_h2d(void* dev_dst, const void* host_src, size_t bytes);
int ffb_d2h(void* host_dst, const void* dev_src, size_t bytes);   /* blocks until the copy has landed */
int ffb_d2d(void* dev_dst, const void* dev_src, size_t bytes);
int ffb_host_alloc_pinned(void** host_ptr, size_t bytes);
int ffb_host_free_pinned(void* host_ptr);
int ffb_mem_info(size_t* free_bytes, size_t* total_bytes);

/* Asynchronous output path: `saveoutput` (src/output.jl:61-79) downloads every field with a blocking `Array(data)`.
 * A snapshot ring of nbuf (device staging buffer, pinned host buffer) pairs decouples it from the step loop: ffb_snapshot_begin
 * enqueues a device-to-device copy on the library stream (ordered with the steps) and the device-to-host copy on a dedicated copy
 * stream; ffb_snapshot_wait blocks until that slot has landed and returns the pinned host pointer; ffb_snapshot_release frees the
 * slot.  Decomposed fields: every rank snapshots its slab, the launcher gathers the host buffers on rank 0. */
int ffb_snapshot_create(ffb_snapshot** snap, size_t bytes, int nbuf);
int ffb_snapshot_destroy(ffb_snapshot* snap);
int ffb_snapshot_begin(ffb_snapshot* snap, const void* dev_src, size_t bytes, int* slot);
int ffb_snapshot_wait(ffb_snapshot* snap, int slot, void** host_ptr);
int ffb_snapshot_ready(ffb_snapshot* snap, int slot, int* ready);
int ffb_snapshot_release(ffb_snapshot* snap, int slot);

/* ---------------------------------------------------------------- B2 FFT plan protocol
 * `plan_flows_rfft` / `plan_flows_fft` src/domains.jl:2-5, called from the grid constructors :86-87, :207-208,
 * :348-349.  n[] = (nx[, ny[, nz]]) in Julia order (x fastest).  Odd n -> FFB_EDOMAIN.
 * nbatch = number of trailing fields transformed per call (contiguous).  `flags` is a bit set of FFB_PLAN_*. */
#define FFB_PLAN_DEFAULT 0
#define FFB_PLAN_FORCE_GENERIC 1   /* use the arbitrary-size mixed-radix path even for powers of two (testing) */
int ffb_plan_create(ffb_plan** plan, int ndim, const int64_t* n, int dtype, int kind, int nbatch, int flags);
int ffb_plan_destroy(ffb_plan* plan);
int ffb_plan_workspace_bytes(const ffb_plan* plan, size_t* bytes);
/* which code path each 1-D pass uses: writes a NUL-terminated summary (for tests / DESIGN evidence) */
int ffb_plan_describe(const ffb_plan* plan, char* buf, size_t buflen);

/* `mul!(out, plan, in)` (src/diffusion.jl:139,171): forward, unnormalised, sign -1.
 *   R2C: in real (nx,ny,nz[,nbatch]) -> out complex (nx/2+1,ny,nz[,nbatch]); in is preserved.
 *   C2C: in/out complex (nx,ny,nz[,nbatch]); in == out allowed. */
int ffb_fft_forward(ffb_plan* plan, const void* in, void* out);
/* `ldiv!(out, plan, in)` (src/diffusion.jl:137,154-155): inverse scaled by 1/(nx*ny*nz); `in` is preserved
 * (the reference allows c2r to destroy it; preserving is a superset). */
int ffb_fft_inverse(ffb_plan* plan, const void* in, void* out);

/* Fused forms (SURVEY 8b B2 "optional fused forms"; north star: the linear-operator multiply, the dealias! mask and the
 * physical-space products are folded into the FFT passes).  All-power-of-two plans with ndim >= 2 only.
 *   inverse_ex:  out = irfft( F .* in ) .* mul       F[k,l,m] = (cr + i ci) * kx[k] * l[l] * m[m] * w[k,l,m]
 *   forward_ex:  out = dealias!( F .* rfft(in) + G .* acc )   G = (ar + i ai) * akx[k] * al[l] * am[m]
 * forward_ex with square_input != 0 transforms in.^2.  On slab-decomposed plans forward_ex supports square_input, the scalar /
 * wavenumber factors and dealias (l = this rank's slice of the y wavenumbers, alias range 1 = the local range; no acc / w; NCCL and
 * peer-store exchanges).
 * Any vector / array pointer may be NULL (factor 1).  `w`, `acc` have the layout of the spectral array, `mul` of the
 * physical array.  dealias = 1 zeroes the alias box given by alias_lo/hi (as in ffb_desc); dealias = 2 declares the box don't-care:
 * the caller discards it (e.g. the array is only used as `acc` of a later dealiased forward_ex), so the transform may leave it
 * unwritten.  Either way the aliased columns are not stored by the x pass and skipped by the strided passes. */
typedef struct {
  double cr, ci;
  const void *kx, *l, *m, *w;
  const void* acc;
  double ar, ai;
  const void *akx, *al, *am;
  int dealias;
  int32_t alias_lo[3], alias_hi[3];
  const void* mul;
  int square_input;   /* forward_ex: the real input is squared on load (`@. c = c * c` before `mul!`) */
  /* slab-decomposed plans: the GLOBAL 1-based alias ranges (alias_lo/hi hold this rank's local ones); lets the pass that feeds the
   * exchange skip the aliased modes instead of sending them.  All zero: unknown, everything is exchanged. */
  int32_t galias_lo[3], galias_hi[3];
} ffb_fuse;
int ffb_fft_forward_ex(ffb_plan* plan, const void* in, void* out, const ffb_fuse* fuse);
int ffb_fft_inverse_ex(ffb_plan* plan, const void* in, void* out, const ffb_fuse* fuse);
/* n inverse_ex transforms of the SAME spectral array: outs[v] = irfft( F_v .* in ) .* mul_v, v = 0..n-1, evaluated in that order (so
 * fuses[v].mul may be an earlier outs[u]).  calcN! of the 2-D vorticity equation takes zeta, u and v from one `sol`
 * (three `ldiv!` calls on one input, src/diffusion.jl:137 being the single-field case); when the last dimension is long enough for the
 * four-step split (>= 4096) its first sub-pass runs once for all variants and reads `in`, and a dense factor `w` the variants share,
 * from DRAM once.  Costs n scratch arrays of the spectral size on first use.  Same results as n inverse_ex calls. */
int ffb_fft_inverse_multi(ffb_plan* plan, const void* in, int n, void* const* outs, const ffb_fuse* fuses);

/* ---------------------------------------------------------------- multi-GPU slab decomposition
 * No reference counterpart: FourierFlows.jl is single-device (README.md:58, docs/src/gpu.md:57); this is the new
 * capability named by BASELINE.json north_star.  One process per GPU; the caller (torch.distributed, MPI, ...) moves the
 * 128-byte NCCL unique id from rank 0 to the other ranks.  Physical arrays are split along z: rank r holds
 * (nx, ny, nz/P); spectral arrays along y: rank r holds (nx/2+1, ny/P, nz).  ffb_fft_forward / ffb_fft_inverse on a
 * distributed plan perform the local passes plus one exchange (ffb_plan_dist_set_exchange; default: NCCL all-to-all,
 * overlapped chunk by chunk). */
int ffb_dist_unique_id(void* host_id128);
int ffb_dist_init(ffb_dist** dist, int rank, int nranks, const void* host_id128);
int ffb_dist_destroy(ffb_dist* dist);
int ffb_dist_info(const ffb_dist* dist, int* rank, int* nranks);
int ffb_dist_alltoall(ffb_dist* dist, const void* sendbuf, void* recvbuf, size_t block_bytes);
int ffb_plan_create_dist(ffb_plan** plan, int ndim, const int64_t* n, int dtype, ffb_dist* dist, int nchunks /* 0 = default */);
/* Fused pass + collective: the strided pass before the exchange stores its output straight into the peers' receive
 * buffers over NVLink (buffers mapped through CUDA IPC), so the all-to-all disappears and one stream-ordered barrier
 * remains.  Setup: each rank asks its plan for its two receive buffers, exports their IPC handles, the launcher gathers
 * them, every rank opens its peers' handles and hands the mapped pointers (own rank: its local buffer) to the plan. */
int ffb_plan_dist_recv_buffers(ffb_plan* plan, void** buf0, void** buf1, size_t* bytes_each);
int ffb_plan_dist_set_peers(ffb_plan* plan, void* const* peers_buf0, void* const* peers_buf1);   /* arrays of nranks pointers */
/* Receive buffers are pooled per process (peers keep them mapped) and may have served an earlier plan: between
 * ffb_plan_dist_set_peers and the first transform every rank must drain its stream (ffb_sync) and all ranks must meet in a
 * host barrier of the launcher. */
/* Exchange used by a slab-decomposed plan.  NCCL: chunked grouped send/recv (default, needs no peer mapping).
 * PEER_STORE: the fused pass described above; the receive layout is blocked ([kx block][z][y_local][B] forward,
 * [kx block][y][z_local][B] inverse, B complex = 64 bytes) so that a warp's stores are 256 contiguous bytes in the peer's
 * memory; needs 16 <= ny, nz <= 2048 (Float32; 1024 in Float64) and at most 8 ranks -- one NVSwitch domain -- (FFB_EUNSUPPORTED otherwise).  COPY_ENGINE: the passes write
 * destination-major kx-chunks and cudaMemcpyAsync pushes them into the peers' receive buffers while the neighbouring
 * chunks are being transformed (no SM involved); a one-element all-reduce per chunk is the arrival barrier.
 * PEER_STORE and COPY_ENGINE need ffb_plan_dist_set_peers first. */
enum { FFB_EXCHANGE_NCCL = 0, FFB_EXCHANGE_PEER_STORE = 1, FFB_EXCHANGE_COPY_ENGINE = 2 };
int ffb_plan_dist_set_exchange(ffb_plan* plan, int mode);
int ffb_plan_dist_get_exchange(const ffb_plan* plan, int* mode);
/* export: handle of the allocation that holds dev_ptr + offset of dev_ptr inside it; open: maps the allocation once per
 * process (reference counted) and returns base + offset; close: drops one reference */
int ffb_dist_ipc_export(void* dev_ptr, void* host_handle64, size_t* offset);
int ffb_dist_ipc_open(const void* host_handle64, size_t offset, void** dev_ptr);
int ffb_dist_ipc_close(void* dev_ptr);
int ffb_dist_barrier(ffb_dist* dist);

/* ---------------------------------------------------------------- grid-side kernels (src/domains.jl)
 * wavenumber vectors `k, l, m, kr` (:77-78,193-195,333-336): fftfreq/rfftfreq computed in Float64, stored as T. */
int ffb_wavenumbers(void* out, int64_t n, double L, int dtype, int real_half /* 1 = rfftfreq */);
/* dense `Ksq`/`Krsq` (= kx^2 + l^2 + m^2 in T arithmetic) and `invKsq`/`invKrsq` with [1,1,1] = 0 (:80-83,197-203,338-344).
 * kx, l, m are device vectors of length dims[0..2] (l, m NULL when ndim < 2, 3).  Either output may be NULL. */
int ffb_ksq(void* ksq, void* invksq, const void* kx, const void* l, const void* m, const ffb_desc* desc);
/* `dealias!(fh, grid)` (:428-476): zero fh[alias_x,:,:,:], fh[:,alias_y,:,:], fh[:,:,alias_z,:].  The caller puts
 * kralias or kalias into alias_lo/hi[0] according to `size(fh,1) == grid.nkr` (:437,450,464). */
int ffb_dealias(void* fh, const ffb_desc* desc);
/* `makefilter(grid, T, dims; order, innerK, outerK, tol)` (:506-546): dense real filter of desc->dims.
 * kx/l/m as in ffb_ksq (pass kr for real-variable spectra, k otherwise); dx, dy, dz = grid spacings. */
int ffb_make_filter(void* filter, const void* kx, const void* l, const void* m, double dx, double dy, double dz,
                    double order, double innerK, double outerK, double tol, const ffb_desc* desc);

/* ---------------------------------------------------------------- B3 stepper stages (src/timesteppers.jl)
 * n = number of complex elements of `sol` (prod(eq.dims)); T = desc dtype; L / coefficients as ffb_coef.
 * filter (dense real T, may be NULL) folds `@. sol *= filter` (:279,408,552,658) into the final stage. */
/* ETD coefficient precompute: `getexpLs` :673-678 and `getetdcoeffs` :689-721 (32-point contour mean in
 * Complex{Float64}).  Outputs have L's kind (real/complex) and `coef_dtype` storage; n elements (1 for scalar L,
 * then results are written to host_scalars[12] as (re,im) pairs of expLdt, exphLdt, zeta, alpha, beta, gamma). */
int ffb_etd_coeffs(double dt, const ffb_coef* L, int dtype, int coef_dtype, int64_t n, void* expLdt, void* exphLdt,
                   void* zeta, void* alpha, void* beta, void* gamma, double* host_scalars);
/* ForwardEuler :113 `sol += dt*(L*sol + N)`; with filter :144 `sol = filter*(sol + dt*(N + L*sol))` */
int ffb_stage_fe(void* sol, const void* N, const ffb_coef* L, double dt, const void* filter, int dtype, int64_t n);
/* RK4 :225-264.  `rhs += L*u` (addlinearterm!) fused with the next substep `sol1 = sol + c*rhs` (substepsol!). */
int ffb_stage_rk4_substep(void* sol1, void* rhs, const void* u, const void* sol, const ffb_coef* L, double c,
                          int dtype, int64_t n);
/* last RK4 stage: `rhs4 += L*sol1` then `sol += dt/6*(rhs1 + 2 rhs2 + 2 rhs3 + rhs4)` [`*= filter`]; rhs4 is updated
 * in memory only when store_rhs4 != 0 (the reference leaves RHS4 = N4 + L*sol1 in ts.RHS4). */
int ffb_stage_rk4_final(void* sol, const void* rhs1, const void* rhs2, const void* rhs3, void* rhs4, const void* sol1,
                        const ffb_coef* L, double dt, const void* filter, int store_rhs4, int dtype, int64_t n);
/* LSRK54 stage i :386-392: `rhs += L*sol; S2 = A*S2 + dt*rhs; sol += B*S2` [`*= filter` on the last stage].
 * first != 0 treats S2 as zero on input (folds `@. S2 = 0`, :384). */
int ffb_stage_lsrk54(void* sol, void* S2, void* rhs, const ffb_coef* L, double A, double B, double dt, int first,
                     const void* filter, int dtype, int64_t n);
/* ETDRK4 :501-516 */
int ffb_stage_etdrk4_substep12(void* out, const ffb_coef* exphLdt, const void* sol, const ffb_coef* zeta, const void* N,
                               int dtype, int64_t n);
int ffb_stage_etdrk4_substep3(void* out, const ffb_coef* exphLdt, const void* sol1, const ffb_coef* zeta, const void* N1,
                              const void* N3, int dtype, int64_t n);
int ffb_stage_etdrk4_update(void* sol, const ffb_coef* expLdt, const ffb_coef* alpha, const ffb_coef* beta,
                            const ffb_coef* gamma, const void* N1, const void* N2, const void* N3, const void* N4,
                            const void* filter, int dtype, int64_t n);
/* AB3 :628-651: `rhs += L*sol`; Euler when step < 3 else AB3 update; [`*= filter`].  History rotation is done by the
 * caller swapping pointers (the reference copies RHS -> RHS_1 -> RHS_2 after the clock tick). */
int ffb_stage_ab3(void* sol, void* rhs, const void* rhs_m1, const void* rhs_m2, const ffb_coef* L, double dt,
                  int64_t step, const void* filter, int dtype, int64_t n);

/* ---------------------------------------------------------------- elementwise vocabulary for user calcN!
 * (the broadcast shapes of src/diffusion.jl:136-140 and of the 2-D vorticity / 3-D Burgers test equations) */
/* out[i] = a * x[i] + b * y[i] over n elements of type T (real if is_complex == 0); y may be NULL (b ignored) */
int ffb_ew_axpby(void* out, double a, const void* x, double b, const void* y, int is_complex, int dtype, int64_t n);
/* out = x * y (real arrays, physical space products such as `cx *= kappa`, `u *= zeta`) */
int ffb_ew_mul_real(void* out, const void* x, const void* y, int dtype, int64_t n);
/* out[k,l,m,f] = (ar + i*ai) * kx[k]^px * l[l]^py * m[m]^pz * w[k,l,m] * in[k,l,m,f]; w (dense real) may be NULL;
 * accumulate != 0 adds into out; dealias != 0 zeroes the aliased box afterwards (desc alias ranges). */
int ffb_ew_spectral_mul(void* out, const void* in, double ar, double ai, const void* kx, int px, const void* l, int py,
                        const void* m, int pz, const void* w, int accumulate, int dealias, const ffb_desc* desc);

/* out = x * y with x real or complex and y real or complex arrays of the same shape (`a .* by` for complex fields in `jacobianh`,
 * src/utils.jl:198-203; a real factor multiplies both parts, as Julia's real*complex does) */
int ffb_ew_mul(void* out, const void* x, int x_complex, const void* y, int y_complex, int dtype, int64_t n);
/* `jacobianh(a, b, grid)` for real fields on a TwoDGrid (src/utils.jl:190-197):
 *   out = im*kr .* rfft(a .* irfft(im*l .* rfft(b))) - im*l .* rfft(a .* irfft(im*kr .* rfft(b)))
 * as five transforms with every multiply folded into a pass (no elementwise kernel).  rfftplan: the grid's 2-D r2c plan; a, b: real
 * (nx, ny); out, scratch_h: complex (nx/2+1, ny); scratch_p1, scratch_p2: real (nx, ny); kr, l: the grid's wavenumber vectors. */
int ffb_jacobianh(ffb_plan* rfftplan, void* out, const void* a, const void* b, const void* kr, const void* l, void* scratch_h,
                  void* scratch_p1, void* scratch_p2);

/* ---------------------------------------------------------------- diagnostics (src/utils.jl:113-183) */
/* `parsevalsum2(uh, grid)` / `parsevalsum(uh, grid)` partial: returns Sum over modes with the half-spectrum
 * double counting of 0 < k < nx/2 when half != 0; the caller applies the L/n^2 normalisation. */
int ffb_parseval_sum(double* host_result_re, const void* uh, int abs2, int half, const ffb_desc* desc);

/* ---------------------------------------------------------------- C-driven problem (bench / Julia `@cfunction` form)
 * Mirrors `Problem(eqn, stepper, dt, grid, vars, params)` src/problem.jl:99-111 and `stepforward!` src/timesteppers.jl:6-35. */
typedef int (*ffb_calcN_fn)(void* N, const void* sol, double t, void* user);

typedef enum {
  FFB_CALCN_CALLBACK = 0,
  FFB_CALCN_ZERO = 1,        /* `@. N = 0` src/diffusion.jl:129-133 */
  FFB_CALCN_DIFFUSION = 2,   /* array-kappa diffusion, src/diffusion.jl:135-143 */
  FFB_CALCN_VORTICITY2D = 3, /* 2-D Navier-Stokes advection + dealias! */
  FFB_CALCN_BURGERS3D = 4    /* N = -1/2 i kr rfft(irfft(sol)^2) + dealias! */
} ffb_calcN_kind;

typedef enum {
  FFB_FORWARD_EULER = 0, FFB_RK4 = 1, FFB_LSRK54 = 2, FFB_ETDRK4 = 3, FFB_AB3 = 4
} ffb_stepper_kind;

typedef struct {
  int ndim;
  int64_t n[3];          /* physical grid size (nx, ny, nz) */
  double L[3];           /* domain extents (Lx, Ly, Lz) */
  int dtype;             /* T */
  double aliased_fraction;
  int stepper;           /* ffb_stepper_kind */
  int filtered;          /* Filtered* variant; filter parameters below */
  /* makefilter keywords (src/domains.jl:506).  Each field has its own sentinel: filter_order, filter_outerK, filter_tol <= 0 select the
   * reference defaults (4, 1, 1e-15).  filter_innerK < 0 selects 2/3; filter_innerK == 0 means 0 when filter_outerK or filter_tol is
   * given (> 0) and the default 2/3 in an all-zero (memset) configuration. */
  double filter_order, filter_innerK, filter_outerK, filter_tol;
  double dt;
  int calcN;             /* ffb_calcN_kind */
  ffb_calcN_fn callback; void* user;
  double nu;             /* L = -nu * Krsq (dense real); nu_scalar_L != 0 uses scalar L = 0 instead */
  int scalar_zero_L;
  const void* kappa;     /* device real array (nx) for FFB_CALCN_DIFFUSION */
  int coef_dtype;        /* storage of ETD coefficients: FFB_F64 (reference-faithful) or dtype */
  int fused;             /* 1: fold spectral multiplies / products / dealias into FFT passes where implemented */
  ffb_dist* dist;        /* non-NULL: slab-decomposed 3-D problem; all arrays are the local slabs (see ffb_plan_create_dist) */
} ffb_problem_config;

int ffb_problem_create(ffb_problem** prob, const ffb_problem_config* cfg);
int ffb_problem_destroy(ffb_problem* prob);
int ffb_problem_sol(ffb_problem* prob, void** sol, int64_t* n_complex);   /* device pointer of `prob.sol` */
int ffb_problem_plan(ffb_problem* prob, ffb_plan** plan);                 /* `prob.grid.rfftplan` (borrowed handle) */
int ffb_problem_clock(ffb_problem* prob, double* t, int64_t* step, double* dt);
int ffb_problem_set_dt(ffb_problem* prob, double dt);
int ffb_problem_bytes(ffb_problem* prob, size_t* device_bytes);
/* `set_c!`-like: real physical field (host) -> sol = rfft(field) */
int ffb_problem_set_physical(ffb_problem* prob, const void* host_field);
int ffb_problem_get_physical(ffb_problem* prob, void* host_field);
/* `stepforward!(prob, nsteps)` src/timesteppers.jl:14-20 */
int ffb_step(ffb_problem* prob, int64_t nsteps);
/* `step_until!(prob, stop_time)` src/timesteppers.jl:734-760 (bug-compatible final step, FFB_ESTEPPER for ETDRK4) */
int ffb_step_until(ffb_problem* prob, double stop_time);

/* Host-buffer pipeline (no reference counterpart: with CUDA.jl the user writes `prob.sol .= device_array(dev)(host)`,
 * `stepforward!`, `Array(prob.sol)` -- three blocking calls, src/utils.jl:330, src/timesteppers.jl:14, src/output.jl:79 -- and the GPU
 * idles during both copies).  Each submission is an independent spectral state in pinned host memory (an ensemble member, a request):
 * it is uploaded on a copy-in stream, moved into `sol`, stepped `nsteps` times on the library stream and downloaded into `host_out`
 * on a copy-out stream, so that the copies of neighbouring submissions overlap the steps of this one.  `depth` = submissions in
 * flight (device staging: 2 x depth spectral arrays).  ffb_pipeline_submit blocks only when the ring is full; `host_out` is complete
 * after ffb_pipeline_wait(ticket).  Results equal h2d + ffb_step + d2h bit for bit.  FFB_EUNSUPPORTED for AB3 (history). */
typedef struct ffb_pipeline ffb_pipeline;
int ffb_pipeline_create(ffb_pipeline** pipe, ffb_problem* prob, int depth);
int ffb_pipeline_destroy(ffb_pipeline* pipe);
int ffb_pipeline_submit(ffb_pipeline* pipe, const void* host_in, void* host_out, int64_t nsteps, int* ticket);
int ffb_pipeline_wait(ffb_pipeline* pipe, int ticket);

#ifdef __cplusplus
}
#endif
#endif /* FOURIERFLOWS_B200_H */
