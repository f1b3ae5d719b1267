/*
 * qrb200.h -- C ABI of libqrb200.so: B200-native (sm_100a) tiled Householder QR.
 *
 * Drop-in boundary for the GPU path of s10m/GPU-Tiled-QR-Decomposition.  The four
 * legacy symbols of the reference's include/gpucalc.h are re-exported with the same
 * signatures (see include/gpucalc.h in this repo); the b200qr_* entry points are the
 * extended, device-pointer interface the legacy calls are thin wrappers over.
 *
 * Conventions (identical to the reference, qrdecomp.c:18 / gpucalc.cu:9):
 *   - matrices are column-major float32, element (i,j) at mat[j*ldm + i];
 *   - the tile size is 32 (baked into the reference, gpucalc.cu:1582, qrdecomp.c:506);
 *     m and n must be multiples of 32; p = m/32 tile rows, q = n/32 tile columns;
 *   - output, in place: R in the global upper triangle; strict lower part of diagonal
 *     tile (k,k) = GEQRT reflectors (unit diagonal implied); every tile (i,k), i>k,
 *     = the full 32x32 TSQRT reflector block V2 (reference rows a1-a4 of SURVEY 8);
 *   - tau: one scalar per reflector, tau = 2/(1+|v_ess|^2).  The reference keeps it in
 *     an m x n matrix whose tile (i,k) holds tau_0..tau_31 in rows 0..31 of the tile's
 *     first column (qrdecomp.c:486,686).  This library stores it compactly,
 *     tau_c[(k*p + i)*32 + r], and b200qr_tau_to_reference() expands it.
 *
 * All functions return 0 (B200QR_OK) or a negative B200QR_E* code; none prints,
 * none resets the device.  No torch types: plain pointers and sizes only.
 */
#ifndef QRB200_H
#define QRB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200QR_TILE 32

enum {
    B200QR_OK = 0,
    B200QR_EINVAL = -1,      /* bad shape / pointer / option                       */
    B200QR_ECUDA = -2,       /* a CUDA runtime call failed (b200qr_last_cuda_error) */
    B200QR_ENOMEM = -3,
    B200QR_ESTALL = -4       /* a wait inside the persistent kernel made no progress for QRB200_WATCHDOG_MS
                                (default 60 000): the kernel trapped; reported by the synchronous entry points */
};

/* Task vocabulary of the reference (include/gridscheduler.h:10-17). */
enum { B200QR_QRS = 0, B200QR_SAPP = 1, B200QR_QRD = 2, B200QR_DAPP = 3 };
enum { B200QR_TASK_AVAIL = 0, B200QR_TASK_NONE = 1, B200QR_TASK_DONE = 2 };

typedef struct {
    int type;   /* B200QR_QRS .. B200QR_DAPP          (Task.taskType) */
    int l;      /* tile row                            (Task.l)        */
    int m;      /* tile column                         (Task.m)        */
    int k;      /* step                                (Task.k)        */
} b200qr_task;

typedef struct {
    int grid_ctas;      /* persistent grid size; 0 = all co-resident CTAs of the device     */
    int trace;          /* 1 = record a per-tile-task start/finish sequence (DAG tests);
                           2 = record %globaltimer nanoseconds (low 32 bits) instead            */
    int share_device;   /* b200qr_dist_create only: number of rank objects that will share this device
                           and be launched together with b200qr_dist_factor_group (0/1 = none)    */
    int engine;         /* 0 = automatic (a function of m, n and the world size only), 1 = FP32 SIMT sweeps,
                           2 = tensor-core sweeps (3xTF32 compact-WY products on tcgen05).  Engine 2 needs an
                           operand workspace of 6 x the factored columns: if it does not fit, plan creation
                           fails with B200QR_ENOMEM -- the library never switches engine behind the caller. */
    int reserved[4];
} b200qr_opts;

typedef struct b200qr_plan_s *b200qr_plan_t;

/* ---- library / device info ------------------------------------------------------ */
int         b200qr_version(void);
const char *b200qr_strerror(int code);
int         b200qr_last_cuda_error(void);        /* cudaError_t of the last B200QR_ECUDA */
int         b200qr_device_sm_count(void);
/* Synchronous copy between any two host/device buffers (cudaMemcpyDefault): lets a caller fill and
 * read the library-owned matrices of the multi-GPU API without linking the CUDA runtime itself. */
int         b200qr_copy(void *dst, const void *src, size_t bytes);
/* Measured FP32 FFMA throughput of the current device in GFLOP/s (register-resident FMA loop
 * on every SM, best of 3 after a warm-up): the roofline denominator of the SIMT kernels.   */
int         b200qr_measure_fp32_peak(double *gflops, int iters);

/* ---- plan: workspace + schedule for one (m, n) shape ----------------------------
 * Replaces the per-call cudaMalloc of task grid / queue / tau in cudaQRTask
 * (gpucalc.cu:1601-1612).  A plan owns: per-tile progress words (the dependency
 * counters), compact tau, the step table of the persistent scheduler.             */
int  b200qr_plan_create(b200qr_plan_t *plan, int m, int n, const b200qr_opts *opts);
int  b200qr_plan_destroy(b200qr_plan_t plan);
int  b200qr_plan_grid(b200qr_plan_t plan);               /* CTAs the kernel launches with */
/* After B200QR_ESTALL (or a launch failure the caller's own synchronize reported): where the first wait that gave up was
 * stuck.  out5 = {valid, block, warp, tag, arg}; tag = shared-memory address of the mbarrier it waited on (arg = parity),
 * or 0xF0000000 + n for a progress-word wait (arg = the value waited for).  The note lives in pinned host memory, so it
 * can be read although the CUDA context is gone. */
int  b200qr_plan_stall_info(b200qr_plan_t plan, unsigned *out5);
/* Every warp that gave up leaves a note and waits one more second before it traps, so all the stuck waits are on record:
 * out[2 s] = tag, out[2 s + 1] = arg | 0x80000000 for slot s = 16 * block + warp (0 = none); returns the slots copied. */
int  b200qr_plan_stall_notes(b200qr_plan_t plan, unsigned *out, int cap);
int  b200qr_plan_debug_counters(b200qr_plan_t plan, unsigned long long *out, int n);   /* opts.trace == 3: phase timers of the tensor-core engine */
int  b200qr_plan_variant(b200qr_plan_t plan);            /* 1, 2: FP32 SIMT sweeps (1 / 2 matrix columns per thread); 5: tcgen05 sweeps */
long b200qr_plan_tile_tasks(b200qr_plan_t plan);         /* sum_k (p-k)(q-k)  (gpucalc.cu:1546-1559) */
long b200qr_plan_macro_tasks(b200qr_plan_t plan);        /* panel + column-sweep tasks actually queued */

/* ---- the hot path: factor a device-resident matrix in place ---------------------
 * Replaces taskKernel + cuda_initScheduling (gpucalc.cu:1466-1544, launched at
 * 1627/1638).  dA: device pointer, column-major, ldm >= m, ldm % 4 == 0, 16-byte
 * aligned.  Asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy
 * default stream).  One persistent kernel launch + one small memset per call.      */
int b200qr_factor_device(b200qr_plan_t plan, float *dA, int ldm, void *stream);

/* Compact tau (device pointer owned by the plan, p*min(p,q)*32 floats). */
const float *b200qr_plan_tau_device(b200qr_plan_t plan);
/* Expand compact tau into the reference's m x n tau-matrix layout (device, ld = ldtau);
 * entries outside the tau vectors are set to 0 (qrdecomp.c:90 pre-zeroes them). */
int b200qr_tau_to_reference(b200qr_plan_t plan, float *dTauMat, int ldtau, void *stream);

/* Host-buffer convenience: H2D, factor, D2H (what cudaQRTask does, gpucalc.cu:1614-1670)
 * without the prints and without cudaDeviceReset.  hTauMat may be NULL; if given it
 * receives tau in the reference layout (leading dimension ldm).  Synchronous.
 * kernel_ms (optional) receives the device time of the factorisation only.         */
int b200qr_factor_host(float *hA, int m, int n, int ldm, float *hTauMat, float *kernel_ms);

/* ---- multi-GPU: 1-D block-cyclic tile columns, one rank (process) per GPU -----------------------
 * The reference is single-GPU; this is the north-star extension.  Tile columns are dealt to the
 * ranks in blocks of b200qr_dist_block() tile columns (block g to rank g % world).  Every rank
 * runs the same persistent kernel on its own GPU: the panels of its tile columns and the sweeps
 * of its blocks.  The reflector tiles, tau and panel progress words of a step whose panel lives
 * on another rank are read directly from that rank's memory over NVLink inside the kernel (peer
 * mappings obtained with CUDA IPC), so there is no host-side broadcast step.
 *
 * Protocol (all ranks):  create -> export -> [exchange the handle bytes, e.g. all_gather] ->
 * connect -> fill the owned columns of b200qr_dist_matrix() -> reset -> [barrier] -> factor ->
 * [stream sync] -> [barrier].  Results (R and V of the owned tile columns, tau of the owned panels)
 * stay in the rank's matrix / tau buffer.  The closing barrier matters when the factorisation is
 * repeated: until EVERY rank's kernel has finished, peers still read this rank's reflector tiles and
 * progress words (SIMT engine) or still write operand tiles and panel progress words into this
 * rank's buffers (tensor-core engine), so a rank must not refill its matrix or reset before it.   */
typedef struct b200qr_dist_s *b200qr_dist_t;
int    b200qr_dist_create(b200qr_dist_t *d, int m, int n, int rank, int world, const b200qr_opts *opts);
int    b200qr_dist_destroy(b200qr_dist_t d);
int    b200qr_dist_handle_bytes(void);                       /* size of one rank's exported handle blob */
int    b200qr_dist_export(b200qr_dist_t d, void *out_bytes);
int    b200qr_dist_connect(b200qr_dist_t d, const void *all_bytes);   /* world blobs, rank order */
float *b200qr_dist_matrix(b200qr_dist_t d);                  /* device pointer, m x n, leading dimension m */
const float *b200qr_dist_tau_device(b200qr_dist_t d);        /* compact tau [(k*p + i)*32 + r], owned panels */
int    b200qr_dist_block(b200qr_dist_t d);                   /* tile columns per distribution block */
int    b200qr_dist_owner(b200qr_dist_t d, int tile_col);     /* rank that owns a tile column */
int    b200qr_dist_grid(b200qr_dist_t d);
int    b200qr_dist_stall_info(b200qr_dist_t d, unsigned *out5);   /* like b200qr_plan_stall_info */
int  b200qr_dist_variant(b200qr_dist_t d);             /* kernel variant of this rank's plan (2: SIMT sweeps, 5: tensor-core engine) */
int    b200qr_dist_read_trace(b200qr_dist_t d, int *out6, long cap);   /* opts.trace; tasks of other ranks read -1 */
long   b200qr_dist_tile_tasks(b200qr_dist_t d);
int    b200qr_dist_reset(b200qr_dist_t d, void *stream);
int    b200qr_dist_factor(b200qr_dist_t d, void *stream);
/* Same-process variants for rank objects that share ONE device (tests / fewer GPUs than ranks):
 * wire them with plain pointers and run all of them in a single cooperative launch (world <= 2). */
/* A rank's factorisation with HOST data, the copies overlapped with the kernel (the multi-GPU sibling of cudaQRTask's
 * overlapped path; the reference has no multi-GPU entry point).  Every rank: reset_host (progress words = "input not
 * here yet"; synchronous) -> [barrier] -> factor_host, which launches the rank's kernel, streams the OWNED tile columns in
 * behind it, copies every finished block of tile columns back while the kernel works on the rest, and returns when
 * they are all back.  h: column-major host matrix (pinned for the overlap to happen), leading dimension ldh >= m;
 * packed = 0: the whole m x n matrix, of which only this rank's tile columns are read and written; packed = 1: this
 * rank's tile columns only, consecutive in increasing order.  kernel_ms (optional): the kernel's run time, which here
 * includes waiting for its input.  As above, a barrier belongs between factor_host and the next reset_host. */
int    b200qr_dist_reset_host(b200qr_dist_t d);
int    b200qr_dist_factor_host(b200qr_dist_t d, float *h, long long ldh, int packed, float *kernel_ms);
int    b200qr_dist_connect_local(b200qr_dist_t *all, int world);
int    b200qr_dist_factor_group(b200qr_dist_t *all, int world, void *stream);

/* ---- batched small problems and TSQR (BASELINE configs[2], configs[3]) ---------------------------
 * nprob independent m x n problems with n <= 256 (at most 8 tile columns), problem b at
 * dA + b*strideA, each factored in place exactly like cudaQRTask would factor it (same tile
 * algorithm and output layout).  dTau receives nprob * K*p*32 floats (compact tau per problem,
 * [(k*p + i)*32 + r]); dCounter is one int of device scratch.  Asynchronous on `stream`.       */
int b200qr_batched_workspace_floats(int m, int n, int nprob, long long *tau_floats);
int b200qr_batched_factor(float *dA, int m, int n, int ldm, long long strideA, int nprob, float *dTau, int *dCounter,
                          void *stream);
int b200qr_batched_last_cuda_error(void);
/* Tall-skinny QR by a reduction tree: rows cut into leaves of leaf_rows (m % leaf_rows == 0), every
 * leaf factored as one batched problem, the R factors stacked and reduced `fanin` at a time until
 * one n x n R is left.  The reference handles a tall panel as one serial TSQRT chain
 * (gridscheduler.c:136-154); the tree changes the Householder-vector layout (leaf reflectors stay in
 * dA, inner nodes in the plan), so parity for this path is on R up to row signs.  For several
 * GPUs: one plan per row block, all-gather the R factors (NCCL), reduce the stack on every rank with
 * b200qr_tsqr_create_rstack (host side: TsqrDist in the Python package, tools/bench_tsqr_dist.py). */
typedef struct b200qr_tsqr_s *b200qr_tsqr_t;
int b200qr_tsqr_create(b200qr_tsqr_t *t, long long m, int n, int leaf_rows, int fanin);
/* Streaming engine for the same job (same b200qr_tsqr_factor / _destroy): `strips` CTAs (<= 0: two per SM) each walk a
 * contiguous strip of rows in blocks of b200qr_tsqr_stream_block_rows(t) rows (48) -- the reference's TSQRT recurrence
 * (qrdecomp.c:689-721) with the tile as wide as the matrix, thread = matrix column, the strip's n x n triangle in L2 --
 * and the strips' R factors are reduced by the leaf-batch tree, four at a time.  Reflectors stay in dA (one block of
 * rows x n per TSQRT step, strips of ceil(m / strips) rows rounded up to whole blocks) and in the plan (tree). */
int b200qr_tsqr_create_streaming(b200qr_tsqr_t *t, long long m, int n, int strips);
/* Tile engine for the same job: `leaves` equal leaves (<= 0: sixteen per SM, rounded down to a divisor of m), each factored
 * by ONE WARP of the batched tile kernel as a flat tree of 32 x 32 TSQRT tiles per tile column (the reference's own chain,
 * gridscheduler.c:136-154, per leaf; leaf reflectors in dA in the reference's tile layout), the leaves' R factors reduced
 * pairwise by the streaming engine's triangular levels. */
int b200qr_tsqr_create_tiles(b200qr_tsqr_t *t, long long m, int n, int leaves);
int b200qr_tsqr_leaves(b200qr_tsqr_t t);             /* leaves (tile engine) or strips (streaming engine) */
int b200qr_tsqr_stream_block_rows(b200qr_tsqr_t t);
/* Reduction of `count` stacked n x n upper-triangular factors ((count*n) x n, column-major) to one by the streaming
 * engine's pairwise levels: the top of the multi-GPU TSQR (all-gather of the ranks' R factors, then every
 * rank reduces the stack).  b200qr_tsqr_factor(t, dStack, ld, dR) overwrites the stack with the tree's reflectors. */
int b200qr_tsqr_create_rstack(b200qr_tsqr_t *t, int count, int n);
int b200qr_tsqr_destroy(b200qr_tsqr_t t);
int b200qr_tsqr_levels(b200qr_tsqr_t t);
int b200qr_tsqr_factor(b200qr_tsqr_t t, float *dA, long long ldm, float *dR, void *stream);

/* ---- TSQR over several GPUs behind the C ABI (BASELINE.json configs[3] at 2 / 4 / 8 GPUs) -------------------------
 * One rank (process) per GPU; rank r owns rows [r * m_local, (r + 1) * m_local).  factor(): local reduction -> the R
 * factors are exchanged over peer memory (each rank copies its n x n factor into every rank's stack buffer over NVLink
 * and raises an epoch flag there; nobody calls a collective library or synchronises on the host) -> every rank reduces
 * the stack of W factors, so all ranks hold the same R.  Protocol: create -> export -> [all-gather the handle bytes]
 * -> connect -> factor (any number of times, asynchronous on `stream`).  tiles != 0 selects the tile engine for the
 * local reduction, 0 the streaming engine.  The reflectors stay in dA (local), in the plans (trees).            */
typedef struct b200qr_tsqr_dist_s *b200qr_tsqr_dist_t;
int b200qr_tsqr_dist_create(b200qr_tsqr_dist_t *h, long long m_local, int n, int rank, int world, int tiles);
int b200qr_tsqr_dist_destroy(b200qr_tsqr_dist_t h);
int b200qr_tsqr_dist_handle_bytes(void);
int b200qr_tsqr_dist_export(b200qr_tsqr_dist_t h, void *out_bytes);
int b200qr_tsqr_dist_connect(b200qr_tsqr_dist_t h, const void *all_bytes);            /* world blobs, rank order */
int b200qr_tsqr_dist_connect_local(b200qr_tsqr_dist_t *all, int world);               /* rank objects of one process */
int b200qr_tsqr_dist_factor(b200qr_tsqr_dist_t h, float *dA, long long ldm, float *dR, void *stream);
/* the two halves of factor(), for a single process that drives several rank objects (all begins before any end) */
int b200qr_tsqr_dist_factor_begin(b200qr_tsqr_dist_t h, float *dA, long long ldm, void *stream);
int b200qr_tsqr_dist_factor_end(b200qr_tsqr_dist_t h, float *dR, void *stream);
const float *b200qr_tsqr_dist_local_r(b200qr_tsqr_dist_t h);                          /* the rank's own n x n factor */

/* ---- apply Q (the reference has no such routine, qrdecomp.c:1292 multAB is unused) ---
 * C <- Q^T C (transpose != 0) or Q C for the Q held implicitly by a factored matrix: dA = factored
 * m x n matrix (ldm), dTauCompact = [(k*p + i)*32 + r] (b200qr_plan_tau_device), dC = m x ncols
 * column-major (ldc), ncols % 32 == 0.  Asynchronous on `stream`.                               */
int b200qr_apply_q(const float *dA, int m, int n, int ldm, const float *dTauCompact, float *dC, int ldc, int ncols,
                   int transpose, void *stream);

/* Q(:, 0:ncols) of a factored matrix, explicitly: dQ (m x ncols, ldq) <- Q I(:, 0:ncols); ncols % 32 == 0, ncols <= m. */
int b200qr_form_q(const float *dA, int m, int n, int ldm, const float *dTauCompact, float *dQ, int ldq, int ncols, void *stream);
/* X <- R^-1 X for the n x n upper triangle R of a factored matrix (dR = its first element, ldr): dB is n x nrhs (ldb). */
int b200qr_trsm_upper(const float *dR, int n, int ldr, float *dB, int ldb, int nrhs, void *stream);
/* Least squares min |A x - b| with the factored A (m >= n, full column rank): dB (m x nrhs, ldb, nrhs % 32 == 0) is
 * overwritten with Q^T B, whose first n rows are then overwritten with the solution X = R^-1 (Q^T B)(0:n, :). */
int b200qr_solve_ls(const float *dA, int m, int n, int ldm, const float *dTauCompact, float *dB, int ldb, int nrhs, void *stream);

/* ---- arbitrary sizes: the reference truncates m and n to multiples of 32 (qrdecomp.c:162,
 * gpucalc.cu:1582); this front-end embeds the m x n matrix in a zero-padded (mp x np) device buffer owned by the handle
 * (mp, np = m, n rounded up to 32), factors that, and serves R, Q and least-squares solves for the REAL sizes.  The m x n
 * part of the factored buffer is the factorisation of A: padded rows carry zero reflector entries, padded columns the
 * reference's zero-column case (qrdecomp.c:1219).  b200qr_any_factored / _tau expose the padded result (ld = mp) in the
 * library's usual layout.  Every pointer is a device pointer; calls are asynchronous on `stream`. */
typedef struct b200qr_any_s *b200qr_any_t;
int  b200qr_any_create(b200qr_any_t *h, int m, int n, const b200qr_opts *opts);
int  b200qr_any_destroy(b200qr_any_t h);
int  b200qr_any_padded_rows(b200qr_any_t h);
int  b200qr_any_padded_cols(b200qr_any_t h);
int  b200qr_any_factor(b200qr_any_t h, const float *dA, int ldm, void *stream);          /* dA: m x n, not modified */
int  b200qr_any_get_r(b200qr_any_t h, float *dR, int ldr, void *stream);                /* min(m,n) x n, upper */
int  b200qr_any_apply_q(b200qr_any_t h, float *dC, int ldc, int ncols, int transpose, void *stream);   /* C: m x ncols */
int  b200qr_any_solve_ls(b200qr_any_t h, const float *dB, int ldb, int nrhs, float *dX, int ldx, void *stream);
const float *b200qr_any_factored(b200qr_any_t h);
const float *b200qr_any_tau(b200qr_any_t h);
int  b200qr_frontend_last_cuda_error(void);

/* ---- other tile sizes (BASELINE.json configs[1]: tile-size sweep 32 / 64 / 128) -------------------------------------
 * The reference bakes 32 into both paths (gpucalc.cu:1582, qrdecomp.c:506) and the drop-in layout is defined on 32 x 32
 * tiles; this is the documented non-default mode for b = 64 and 128 (32 is accepted too: the same generic kernels,
 * the host-ordered baseline the persistent engines are compared with).  Same four tile kernels with the tile size a
 * parameter, the reference's flat-tree order, host-ordered launches (csrc/qr_tiles.cu).  Output layout = the
 * reference's with 32 -> b; tau compact [(k*p + i)*b + r], p = m / b.  m % b == 0, n % b == 0.  Asynchronous on `stream`. */
typedef struct b200qr_tileplan_s *b200qr_tileplan_t;
int  b200qr_tiles_create(b200qr_tileplan_t *plan, int m, int n, int b);
int  b200qr_tiles_destroy(b200qr_tileplan_t plan);
int  b200qr_tiles_tile(b200qr_tileplan_t plan);
int  b200qr_tiles_factor(b200qr_tileplan_t plan, float *dA, int ldm, void *stream);
const float *b200qr_tiles_tau(b200qr_tileplan_t plan);
/* C (m x ncols, ncols % b == 0) <- Q^T C (transpose != 0) or Q C with the reflectors of a matrix factored by this plan */
int  b200qr_tiles_apply_q(b200qr_tileplan_t plan, const float *dA, int ldm, float *dC, int ldc, int ncols, int transpose, void *stream);
int  b200qr_tiles_last_cuda_error(void);

/* cudaQRTask keeps its plan and device buffer between calls of the same shape; this frees them. */
void b200qr_release_cache(void);

/* ---- trace of the device scheduler (opts.trace = 1) -------------------------------
 * For every tile task (k, i, j) the kernel records a global sequence number when the
 * task starts reading its inputs and when its outputs are published.  Records come
 * back as n x 6 ints: type, l, m, k, start_seq, finish_seq (n = tile task count). */
int b200qr_plan_read_trace(b200qr_plan_t plan, int *out6, long cap);

/* ---- single tile tasks (per-kernel parity tests; doATask, qrdecomp.c:377-449) ------
 * Runs ONE task of the reference's vocabulary on device memory with the reference's
 * pointer arithmetic: mat and tauMat are m x n device matrices with leading dimension
 * ldm (tauMat in the reference tau-matrix layout).  Synchronous.                    */
int b200qr_tile_task(int type, int l, int m, int k, float *dMat, float *dTauMat, int ldm);

/* ---- host mirror of the tile-DAG scheduler (gridscheduler.h:19-21) ----------------
 * Progress-word formulation used by the device scheduler: prog[l][m] = number of
 * completed steps of tile (l,m); a task is ready when its tile's word equals k and the
 * words it depends on have passed k.  Driven serially with the reference's pick rule
 * (highest row, then highest column) it emits gridscheduler.c's exact sequence.    */
typedef struct b200qr_sched_s *b200qr_sched_t;
b200qr_sched_t b200qr_sched_init(int M, int N);                    /* initScheduler  */
int  b200qr_sched_next(b200qr_sched_t s, b200qr_task *out);        /* getNextTask    */
void b200qr_sched_done(b200qr_sched_t s, const b200qr_task *t);    /* doneATask      */
void b200qr_sched_free(b200qr_sched_t s);
int  b200qr_sched_progress(b200qr_sched_t s, int l, int m);        /* prog[l][m]     */
int  b200qr_sched_final_progress(int p, int q, int l, int m);       /* device word of a finished tile column (diagonal: +1) */
int  b200qr_sched_ready_count(b200qr_sched_t s);
int  b200qr_sched_dep_count(int type, int k);   /* initial dependency counter of a task */
/* Predecessors of task t (at most 3); returns how many were written to pred[].      */
int  b200qr_sched_predecessors(const b200qr_task *t, b200qr_task pred[3]);
/* Sweep ticket order of the persistent kernel for an M x N tile grid and G tile columns per
 * sweep: (k, g) pairs sorted by k + G*g.  Returns the number of sweeps (out2 may be NULL).    */
long b200qr_sched_sweep_order(int M, int N, int G, int *out2, long cap);
/* Same with an explicit skew B (key k + B*g; B = G above): any B >= 0 is a topological order.       */
long b200qr_sched_sweep_order_skew(int M, int N, int G, int B, int *out2, long cap);
/* Lagged order of the tensor-core engine 3: far sweeps of a step adjacent (key k + LAG), near-diagonal ones in step order. */
long b200qr_sched_sweep_order_lag(int M, int N, int G, int LAG, int *out2, long cap);
long b200qr_sched_rank_tickets_skew(int M, int N, int G, int B, int rank, int world, int *sweeps2, long cap, int *panels,
                                    long capp, long *npanels);
/* Tickets of one rank of the block-cyclic distribution (block g of G tile columns on rank g % world):
 * its sweeps (k, g) in the global order and its panels k, ascending.  Returns the sweep count.  */
long b200qr_sched_rank_tickets(int M, int N, int G, int rank, int world, int *sweeps2, long cap, int *panels,
                               long capp, long *npanels);
/* One serial order consistent with both ticket streams of the device scheduler, expanded to
 * tile tasks (n x 4 ints: type,l,m,k).                                               */
long b200qr_sched_device_order(int M, int N, int *out4, long cap);

#ifdef __cplusplus
}
#endif
#endif /* QRB200_H */
