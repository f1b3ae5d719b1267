// qr_persistent.cu -- the persistent-kernel tiled QR and its device-side scheduler.
//
// One cooperative launch factors the whole matrix.  Replaces the reference's
// taskKernel / executeTask / completeATask / ring queue / per-tile spin-locks
// (src/gpucalc.cu:117-351, 691-814, 1397-1535).
//
// Scheduler.  The tile DAG of the reference (gridscheduler.c:112-174) is tracked with one
// 32-bit progress word per tile, prog[j*p + i] = number of completed steps of tile (i,j) --
// the reference's (k, taskStatus) pair collapsed.  Work is handed out as
// MACRO TASKS through a single atomic ticket counter, in an order that is a topological
// order of the DAG:
//     P(k)   "panel":  GEQRT(k,k) then the TSQRT chain (i,k), i = k+1..p-1      (1 warp)
//     S(k,g) "sweep":  for the G tile columns of group g: UNMQR(k,j) then the TSMQR
//                      chain (i,j), i = k+1..p-1, with tile row k resident on chip.
//   tickets: P(0); then per step k:  S(k,g0)  P(k+1)  S(k,g0+1)  S(k,g0+2) ...
// (g0 = group holding tile column k+1, so the next panel's column is updated first:
// look-ahead).  A CTA that holds a ticket waits, tile row by tile row, on the progress
// words of its predecessors (ld.acquire.gpu spin) and publishes its own with st.release.gpu.
// Because every dependency points to an earlier ticket and all CTAs are co-resident
// (cooperative launch), the oldest unfinished ticket can always run: no deadlock, no
// queue, no locks, no host round trips.  Panels of successive steps and the sweeps behind
// them pipeline as a wavefront down the matrix.
#include "qr_device.cuh"
#include "../../include/qrb200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace qrb200 {

constexpr int G = 4;                  // tile columns per sweep task (one warp each)
constexpr int NT = 32 * G;            // threads per CTA

struct Params {
    float *A;
    int ldm, p, q, K;
    float *tauc;          // compact tau: [(k*p + i)*32 + r]
    int *prog;            // progress words: [j*p + i]
    int *ctr;             // [0] ticket counter, [1] trace sequence counter
    const int *step_start;  // K+1 entries: first ticket of step k (ticket 0 is P(0))
    int ntickets;
    int *trace;           // optional: 2 ints (start, finish sequence) per tile task
    const long long *trace_off;  // K entries: first trace slot of step k
    unsigned long long watchdog_ns;
};

struct __align__(16) Smem {
    uint64_t sfull[2];    // sweep stage barriers (NT arrivals)
    uint64_t pfull[2];    // panel stage barriers (32 arrivals)
    union {
        struct {
            float col[2][NT][COL_STRIDE];   // staged tile columns, double buffered
            float top[TB][NT];              // tile row k of this column group (resident)
            float v[2][TB * TB];            // reflector block of the current tile row
            float tau[2][TB];
        } s;
        struct {
            float col[2][TB][COL_STRIDE];
            float Rs[TB * 33];
        } p;
    } u;
};

__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Spin until *p >= target.  A scheduler bug must not hang the GPU: after watchdog_ns of
// kernel time the kernel traps (host sees a launch failure -> B200QR_ESTALL).
__device__ __forceinline__ void wait_ge(const int *p, int target, unsigned long long t0, unsigned long long limit)
{
    if (ld_acquire(p) >= target) return;
    unsigned spins = 0;
    while (ld_acquire(p) < target) {
        __nanosleep(64);
        if (((++spins) & 0xfffu) == 0 && gtimer() - t0 > limit) __trap();
    }
}

__device__ __forceinline__ long long trace_slot(const Params &P, int k, int i, int j)
{
    return P.trace_off[k] + (long long)(j - k) * (P.p - k) + (i - k);
}
__device__ __forceinline__ void trace_mark(const Params &P, int k, int i, int j, int which)
{
    if (P.trace) P.trace[2 * trace_slot(P, k, i, j) + which] = atomicAdd(&P.ctr[1], 1);
}

// ------------------------------------------------------------------ P(k): panel, warp 0
__device__ void run_panel(const Params &P, int k, Smem &sm, uint32_t (&ph)[2], unsigned long long t0)
{
    const int lane = threadIdx.x & 31;
    const int p = P.p;
    float *colbase = P.A + (size_t)(k * TB + lane) * P.ldm;   // my column of tile column k
    int *prog_col = P.prog + (size_t)k * p;
    float *tauc = P.tauc + (size_t)k * p * TB;

    auto issue = [&](int row, int s) {
        fence_proxy_async();
        mbar_arrive_expect_tx(&sm.pfull[s], 128);
        bulk_g2s(&sm.u.p.col[s][lane][0], colbase + row * TB, 128, &sm.pfull[s]);
    };

    // ---- GEQRT on the diagonal tile
    wait_ge(prog_col + k, k, t0, P.watchdog_ns);
    issue(k, 0);
    mbar_wait(&sm.pfull[0], ph[0]); ph[0] ^= 1;
    float a[TB];
    load_col_smem(a, sm.u.p.col[0][lane]);
    if (lane == 0) trace_mark(P, k, k, k, 0);
    float t = geqrt_warp(a, lane);
    store_col_gmem(colbase + k * TB, a);      // reflectors final; R part rewritten after the chain
    tauc[k * TB + lane] = t;
#pragma unroll
    for (int r = 0; r < TB; ++r) sm.u.p.Rs[r * 33 + lane] = a[r];
    __threadfence();
    __syncwarp();
    if (lane == 0) { trace_mark(P, k, k, k, 1); st_release(prog_col + k, k + 1); }

    // ---- TSQRT chain down the tile column; R stays in shared memory
    bool have_cur = false;
    for (int i = k + 1; i < p; ++i) {
        const int s = (i - k) & 1;
        if (!have_cur) { wait_ge(prog_col + i, k, t0, P.watchdog_ns); issue(i, s); }
        bool have_next = false;
        if (i + 1 < p && ld_acquire(prog_col + i + 1) >= k) { issue(i + 1, s ^ 1); have_next = true; }
        mbar_wait(&sm.pfull[s], ph[s]); ph[s] ^= 1;
        float b[TB];
        load_col_smem(b, sm.u.p.col[s][lane]);
        if (lane == 0) trace_mark(P, k, i, k, 0);
        float ti = tsqrt_warp(b, sm.u.p.Rs, lane);
        store_col_gmem(colbase + i * TB, b);
        tauc[i * TB + lane] = ti;
        __threadfence();
        __syncwarp();
        if (lane == 0) { trace_mark(P, k, i, k, 1); st_release(prog_col + i, k + 1); }
        have_cur = have_next;
    }
    // ---- final R of the diagonal tile (upper triangle only; reflectors below stay)
#pragma unroll
    for (int r = 0; r < TB; ++r)
        if (r <= lane) colbase[k * TB + r] = sm.u.p.Rs[r * 33 + lane];
}

// ------------------------------------------------------------------ S(k,g): column sweep
__device__ void run_sweep(const Params &P, int k, int g, Smem &sm, uint32_t (&ph)[2], unsigned long long t0)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = P.p;
    const int jt = g * G + warp;                        // my warp's tile column
    const bool active = jt > k && jt < P.q;
    const int nrows = p - k;
    float *colbase = P.A + (size_t)(active ? jt * TB + lane : 0) * P.ldm;
    const float *vbase = P.A + (size_t)(k * TB + lane) * P.ldm;   // reflector tiles live in tile column k
    const float *taubase = P.tauc + (size_t)k * p * TB;
    const int *prog_panel = P.prog + (size_t)k * p;
    int *prog_mine = P.prog + (size_t)(active ? jt : 0) * p;

    auto ready = [&](int row) -> bool {
        if (ld_acquire(prog_panel + row) < k + 1) return false;          // reflectors of (row,k) published
        if (active && ld_acquire(prog_mine + row) < k) return false;     // my tile finished step k-1
        return true;
    };
    auto issue = [&](int row, int s) {
        uint32_t bytes = (active ? 128u : 0u) + (warp == 0 ? 128u : 0u) + (tid == 0 ? 128u : 0u);
        fence_proxy_async();
        if (bytes) mbar_arrive_expect_tx(&sm.sfull[s], bytes);
        else mbar_arrive(&sm.sfull[s]);
        if (active) bulk_g2s(&sm.u.s.col[s][tid][0], colbase + row * TB, 128, &sm.sfull[s]);
        if (warp == 0) {
            bulk_g2s(&sm.u.s.v[s][lane * TB], vbase + row * TB, 128, &sm.sfull[s]);
            if (lane == 0) bulk_g2s(sm.u.s.tau[s], taubase + row * TB, 128, &sm.sfull[s]);
        }
    };

    bool have_cur = false;
    for (int idx = 0; idx < nrows; ++idx) {
        const int row = k + idx, s = idx & 1;
        __syncthreads();                                   // everyone is done with stage s^1's reflectors
        if (!have_cur) {
            unsigned spins = 0;
            while (!ready(row)) {
                __nanosleep(64);
                if (((++spins) & 0xfffu) == 0 && gtimer() - t0 > P.watchdog_ns) __trap();
            }
            issue(row, s);
        }
        bool have_next = false;
        if (idx + 1 < nrows && ready(row + 1)) { issue(row + 1, s ^ 1); have_next = true; }
        mbar_wait(&sm.sfull[s], ph[s]); ph[s] ^= 1;

        if (idx == 0) {
            // tile row k: UNMQR with the GEQRT reflectors of the diagonal tile
            make_unit_lower(sm.u.s.v[s], tid, NT);
            __syncthreads();
            if (active) {
                float b[TB];
                load_col_smem(b, sm.u.s.col[s][tid]);
                if (lane == 0) trace_mark(P, k, k, jt, 0);
                apply_block_reflectors<false>(b, nullptr, 0, sm.u.s.v[s], sm.u.s.tau[s]);
#pragma unroll
                for (int r = 0; r < TB; ++r) sm.u.s.top[r][tid] = b[r];
            }
        } else if (active) {
            // tile row `row`: TSMQR against the resident tile row k
            float b[TB];
            load_col_smem(b, sm.u.s.col[s][tid]);
            if (lane == 0) trace_mark(P, k, row, jt, 0);
            apply_block_reflectors<true>(b, &sm.u.s.top[0][tid], NT, sm.u.s.v[s], sm.u.s.tau[s]);
            store_col_smem(sm.u.s.col[s][tid], b);
            fence_proxy_async_smem();
            bulk_s2g(colbase + row * TB, sm.u.s.col[s][tid], 128);
            bulk_commit();
            bulk_wait_all();
            __threadfence();
            __syncwarp();
            if (lane == 0) { trace_mark(P, k, row, jt, 1); st_release(prog_mine + row, k + 1); }
        }
        have_cur = have_next;
    }
    // tile row k is final for this step: write it back and publish
    if (active) {
        float *stg = sm.u.s.col[nrows & 1][tid];
        float b[TB];
#pragma unroll
        for (int r = 0; r < TB; ++r) b[r] = sm.u.s.top[r][tid];
        store_col_smem(stg, b);
        fence_proxy_async_smem();
        bulk_s2g(colbase + k * TB, stg, 128);
        bulk_commit();
        bulk_wait_all();
        __threadfence();
        __syncwarp();
        if (lane == 0) { trace_mark(P, k, k, jt, 1); st_release(prog_mine + k, k + 1); }
    }
}

// ticket -> macro task.  Keep in lock-step with b200qr_sched_device_order() (sched_host.cpp).
__device__ __forceinline__ void decode_ticket(const Params &P, int t, int &is_panel, int &k, int &g)
{
    if (t == 0) { is_panel = 1; k = 0; g = 0; return; }
    int lo = 0, hi = P.K - 1;                    // largest step with step_start[step] <= t
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (P.step_start[mid] <= t) lo = mid; else hi = mid - 1;
    }
    int r = t - P.step_start[lo];
    int next_panel = (lo + 1 < P.K) ? 1 : 0;
    int g0 = (lo + 1) / G;
    if (r == 0) { is_panel = 0; k = lo; g = g0; }
    else if (r == 1 && next_panel) { is_panel = 1; k = lo + 1; g = 0; }
    else { is_panel = 0; k = lo; g = g0 + r - next_panel; }
}

__global__ void __launch_bounds__(NT, 3) tiledqr_persistent(const Params P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    __shared__ int s_task[3];
    const unsigned long long t0 = gtimer();

    if (threadIdx.x == 0) {
        mbar_init(&sm.sfull[0], NT); mbar_init(&sm.sfull[1], NT);
        mbar_init(&sm.pfull[0], 32); mbar_init(&sm.pfull[1], 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t ph_s[2] = {0, 0}, ph_p[2] = {0, 0};

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = atomicAdd(&P.ctr[0], 1);
            int is_panel = -1, k = 0, g = 0;
            if (t < P.ntickets) decode_ticket(P, t, is_panel, k, g);
            s_task[0] = is_panel; s_task[1] = k; s_task[2] = g;
        }
        __syncthreads();
        const int is_panel = s_task[0], k = s_task[1], g = s_task[2];
        if (is_panel < 0) break;
        if (is_panel) {
            if (threadIdx.x < 32) run_panel(P, k, sm, ph_p, t0);
        } else {
            run_sweep(P, k, g, sm, ph_s, t0);
        }
    }
}

// ------------------------------------------------------------------ single tile tasks
// doATask (qrdecomp.c:377-449) on device memory, one warp; used by the per-kernel parity
// tests, doCUDADAPP and cudaQRFull.  taumat is in the reference's tau-matrix layout.
__device__ void tile_task_device(int type, int l, int m, int k, float *mat, float *taumat, int ldm,
                                 float *Rs, float *vs, float *ts, float *top)
{
    const int lane = threadIdx.x & 31;
    auto tile = [&](float *base, int I, int J) { return base + (size_t)(J * TB + lane) * ldm + I * TB; };
    auto tauv = [&](int I, int J) { return taumat + (size_t)(J * TB) * ldm + I * TB; };
    float x[TB];
    switch (type) {
    case B200QR_QRS: {
        load_col_gmem(x, tile(mat, k, k));
        float t = geqrt_warp(x, lane);
        store_col_gmem(tile(mat, k, k), x);
        tauv(k, k)[lane] = t;
        break;
    }
    case B200QR_QRD: {
        load_col_gmem(x, tile(mat, k, k));
#pragma unroll
        for (int r = 0; r < TB; ++r) Rs[r * 33 + lane] = x[r];
        __syncwarp();
        load_col_gmem(x, tile(mat, l, k));
        float t = tsqrt_warp(x, Rs, lane);
        store_col_gmem(tile(mat, l, k), x);
        float *rcol = tile(mat, k, k);
#pragma unroll
        for (int r = 0; r < TB; ++r)
            if (r <= lane) rcol[r] = Rs[r * 33 + lane];
        tauv(l, k)[lane] = t;
        break;
    }
    case B200QR_SAPP: {
        load_col_gmem(x, tile(mat, k, k));
        store_col_smem(vs + lane * TB, x);
        ts[lane] = tauv(k, k)[lane];
        __syncwarp();
        make_unit_lower(vs, lane, 32);
        __syncwarp();
        load_col_gmem(x, tile(mat, k, m));
        apply_block_reflectors<false>(x, nullptr, 0, vs, ts);
        store_col_gmem(tile(mat, k, m), x);
        break;
    }
    case B200QR_DAPP: {
        load_col_gmem(x, tile(mat, l, k));
        store_col_smem(vs + lane * TB, x);
        ts[lane] = tauv(l, k)[lane];
        load_col_gmem(x, tile(mat, k, m));
#pragma unroll
        for (int r = 0; r < TB; ++r) top[r * TB + lane] = x[r];
        __syncwarp();
        load_col_gmem(x, tile(mat, l, m));
        apply_block_reflectors<true>(x, top + lane, TB, vs, ts);
        store_col_gmem(tile(mat, l, m), x);
#pragma unroll
        for (int r = 0; r < TB; ++r) x[r] = top[r * TB + lane];
        store_col_gmem(tile(mat, k, m), x);
        break;
    }
    }
}

__global__ void __launch_bounds__(32) tile_task_kernel(int type, int l, int m, int k, float *mat, float *taumat,
                                                       int ldm)
{
    __shared__ __align__(16) float Rs[TB * 33];
    __shared__ __align__(16) float vs[TB * TB];
    __shared__ __align__(16) float ts[TB];
    __shared__ __align__(16) float top[TB * TB];
    tile_task_device(type, l, m, k, mat, taumat, ldm, Rs, vs, ts, top);
}

// One step of the host-ordered driver (cudaQRFull): the panel of step k as one block,
// then one block per trailing tile column (UNMQR + TSMQR chain).  mode 0 = panel, 1 = update.
__global__ void __launch_bounds__(32) step_kernel(int mode, int k, int p, float *mat, float *taumat, int ldm)
{
    __shared__ __align__(16) float Rs[TB * 33];
    __shared__ __align__(16) float vs[TB * TB];
    __shared__ __align__(16) float ts[TB];
    __shared__ __align__(16) float top[TB * TB];
    if (mode == 0) {
        tile_task_device(B200QR_QRS, k, k, k, mat, taumat, ldm, Rs, vs, ts, top);
        for (int i = k + 1; i < p; ++i) {
            __syncwarp();
            tile_task_device(B200QR_QRD, i, k, k, mat, taumat, ldm, Rs, vs, ts, top);
        }
    } else {
        int j = k + 1 + blockIdx.x;
        tile_task_device(B200QR_SAPP, k, j, k, mat, taumat, ldm, Rs, vs, ts, top);
        for (int i = k + 1; i < p; ++i) {
            __syncwarp();
            tile_task_device(B200QR_DAPP, i, j, k, mat, taumat, ldm, Rs, vs, ts, top);
        }
    }
}

// TSMQR micro-benchmark kernel (testDAPP / doCUDADAPP): block x updates the column pair
// (A, B) = (pairs + x*pair_stride, +32 rows) with reflectors V2 and tau.
__global__ void __launch_bounds__(32) dapp_pairs_kernel(const float *V2, const float *tau, float *pairs,
                                                        size_t pair_stride, int ldm)
{
    __shared__ __align__(16) float vs[TB * TB];
    __shared__ __align__(16) float ts[TB];
    __shared__ __align__(16) float top[TB * TB];
    const int lane = threadIdx.x;
    float x[TB];
    load_col_gmem(x, V2 + (size_t)lane * ldm);
    store_col_smem(vs + lane * TB, x);
    ts[lane] = tau[lane];
    float *A = pairs + blockIdx.x * pair_stride + (size_t)lane * ldm;
    float *B = A + TB;
    load_col_gmem(x, A);
#pragma unroll
    for (int r = 0; r < TB; ++r) top[r * TB + lane] = x[r];
    __syncwarp();
    load_col_gmem(x, B);
    apply_block_reflectors<true>(x, top + lane, TB, vs, ts);
    store_col_gmem(B, x);
#pragma unroll
    for (int r = 0; r < TB; ++r) x[r] = top[r * TB + lane];
    store_col_gmem(A, x);
}

// compact tau -> reference tau-matrix layout (qrdecomp.c:486,686: rows 0..31 of the first
// column of tile (i,k)); everything else zero.
__global__ void tau_expand_kernel(const float *tauc, float *taumat, int ldtau, int p, int K)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;    // over K*p*32
    if (idx >= K * p * TB) return;
    int r = idx & 31, i = (idx >> 5) % p, k = (idx >> 5) / p;
    if (i >= k) taumat[(size_t)(k * TB) * ldtau + i * TB + r] = tauc[idx];
}

}  // namespace qrb200

// ====================================================================== host side / C ABI
using namespace qrb200;

static thread_local int g_last_cuda = 0;
#define CK(call)                                                    \
    do {                                                            \
        cudaError_t e_ = (call);                                    \
        if (e_ != cudaSuccess) { g_last_cuda = (int)e_; return B200QR_ECUDA; } \
    } while (0)

struct b200qr_plan_s {
    int m, n, p, q, K;
    int grid;
    int ntickets;
    long tile_tasks;
    int trace;
    int *d_prog = nullptr, *d_ctr = nullptr, *d_step_start = nullptr, *d_trace = nullptr;
    long long *d_trace_off = nullptr;
    float *d_tauc = nullptr;
    std::vector<long long> trace_off;
    unsigned long long watchdog_ns;
};

static size_t smem_bytes() { return sizeof(Smem); }

extern "C" {

int b200qr_version(void) { return 100; }

const char *b200qr_strerror(int code)
{
    switch (code) {
    case B200QR_OK: return "ok";
    case B200QR_EINVAL: return "invalid argument";
    case B200QR_ECUDA: return "CUDA runtime error";
    case B200QR_ENOMEM: return "out of memory";
    case B200QR_ESTALL: return "persistent scheduler stalled (watchdog)";
    }
    return "unknown";
}

int b200qr_last_cuda_error(void) { return g_last_cuda; }

int b200qr_device_sm_count(void)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return sms;
}

int b200qr_plan_create(b200qr_plan_t *out, int m, int n, const b200qr_opts *opts)
{
    if (!out || m <= 0 || n <= 0 || m % TB || n % TB) return B200QR_EINVAL;
    b200qr_plan_s *pl = new b200qr_plan_s;
    pl->m = m; pl->n = n; pl->p = m / TB; pl->q = n / TB;
    pl->K = pl->p < pl->q ? pl->p : pl->q;
    pl->trace = opts ? opts->trace : 0;
    const char *wd = getenv("QRB200_WATCHDOG_MS");
    pl->watchdog_ns = (unsigned long long)(wd ? atol(wd) : 60000) * 1000000ull;

    // step table of the ticket order (see decode_ticket)
    const int p = pl->p, q = pl->q, K = pl->K;
    std::vector<int> start(K + 1);
    pl->trace_off.resize(K);
    int t = 1;
    long tiles = 0;
    const int GB = (q + G - 1) / G;
    for (int k = 0; k < K; ++k) {
        start[k] = t;
        if (k + 1 < q) t += GB - (k + 1) / G;
        if (k + 1 < K) t += 1;
        pl->trace_off[k] = tiles;
        tiles += (long)(p - k) * (q - k);
    }
    start[K] = t;
    pl->ntickets = t;
    pl->tile_tasks = tiles;

    int dev = 0, sms = 0, occ = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaFuncSetAttribute(tiledqr_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tiledqr_persistent, NT, smem_bytes()));
    if (occ < 1) { delete pl; return B200QR_EINVAL; }
    int grid = occ * sms;
    if (opts && opts->grid_ctas > 0 && opts->grid_ctas < grid) grid = opts->grid_ctas;
    if (grid > pl->ntickets) grid = pl->ntickets;
    pl->grid = grid;

    CK(cudaMalloc(&pl->d_prog, sizeof(int) * (size_t)p * q));
    CK(cudaMalloc(&pl->d_ctr, sizeof(int) * 64));
    CK(cudaMalloc(&pl->d_step_start, sizeof(int) * (K + 1)));
    CK(cudaMalloc(&pl->d_tauc, sizeof(float) * (size_t)K * p * TB));
    CK(cudaMemcpy(pl->d_step_start, start.data(), sizeof(int) * (K + 1), cudaMemcpyHostToDevice));
    CK(cudaMemset(pl->d_tauc, 0, sizeof(float) * (size_t)K * p * TB));
    if (pl->trace) {
        CK(cudaMalloc(&pl->d_trace, sizeof(int) * 2 * (size_t)tiles));
        CK(cudaMalloc(&pl->d_trace_off, sizeof(long long) * K));
        CK(cudaMemcpy(pl->d_trace_off, pl->trace_off.data(), sizeof(long long) * K, cudaMemcpyHostToDevice));
    }
    *out = pl;
    return B200QR_OK;
}

int b200qr_plan_destroy(b200qr_plan_t pl)
{
    if (!pl) return B200QR_EINVAL;
    cudaFree(pl->d_prog); cudaFree(pl->d_ctr); cudaFree(pl->d_step_start); cudaFree(pl->d_tauc);
    cudaFree(pl->d_trace); cudaFree(pl->d_trace_off);
    delete pl;
    return B200QR_OK;
}

int b200qr_plan_grid(b200qr_plan_t pl) { return pl ? pl->grid : B200QR_EINVAL; }
long b200qr_plan_tile_tasks(b200qr_plan_t pl) { return pl ? pl->tile_tasks : B200QR_EINVAL; }
long b200qr_plan_macro_tasks(b200qr_plan_t pl) { return pl ? pl->ntickets : B200QR_EINVAL; }
const float *b200qr_plan_tau_device(b200qr_plan_t pl) { return pl ? pl->d_tauc : nullptr; }

int b200qr_factor_device(b200qr_plan_t pl, float *dA, int ldm, void *stream_)
{
    if (!pl || !dA || ldm < pl->m || (ldm & 3) || ((uintptr_t)dA & 15)) return B200QR_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(cudaMemsetAsync(pl->d_prog, 0, sizeof(int) * (size_t)pl->p * pl->q, stream));
    CK(cudaMemsetAsync(pl->d_ctr, 0, sizeof(int) * 64, stream));
    if (pl->trace) CK(cudaMemsetAsync(pl->d_trace, 0xff, sizeof(int) * 2 * (size_t)pl->tile_tasks, stream));
    Params P;
    P.A = dA; P.ldm = ldm; P.p = pl->p; P.q = pl->q; P.K = pl->K;
    P.tauc = pl->d_tauc; P.prog = pl->d_prog; P.ctr = pl->d_ctr;
    P.step_start = pl->d_step_start; P.ntickets = pl->ntickets;
    P.trace = pl->trace ? pl->d_trace : nullptr; P.trace_off = pl->d_trace_off;
    P.watchdog_ns = pl->watchdog_ns;
    void *args[] = {&P};
    CK(cudaLaunchCooperativeKernel((void *)tiledqr_persistent, dim3(pl->grid), dim3(NT), args, smem_bytes(), stream));
    return B200QR_OK;
}

int b200qr_tau_to_reference(b200qr_plan_t pl, float *dTauMat, int ldtau, void *stream_)
{
    if (!pl || !dTauMat || ldtau < pl->m) return B200QR_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(cudaMemset2DAsync(dTauMat, sizeof(float) * ldtau, 0, sizeof(float) * pl->m, pl->n, stream));
    int total = pl->K * pl->p * TB;
    tau_expand_kernel<<<(total + 255) / 256, 256, 0, stream>>>(pl->d_tauc, dTauMat, ldtau, pl->p, pl->K);
    CK(cudaGetLastError());
    return B200QR_OK;
}

int b200qr_factor_host(float *hA, int m, int n, int ldm, float *hTauMat, float *kernel_ms)
{
    if (!hA || ldm < m) return B200QR_EINVAL;
    b200qr_plan_t pl = nullptr;
    int rc = b200qr_plan_create(&pl, m, n, nullptr);
    if (rc) return rc;
    float *dA = nullptr, *dT = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() {
        cudaFree(dA); cudaFree(dT);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        b200qr_plan_destroy(pl);
    };
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_last_cuda = (int)e_; cleanup(); return B200QR_ECUDA; } } while (0)
    CKC(cudaMalloc(&dA, sizeof(float) * (size_t)m * n));
    // one strided copy instead of the reference's n column copies (gpucalc.cu:1614-1619)
    CKC(cudaMemcpy2D(dA, sizeof(float) * m, hA, sizeof(float) * ldm, sizeof(float) * m, n, cudaMemcpyHostToDevice));
    CKC(cudaEventCreate(&e0)); CKC(cudaEventCreate(&e1));
    CKC(cudaEventRecord(e0, 0));
    rc = b200qr_factor_device(pl, dA, m, nullptr);
    if (rc) { cleanup(); return rc; }
    CKC(cudaEventRecord(e1, 0));
    cudaError_t se = cudaEventSynchronize(e1);
    if (se != cudaSuccess) { g_last_cuda = (int)se; cleanup(); return se == cudaErrorLaunchFailure ? B200QR_ESTALL : B200QR_ECUDA; }
    if (kernel_ms) CKC(cudaEventElapsedTime(kernel_ms, e0, e1));
    CKC(cudaMemcpy2D(hA, sizeof(float) * ldm, dA, sizeof(float) * m, sizeof(float) * m, n, cudaMemcpyDeviceToHost));
    if (hTauMat) {
        CKC(cudaMalloc(&dT, sizeof(float) * (size_t)m * n));
        rc = b200qr_tau_to_reference(pl, dT, m, nullptr);
        if (rc) { cleanup(); return rc; }
        CKC(cudaMemcpy2D(hTauMat, sizeof(float) * ldm, dT, sizeof(float) * m, sizeof(float) * m, n, cudaMemcpyDeviceToHost));
    }
    cleanup();
    return B200QR_OK;
#undef CKC
}

int b200qr_plan_read_trace(b200qr_plan_t pl, int *out6, long cap)
{
    if (!pl || !pl->trace || !out6) return B200QR_EINVAL;
    std::vector<int> raw(2 * (size_t)pl->tile_tasks);
    CK(cudaMemcpy(raw.data(), pl->d_trace, sizeof(int) * raw.size(), cudaMemcpyDeviceToHost));
    long n = 0;
    for (int k = 0; k < pl->K; ++k)
        for (int j = k; j < pl->q; ++j)
            for (int i = k; i < pl->p; ++i) {
                long long slot = pl->trace_off[k] + (long long)(j - k) * (pl->p - k) + (i - k);
                if (n < cap) {
                    int type = (i == k) ? (j == k ? B200QR_QRS : B200QR_SAPP) : (j == k ? B200QR_QRD : B200QR_DAPP);
                    int *o = out6 + 6 * n;
                    o[0] = type; o[1] = i; o[2] = j; o[3] = k; o[4] = raw[2 * slot]; o[5] = raw[2 * slot + 1];
                }
                ++n;
            }
    return B200QR_OK;
}

int b200qr_tile_task(int type, int l, int m, int k, float *dMat, float *dTauMat, int ldm)
{
    if (!dMat || !dTauMat || type < 0 || type > 3 || (ldm & 3)) return B200QR_EINVAL;
    tile_task_kernel<<<1, 32>>>(type, l, m, k, dMat, dTauMat, ldm);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return B200QR_OK;
}

// ---------------------------------------------------------------- legacy entry points
// (reference include/gpucalc.h:3-6)

static bool env_on(const char *name)
{
    const char *v = getenv(name);
    return v && *v && strcmp(v, "0") != 0;
}

void cudaQRTask(float *mat, int m, int n, int ldm, int maxblocks)
{
    // reference: gpucalc.cu:1579-1686.  Same observable behaviour: in-place result in the
    // caller's host matrix, the two progress prints; errors are printed, not returned.
    const bool quiet = env_on("QRB200_QUIET");
    const int p = m / TB, q = n / TB;
    m = p * TB; n = q * TB;                       // the reference truncates silently (gpucalc.cu:1582)
    if (p <= 0 || q <= 0 || !mat) { printf("error: bad matrix shape\n"); return; }
    long total = 0;
    for (int k = 0; k < (p < q ? p : q); ++k) total += (long)(p - k) * (q - k);
    if (!quiet) printf("(%d,%d)\n%ld tasks\n", m, n, total);   // calcTotalTasks, gpucalc.cu:1557

    b200qr_opts opts;
    memset(&opts, 0, sizeof opts);
    if (env_on("QRB200_HONOR_MAXBLOCKS") && maxblocks > 0) opts.grid_ctas = maxblocks;
    b200qr_plan_t pl = nullptr;
    float *dA = nullptr;
    cudaEvent_t e0, e1;
    int rc = b200qr_plan_create(&pl, m, n, &opts);
    if (rc) { printf("error creating plan: %s\n", b200qr_strerror(rc)); return; }
    if (cudaMalloc(&dA, sizeof(float) * (size_t)m * n) != cudaSuccess) {
        printf("error allocating dev mat\n");
        b200qr_plan_destroy(pl);
        return;
    }
    if (cudaMemcpy2D(dA, sizeof(float) * m, mat, sizeof(float) * ldm, sizeof(float) * m, n,
                     cudaMemcpyHostToDevice) != cudaSuccess)
        printf("error cpying dev mat\n");
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    rc = b200qr_factor_device(pl, dA, m, nullptr);
    cudaEventRecord(e1, 0);
    cudaError_t se = cudaEventSynchronize(e1);
    if (rc || se != cudaSuccess) printf("error running QR kernel: %s (%d)\n", b200qr_strerror(rc), (int)se);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (!quiet) printf("GPU: %5.3f ms\n", ms);                 // gpucalc.cu:1647
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (cudaMemcpy2D(mat, sizeof(float) * ldm, dA, sizeof(float) * m, sizeof(float) * m, n,
                     cudaMemcpyDeviceToHost) != cudaSuccess)
        printf("copy back failed\n");
    cudaFree(dA);
    b200qr_plan_destroy(pl);
    if (env_on("QRB200_COMPAT_RESET")) cudaDeviceReset();       // gpucalc.cu:1685
}

void cudaQRFull(float *mat, int m, int n)
{
    // Declared at gpucalc.h:4, never defined by the reference (gpucalc.cu:1801-1877 is a
    // comment).  Host-ordered driver: per step one panel launch and one update launch.
    const int p = m / TB, q = n / TB, ldm = m;
    if (p <= 0 || q <= 0 || !mat) { printf("error: bad matrix shape\n"); return; }
    float *dA = nullptr, *dT = nullptr;
    size_t bytes = sizeof(float) * (size_t)ldm * (q * TB);
    if (cudaMalloc(&dA, bytes) != cudaSuccess || cudaMalloc(&dT, bytes) != cudaSuccess) {
        printf("error allocating dev mat\n");
        cudaFree(dA);
        return;
    }
    cudaMemcpy(dA, mat, bytes, cudaMemcpyHostToDevice);
    cudaMemset(dT, 0, bytes);
    const int K = p < q ? p : q;
    for (int k = 0; k < K; ++k) {
        step_kernel<<<1, 32>>>(0, k, p, dA, dT, ldm);
        if (q - k - 1 > 0) step_kernel<<<q - k - 1, 32>>>(1, k, p, dA, dT, ldm);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("error %d\n", (int)e);
    cudaMemcpy(mat, dA, bytes, cudaMemcpyDeviceToHost);
    cudaFree(dA); cudaFree(dT);
}

void testDAPP(float *timings, int n, int nblocks)
{
    // reference: gpucalc.cu:1706-1774 (same data recipe: srand(5), values on a 101-point grid)
    if (!timings || n <= 0 || nblocks <= 0) return;
    const size_t chunk = 64 * 32;
    for (int t = 0; t < n; ++t) {
        srand(5);
        std::vector<float> h((1 + (size_t)nblocks) * chunk), htau(32);
        for (size_t i = 0; i < chunk; ++i) h[i] = ((float)(rand() % 101) - 50.0) / 50.0;
        for (int i = 0; i < 32; ++i) htau[i] = ((float)(rand() % 101) - 50.0) / 50.0;
        for (int b = 1; b <= nblocks; ++b) memcpy(&h[b * chunk], &h[0], chunk * sizeof(float));
        float *d = nullptr, *dt = nullptr;
        cudaMalloc(&d, h.size() * sizeof(float));
        cudaMalloc(&dt, 32 * sizeof(float));
        cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
        cudaMemcpy(dt, htau.data(), 32 * sizeof(float), cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, 0);
        dapp_pairs_kernel<<<nblocks, 32>>>(d + 32, dt, d + chunk, chunk, 64);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        timings[t] = ms * nblocks;
        cudaError_t e = cudaGetLastError();
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        cudaFree(d); cudaFree(dt);
        if (e != cudaSuccess) { printf("ERROR!!! %d\n", (int)e); break; }
    }
}

void doCUDADAPP(float *mat)
{
    // reference: gpucalc.cu:1776-1799
    if (!mat) return;
    float *d = nullptr;
    cudaMalloc(&d, 64 * 64 * sizeof(float));
    cudaMemcpy(d, mat, 64 * 64 * sizeof(float), cudaMemcpyHostToDevice);
    dapp_pairs_kernel<<<1, 32>>>(d + 32, d, d + 32 * 64, 0, 64);
    cudaMemcpy(mat, d, 64 * 64 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error %d\n", (int)e);
    cudaFree(d);
}

}  // extern "C"
