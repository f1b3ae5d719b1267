// qr_batched.cu -- many small tiled QR factorisations at once, and the TSQR tree built on them.
//
// BASELINE.json configs[2] (10 000 independent 256x256 matrices) and configs[3] (tall-skinny
// 4 194 304 x 256, TSQR/CAQR reduction tree).  Each problem is the reference's tile algorithm
// unchanged (same four tile kernels, same flat tree inside a problem, same output layout: R above,
// GEQRT / TSQRT reflectors below, tau per reflector), so a 256x256 problem of the batch equals
// cudaQRTask on that matrix.  What changes is the schedule: a problem has at most 8 tile columns,
// so ONE CTA factors it with a static assignment -- warp w owns tile column w: it applies the
// panels of the steps k < w to its column (UNMQR + TSMQR chain, the step's tile row k resident
// in shared memory), then factors its own column (GEQRT + TSQRT chain) and retires.  Warps follow
// each other through per-step progress counters in shared memory (the tile-DAG dependencies of
// gridscheduler.c:112-174 collapse to "panel k has finished tile row i").  CTAs pull problems from
// an atomic counter; three CTAs share an SM, so one problem's idle warps are another's busy ones.
//
// TSQR: the rows are cut into leaves (leaf_rows x n, a batch), every leaf is factored as above, the
// leaves' R factors are stacked and factored again, `fanin` at a time, until one R is left.  The
// implicit Q is the tree of reflector sets (leaves in place in A, inner nodes in the plan's
// workspace) -- a different Householder layout from the reference's flat chain, as SURVEY 8e
// notes; parity for this path is on R (up to row signs) and on R^T R = A^T A.
#include "qr_device.cuh"
#include "../../include/qrb200.h"

#include <cstdlib>
#include <cstring>
#include <vector>

namespace qrb200 {

constexpr int BQ = 8;                       // max tile columns per problem = warps per CTA
constexpr int WARP_FLOATS = 1088 + 1024 + 32 + 64;   // V/Rs, top row, tau, broadcast slots

struct BatchArgs {
    float *A;
    long long strideA;      // elements between consecutive problems
    int ldm, p, q, nprob;
    float *tauc;            // nprob x K*p*32, compact per problem: [(k*p + i)*32 + r]
    int *counter;
};

__device__ __forceinline__ void spin_ge(volatile int *flag, int target)
{
    while (*flag < target) __nanosleep(20);
}

__global__ void __launch_bounds__(BQ * 32, 2) batched_tiledqr_kernel(const BatchArgs a)
{
    extern __shared__ __align__(16) float bsm[];
    __shared__ int s_prob;
    __shared__ volatile int panel_prog[BQ];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *Vs = bsm + w * WARP_FLOATS;      // reflector tile of the current row; Rs while this warp is the panel
    float *top = Vs + 1088;                 // tile row k of my column: top[c*32 + lane]
    float *ts = top + 1024;
    float *xs = ts + 32;
    const int p = a.p, q = a.q, K = p < q ? p : q, ldm = a.ldm;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_prob = atomicAdd(a.counter, 1);
        if (threadIdx.x < BQ) panel_prog[threadIdx.x] = 0;
        __syncthreads();
        const int b = s_prob;
        if (b >= a.nprob) break;
        if (w >= q) continue;
        float *M = a.A + (size_t)b * a.strideA;
        float *tb = a.tauc + (size_t)b * K * p * TB;
        auto tile = [&](int I, int J) { return M + (size_t)(J * TB + lane) * ldm + I * TB; };

        for (int k = 0; k < K; ++k) {
            float x[TB];
            if (k == w) {
                // ---- my panel: GEQRT + TSQRT chain (qrdecomp.c:532-557, 689-721)
                load_col_gmem(x, tile(k, k));
                float t = geqrt_warp(x, xs, lane);
                store_col_gmem(tile(k, k), x);
                tb[((size_t)k * p + k) * TB + lane] = t;
#pragma unroll
                for (int r = 0; r < TB; ++r) Vs[r * 33 + lane] = x[r];
                __threadfence_block();
                __syncwarp();
                if (lane == 0) panel_prog[k] = 1;
                for (int i = k + 1; i < p; ++i) {
                    load_col_gmem(x, tile(i, k));
                    float ti = tsqrt_warp(x, Vs, xs, top, lane);
                    store_col_gmem(tile(i, k), x);
                    tb[((size_t)k * p + i) * TB + lane] = ti;
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) panel_prog[k] = i - k + 1;
                }
                float *rcol = tile(k, k);
#pragma unroll
                for (int r = 0; r < TB; ++r)
                    if (r <= lane) rcol[r] = Vs[r * 33 + lane];
                break;                                   // columns <= k have no work after step k
            }
            // ---- step k applied to my column: UNMQR + TSMQR chain (qrdecomp.c:620-644, 736-762)
            for (int i = k; i < p; ++i) {
                spin_ge(&panel_prog[k], i - k + 1);
                load_col_gmem(x, tile(i, k));            // reflector tile (L2; written by warp k)
                store_col_smem(Vs + lane * TB, x);
                ts[lane] = __ldcg(tb + ((size_t)k * p + i) * TB + lane);
                __syncwarp();
                load_col_gmem(x, tile(i, w));
                if (i == k) {
                    apply_reflectors_nc<1, false>(x, x, nullptr, nullptr, 0, Vs, ts);
#pragma unroll
                    for (int r = 0; r < TB; ++r) top[r * TB + lane] = x[r];
                } else {
                    apply_reflectors_nc<1, true>(x, x, top + lane, top + lane, TB, Vs, ts);
                    store_col_gmem(tile(i, w), x);
                }
                __syncwarp();
            }
#pragma unroll
            for (int r = 0; r < TB; ++r) x[r] = top[r * TB + lane];
            store_col_gmem(tile(k, w), x);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Round 2: ONE WARP PER PROBLEM, phases in lock-step across the CTA.
// What round 1's kernel above measured (21.2 ms for 10 000 x 256^2; profiles/r1_measurements.md): its warps alternate
// between FFMA-bound update chains and latency-bound panel chains at different times, and this SM's schedulers let an
// FFMA-bound warp starve a panel chain that shares its sub-partition (tools/probe_panel.py: a TSQRT takes 162k cycles
// instead of 13k next to eight FFMA warps) -- while the panel chain is the critical path of every problem.
// Here a CTA of BW warps works on BW problems at a time, one per warp, and all warps are in the SAME phase of their
// problem: step k's panel (GEQRT + TSQRT chain; sixteen latency-bound chains share the SM, 46 tiles per 100k cycles --
// the best panel throughput an SM has), a CTA barrier, step k's update of the trailing tile columns, two columns per
// pass (every warp FFMA-bound at once), a barrier.  No warp ever waits for another one's data: the panel tiles, their
// tau and the trailing tiles of a problem are written and read back by the same threads.  Same four tile kernels, same
// operation order per tile and same output layout as cudaQRTask on each problem (bit-identical: bench.py checks it).
constexpr int BW_FLOATS = 1088 + 1024 + 1024 + 1024 + 64 + 64;   // Rs / V tile A, V tile B, top row of column 0 (V2 out while a panel), of column 1, tau x 2, slots

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// this lane's 128-byte column of a tile, global -> shared, asynchronously (LDGSTS); wait with cp_async_wait_all
__device__ __forceinline__ void cp_async_col(float *dst_smem, const float *src_gmem)
{
    const uint32_t d = smem_u32(dst_smem);
#pragma unroll
    for (int i = 0; i < 8; ++i)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16 * i), "l"(src_gmem + 4 * i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Every warp of the CTA does the same work at the same time, so a warp that waits for a tile waits together with all
// the others: nothing hides a synchronous load (first version, 12 warps, loads in front of every tile: 12.8 ms for
// 10 000 x 256^2, issue slots 44 % busy, 0.7 warps per issue stalled on the long scoreboard;
// profiles/r2_batched_warp_summary.txt).  Hence the software pipeline: the NEXT tile of a chain is requested before
// the current one is worked on -- trailing tiles into registers (eight warps leave 255 registers per thread), the next
// reflector tile by cp.async into the other half of a double buffer.
template <int BW>                                               // warps = problems in flight per CTA
__global__ void __launch_bounds__(BW * 32, 1) batched_warp_kernel(const BatchArgs a)
{
    extern __shared__ __align__(16) float bsm[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *Vs = bsm + w * BW_FLOATS;                            // Rs (33-float rows) while this warp is a panel
    float *VsB = Vs + 1088;
    float *top0 = VsB + 1024, *top1 = top0 + 1024, *ts = top1 + 1024, *xs = ts + 64;
    const int p = a.p, q = a.q, K = p < q ? p : q, ldm = a.ldm;
    // the problems are dealt to the CTAs evenly up front (no tail of one late CTA); a CTA takes its share BW at a time
    const int per = a.nprob / (int)gridDim.x, rem = a.nprob % (int)gridDim.x;
    const int first = (int)blockIdx.x * per + ((int)blockIdx.x < rem ? (int)blockIdx.x : rem);
    const int count = per + ((int)blockIdx.x < rem ? 1 : 0);

    for (int b0 = 0; b0 < count; b0 += BW) {
        const bool live = b0 + w < count;
        const int b = first + b0 + w;
        float *M = a.A + (size_t)(live ? b : first) * a.strideA;
        float *tb = a.tauc + (size_t)(live ? b : first) * K * p * TB;
        auto tile = [&](int I, int J) { return M + (size_t)(J * TB + lane) * ldm + I * TB; };
        for (int k = 0; k < K; ++k) {
            if (live) {
                // ---- panel of step k: GEQRT + TSQRT chain (qrdecomp.c:532-557, 689-721); tile i + 1 is in flight during tile i
                float x[TB], nx[TB];
                load_col_gmem(x, tile(k, k));
                if (k + 1 < p) load_col_gmem(nx, tile(k + 1, k));
                float t = geqrt_warp(x, xs, lane);
                store_col_gmem(tile(k, k), x);
                tb[((size_t)k * p + k) * TB + lane] = t;
#pragma unroll
                for (int r = 0; r < TB; ++r) Vs[r * 33 + lane] = x[r];
                __syncwarp();
                for (int i = k + 1; i < p; ++i) {
#pragma unroll
                    for (int r = 0; r < TB; ++r) x[r] = nx[r];
                    if (i + 1 < p) load_col_gmem(nx, tile(i + 1, k));
                    float ti = tsqrt_warp(x, Vs, xs, top0, lane);
                    store_col_gmem(tile(i, k), x);
                    tb[((size_t)k * p + i) * TB + lane] = ti;
                }
                float *rcol = tile(k, k);
#pragma unroll
                for (int r = 0; r < TB; ++r)
                    if (r <= lane) rcol[r] = Vs[r * 33 + lane];
            }
            __syncthreads();
            if (live) {
                // ---- step k applied to the trailing tile columns, two at a time (qrdecomp.c:620-644, 736-762).
                // Items of a column pair: row k (UNMQR, the diagonal tile's reflectors), then rows k+1 .. p-1 (TSMQR).
                // While item i is applied, item i + 1 is on its way: its reflector tile by cp.async into the other
                // V buffer, its tau and its two trailing tiles into registers.
                for (int j = k + 1; j < q; j += 2) {
                    const bool two = j + 1 < q;
                    float x0[TB], x1[TB], n0[TB], n1[TB];
                    float ntau;
                    __syncwarp();
                    cp_async_col(Vs + lane * TB, tile(k, k));           // (this thread wrote this very column in the panel phase)
                    ntau = tb[((size_t)k * p + k) * TB + lane];
                    load_col_gmem(n0, tile(k, j));
                    if (two) load_col_gmem(n1, tile(k, j + 1));
                    for (int i = k; i < p; ++i) {
                        float *vcur = ((i - k) & 1) ? VsB : Vs, *vnext = ((i - k) & 1) ? Vs : VsB;
                        float *tcur = ts + ((i - k) & 1) * 32;
#pragma unroll
                        for (int r = 0; r < TB; ++r) { x0[r] = n0[r]; x1[r] = n1[r]; }
                        tcur[lane] = ntau;
                        cp_async_wait_all();
                        __syncwarp();                                    // reflector tile i and its tau are in shared memory
                        if (i + 1 < p) {
                            cp_async_col(vnext + lane * TB, tile(i + 1, k));
                            ntau = tb[((size_t)k * p + i + 1) * TB + lane];
                            load_col_gmem(n0, tile(i + 1, j));
                            if (two) load_col_gmem(n1, tile(i + 1, j + 1));
                        }
                        if (i == k) {
                            if (two) apply_reflectors_nc<2, false>(x0, x1, nullptr, nullptr, 0, vcur, tcur);
                            else apply_reflectors_nc<1, false>(x0, x0, nullptr, nullptr, 0, vcur, tcur);
#pragma unroll
                            for (int r = 0; r < TB; ++r) { top0[r * TB + lane] = x0[r]; if (two) top1[r * TB + lane] = x1[r]; }
                        } else {
                            if (two) apply_reflectors_nc<2, true>(x0, x1, top0 + lane, top1 + lane, TB, vcur, tcur);
                            else apply_reflectors_nc<1, true>(x0, x0, top0 + lane, top0 + lane, TB, vcur, tcur);
                            store_col_gmem(tile(i, j), x0);
                            if (two) store_col_gmem(tile(i, j + 1), x1);
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < TB; ++r) x0[r] = top0[r * TB + lane];
                    store_col_gmem(tile(k, j), x0);
                    if (two) {
#pragma unroll
                        for (int r = 0; r < TB; ++r) x1[r] = top1[r * TB + lane];
                        store_col_gmem(tile(k, j + 1), x1);
                    }
                }
            }
            __syncthreads();
        }
    }
}

// stack the R factors (upper triangle of the top n x n of each problem) of a level into the next
// level's input: dst is (nprob*n) x n, column-major, leading dimension nprob*n
__global__ void gather_r_kernel(const float *src, long long strideA, int ldm, int n, int nprob, float *dst)
{
    const long long total = (long long)nprob * n * n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int r = (int)(e % n);
        long long t = e / n;
        int b = (int)(t % nprob), c = (int)(t / nprob);
        float v = r <= c ? src[(size_t)b * strideA + (size_t)c * ldm + r] : 0.f;
        dst[(size_t)c * ((size_t)nprob * n) + (size_t)b * n + r] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Apply Q or Q^T of a factored matrix (tile layout of cudaQRTask / b200qr_factor_device) to C.
// The reference never forms or applies Q; this replays the flat tree:
//   Q^T C:  for k = 0..K-1:  GEQRT reflectors of tile (k,k) on C tile row k, then for l = k+1..p-1
//           the TSQRT reflectors of tile (l,k) on C tile rows (k, l)       (reflector 0 first);
//   Q C:    the same blocks in reverse order, reflectors of a block last-to-first.
// One warp owns 32 columns of C (a thread = one column), C tile row k resident in shared memory.
template <bool FWD>    // FWD = true: Q^T C
__device__ __forceinline__ void apply_block(float (&x)[TB], float *top, int top_rs, const float *vs, const float *ts,
                                            bool diag)
{
    if constexpr (FWD) {
        if (diag) apply_reflectors_nc<1, false>(x, x, nullptr, nullptr, 0, vs, ts);
        else apply_reflectors_nc<1, true>(x, x, top, top, top_rs, vs, ts);
    } else {
#pragma unroll 2
    for (int c = TB - 1; c >= 0; --c) {
        float v[TB];
        load_col_smem(v, vs + c * TB);
        if (diag) {
#pragma unroll
            for (int r = 0; r < TB; ++r) v[r] = (r > c) ? v[r] : ((r == c) ? 1.0f : 0.f);
        }
        const float a0 = diag ? 0.f : top[c * top_rs];
        const float beta = ts[c] * (a0 + dot32(x, v));
        if (!diag) top[c * top_rs] = a0 - beta;
#pragma unroll
        for (int r = 0; r < TB; ++r) x[r] = fmaf(-beta, v[r], x[r]);
    }
    }
}

template <bool FWD>
__global__ void __launch_bounds__(128) apply_q_kernel(const float *A, int ldm, const float *tauc, int p, int K, float *Cm,
                                                      int ldc, int ncols)
{
    __shared__ __align__(16) float sm[4][1024 + 1024 + 32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = (blockIdx.x * 4 + w) * TB + lane;
    if ((blockIdx.x * 4 + w) * TB >= ncols) return;
    float *Vs = sm[w], *top = Vs + 1024, *ts = top + 1024;
    float *ccol = Cm + (size_t)col * ldc;
    float x[TB];
    auto load_v = [&](int I, int J) {
        load_col_gmem(x, A + (size_t)(J * TB + lane) * ldm + I * TB);
        store_col_smem(Vs + lane * TB, x);
        ts[lane] = __ldcg(tauc + ((size_t)J * p + I) * TB + lane);
        __syncwarp();
    };
    for (int kk = 0; kk < K; ++kk) {
        const int k = FWD ? kk : K - 1 - kk;
        // C tile row k resident
        load_col_gmem(x, ccol + k * TB);
        if (FWD) {
            load_v(k, k);
            load_col_gmem(x, ccol + k * TB);
            apply_block<true>(x, nullptr, 0, Vs, ts, true);
        }
#pragma unroll
        for (int r = 0; r < TB; ++r) top[r * TB + lane] = x[r];
        __syncwarp();
        for (int ll = k + 1; ll < p; ++ll) {
            const int l = FWD ? ll : p - 1 - (ll - k - 1);
            load_v(l, k);
            load_col_gmem(x, ccol + l * TB);
            apply_block<FWD>(x, top + lane, TB, Vs, ts, false);
            store_col_gmem(ccol + l * TB, x);
            __syncwarp();
        }
#pragma unroll
        for (int r = 0; r < TB; ++r) x[r] = top[r * TB + lane];
        if (!FWD) {
            float keep[TB];
#pragma unroll
            for (int r = 0; r < TB; ++r) keep[r] = x[r];
            load_v(k, k);                       // clobbers x
            apply_block<false>(keep, nullptr, 0, Vs, ts, true);
#pragma unroll
            for (int r = 0; r < TB; ++r) x[r] = keep[r];
        }
        store_col_gmem(ccol + k * TB, x);
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------------
// Streaming TSQR: the TSQRT recurrence (qrdecomp.c:689-721 STSQRF, 1243-1272 calcvkDouble, 647-687
// updateDoubleQ_WY) with the tile as wide as the matrix.  One CTA owns a strip of rows and walks down it in blocks
// of SRB rows; thread = matrix column, its SRB elements of the block in registers.  Per reflector c: the owner thread
// derives (1/piv, tau) from R(c,c) and its column's sum of squares and broadcasts the column through shared memory;
// every thread j >= c does one dot product and one axpy of length SRB; the reflector V2(:,c) = x / piv replaces
// column c of the block in place.  The n reflectors of a block are a flat-tree TSQRT step exactly like the
// reference's, only n columns wide instead of 32.  The strips' R factors are stacked and reduced by the leaf-batch
// tree above.
//
// What bounds it (ncu, profiles/r1_tsqr_strip_summary.txt): the n reflectors of a block are strictly serial and one
// step is a chain of dependent latencies -- broadcast, dot, update, the next owner's sum of squares, rsqrt / rcp,
// barrier: ~1050 cycles for 64 rows, 48 % of all warp samples wait at the barrier, the FMA pipe is 18 % busy.  So
// the triangle R lives in global memory (L2-resident, 131 KB per strip; a thread only ever touches its own column
// of it and prefetches row c + 1 during step c), which leaves shared memory nearly empty and lets TWO CTAs share an
// SM (128 registers each): two independent chains per SM.

constexpr int TREE_RB = 32;                 // block rows of the reduction levels (divides every n)

struct StripArgs {
    float *A;                               // rows x n, column-major
    long long ldm, rows, rows_per_strip, blocks_per_strip;
    int n;
    float *tau;                             // [strip * blocks_per_strip + block][n]
    float *Rws;                             // [strip][n (n + 1) / 2]: the strip's triangle, packed by rows, zero on entry
    float *Rout;                            // strip s writes its R at rows s*n .. of a column-major stack
    long long ldr;
};

// SRB = rows per block (a thread's block column lives in registers), MINB = CTAs per SM the register budget allows:
// <48, 2> is the default (117 registers, no spills; 64 rows spill at 128), <32, 3> trades block height for a third chain.
// TRI = the rows are a stack of n x n upper-triangular factors (a reduction level): the strip's first factor IS the
// initial triangle, and a block that starts at row r of a factor has nothing but zeros in columns < r, whose
// reflectors are the identity: its steps start at c = r (640 instead of 1024 steps for a pair with 64-row blocks).
template <int SRB, int MINB, bool TRI>
__global__ void __launch_bounds__(256, MINB) tsqr_strip_kernel(const StripArgs a)
{
    __shared__ __align__(16) float xs[2 * SRB];             // two broadcast slots
    __shared__ float hs[4];                                 // two slots of {1/piv, tau}
    const int n = a.n, j = threadIdx.x, warp = j >> 5;
    float *Rg = a.Rws + (size_t)blockIdx.x * (n * (n + 1) / 2);    // (c, col) at c*n - c*(c-1)/2 + (col - c)
    const long long row0 = (long long)blockIdx.x * a.rows_per_strip;
    long long row1 = row0 + a.rows_per_strip;
    if (row1 > a.rows) row1 = a.rows;
    float *col = a.A + (size_t)j * a.ldm;
    float *tau_out = a.tau + (size_t)blockIdx.x * a.blocks_per_strip * n;
    if (TRI) {                                              // thread j: column j of the first factor -> its triangle
        for (int c = 0; c <= j; ++c) __stcg(Rg + c * n - (c * (c - 1)) / 2 + (j - c), col[row0 + c]);
    }

    for (long long r0 = row0 + (TRI ? n : 0); r0 < row1; r0 += SRB, tau_out += n) {
        const int nr = (int)(row1 - r0 < SRB ? row1 - r0 : SRB);     // a multiple of 4
        const int c0 = TRI ? (int)((r0 - row0) % n) : 0;    // first column of this block that is not all zeros
        // Start-up of a strip (the triangle is empty when it begins): after `seen` rows the stack [R; block] has rank
        // <= seen + SRB, so only the reflectors c < c_end are determined by the data.  Behind them the block's columns
        // are zero in exact arithmetic -- rounding noise here, whose sign would pick the sign of that row of R when
        // the next block gives it mass -- so the block stops there: H = I (tau = 0, V = 0) for c >= c_end.  The oracle
        // does the same (oracle_tsqrt_lead), which is what pins R and every reflector block at any width n.
        const long long seen = r0 - row0;
        const int c_end = (TRI || seen + SRB >= n) ? n : (int)(seen + SRB);
        float rc_next = 0.f;                                // R(c0, j): this thread's own store of an earlier block
        if (j >= c0) rc_next = __ldcg(Rg + c0 * n - (c0 * (c0 - 1)) / 2 + (j - c0));
        float b[SRB];
#pragma unroll
        for (int r = 0; r < SRB; r += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nr) v = __ldcs(reinterpret_cast<const float4 *>(col + r0 + r));
            b[r] = v.x; b[r + 1] = v.y; b[r + 2] = v.z; b[r + 3] = v.w;
        }
        auto sumsq = [&]() {
            float s[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] = b[i] * b[i];
#pragma unroll
            for (int r = 8; r < SRB; r += 8)
#pragma unroll
                for (int i = 0; i < 8; ++i) s[i] = fmaf(b[r + i], b[r + i], s[i]);
            return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
        };
        float ssq = sumsq();
        if (TRI) __syncthreads();                           // the broadcast slots of the previous block (c0 may be odd or even)
#pragma unroll 1
        for (int c = c0; c < c_end; ++c) {
            float *x = xs + (c & 1) * SRB;
            float *h2 = hs + (c & 1) * 2;
            const int off = c * n - (c * (c - 1)) / 2;
            const float rc = rc_next;                       // R(c, j), valid for j >= c
            if (j == c) {                                   // owner: broadcast the column and its Householder scalars
#pragma unroll
                for (int r = 0; r < SRB; r += 4) *reinterpret_cast<float4 *>(x + r) = make_float4(b[r], b[r + 1], b[r + 2], b[r + 3]);
                const HouseScalars h = house_scalars(rc, ssq);
                h2[0] = h.inv; h2[1] = h.tau;
                tau_out[c] = h.tau;
            }
            if (j > c) rc_next = __ldcg(Rg + off + (n - c) + (j - c - 1));     // R(c + 1, j): in flight during this step
            __syncthreads();
            const float inv = h2[0], tau = h2[1];
            // the reflector replaces column c of the block
            {
                float *vcol = a.A + (size_t)c * a.ldm + r0;
                for (int t = j; t < nr; t += n) vcol[t] = x[t] * inv;
            }
            if (j >= c) {
                // beta = tau (R(c,j) + v' B(:,j)),  v = x / piv
                float s[8];
#pragma unroll
                for (int i = 0; i < 8; i += 4) {
                    const float4 xv = *reinterpret_cast<const float4 *>(x + i);
                    s[i] = b[i] * xv.x; s[i + 1] = b[i + 1] * xv.y; s[i + 2] = b[i + 2] * xv.z; s[i + 3] = b[i + 3] * xv.w;
                }
#pragma unroll
                for (int r = 8; r < SRB; r += 8)
#pragma unroll
                    for (int i = 0; i < 8; i += 4) {
                        const float4 xv = *reinterpret_cast<const float4 *>(x + r + i);
                        s[i] = fmaf(b[r + i], xv.x, s[i]); s[i + 1] = fmaf(b[r + i + 1], xv.y, s[i + 1]);
                        s[i + 2] = fmaf(b[r + i + 2], xv.z, s[i + 2]); s[i + 3] = fmaf(b[r + i + 3], xv.w, s[i + 3]);
                    }
                const float d = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
                const float beta = tau * fmaf(inv, d, rc);
                __stcg(Rg + off + j - c, rc - beta);
                if (j > c) {
                    const float g = -beta * inv;
#pragma unroll
                    for (int r = 0; r < SRB; r += 4) {
                        const float4 xv = *reinterpret_cast<const float4 *>(x + r);
                        b[r] = fmaf(g, xv.x, b[r]); b[r + 1] = fmaf(g, xv.y, b[r + 1]);
                        b[r + 2] = fmaf(g, xv.z, b[r + 2]); b[r + 3] = fmaf(g, xv.w, b[r + 3]);
                    }
                    if (warp == ((c + 1) >> 5)) ssq = sumsq();      // the next owner lives in this warp
                }
            }
        }
        if (!TRI && j >= c_end) {                           // start-up block: columns behind the last determined reflector
            for (int r = 0; r < nr; r += 4) *reinterpret_cast<float4 *>(col + r0 + r) = make_float4(0.f, 0.f, 0.f, 0.f);
            tau_out[j] = 0.f;
        }
        if (!TRI && c_end < n) __syncthreads();             // (the broadcast slots: the next block starts at c = 0 again)
    }
    float *rcol = a.Rout + (size_t)j * a.ldr + (size_t)blockIdx.x * n;
    for (int c = 0; c < n; ++c) rcol[c] = (c <= j) ? __ldcg(Rg + c * n - (c * (c - 1)) / 2 + (j - c)) : 0.f;
}


// Two variants of this kernel were measured in round 1 and removed in round 2 (profiles/r1_measurements.md): a
// two-barrier form in which the next owner broadcasts its column before its scalar work (56.9 ms against 48.0 ms at
// 4 194 304 x 256) and a form with eight reflectors per barrier and a compact-WY sweep by the other warps (58.9 ms).

// Round 2 tried a second mapping for this kernel -- four matrix columns per thread, n / 4 threads per strip, the
// broadcast column read once into registers for all four (8x less shared-memory traffic per FMA), four strips per SM --
// and measured it slower (profiles/r2_measurements.md): 47.1 ms for the strip kernel alone against 42 ms, because with
// 592 strips the triangles (78 MB) no longer stay in L2 while 8.6 GB of matrix stream past (ncu: 21.6 GB of DRAM writes
// per launch = the 4.3 GB of reflectors + every R store, 131 KB per 32-row block; an L2 persisting window changed
// nothing), and its pairwise levels took 0.92 ms each against 0.55 ms.  Removed; the kernel above is the engine.


// ---- exchange step of the multi-GPU TSQR over peer memory (b200qr_tsqr_dist_*): flags are epoch numbers, monotonic
__global__ void tsqr_signal_kernel(int *const *peer_flags, int world, int rank, int epoch)
{
    const int r = threadIdx.x;
    if (r < world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peer_flags[r] + rank), "r"(epoch) : "memory");
    }
}
__global__ void tsqr_wait_kernel(const int *flags, int world, int epoch, unsigned long long limit_ns)
{
    const int r = threadIdx.x;
    if (r < world) {
        unsigned long long t0 = 0;
        unsigned spins = 0;
        for (;;) {
            int v;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
            if (v >= epoch) break;
            __nanosleep(200);
            if (((++spins) & 0xfffu) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > limit_ns) __trap();        // a peer never arrived: fail loudly instead of hanging the GPU
            }
        }
    }
    __syncthreads();
}

}  // namespace qrb200

using namespace qrb200;

static thread_local int g_batched_cuda = 0;
#define CKB(call)                                                           \
    do {                                                                    \
        cudaError_t e_ = (call);                                            \
        if (e_ != cudaSuccess) { g_batched_cuda = (int)e_; return B200QR_ECUDA; } \
    } while (0)

static int batched_launch(float *dA, int m, int n, int ldm, long long strideA, int nprob, float *dTau, int *dCounter,
                          cudaStream_t stream)
{
    // the shared-memory opt-in is a per-device attribute and the grid a per-device number: both are taken for the
    // CURRENT device on every call (cheap runtime calls), so a process may use several GPUs and threads
    int dev = 0, sms = 0;
    CKB(cudaGetDevice(&dev));
    CKB(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    BatchArgs a;
    a.A = dA; a.strideA = strideA; a.ldm = ldm; a.p = m / TB; a.q = n / TB; a.nprob = nprob;
    a.tauc = dTau; a.counter = dCounter;
    const char *force = getenv("QRB200_BATCHED");           // "colwarp" / "warp": A/B runs
    const bool few = nprob < 2 * sms;                        // too few problems to give the warps of the device one each
    if ((few && !(force && force[0] == 'w')) || (force && force[0] == 'c')) {
        // one CTA per problem, warp w = tile column w: more parallelism inside a problem (the tree levels of the TSQR)
        const int smem = BQ * WARP_FLOATS * (int)sizeof(float);
        int occ = 0;
        CKB(cudaFuncSetAttribute(batched_tiledqr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CKB(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, batched_tiledqr_kernel, BQ * 32, smem));
        const int grid_cap = sms * (occ > 0 ? occ : 1);
        CKB(cudaMemsetAsync(dCounter, 0, sizeof(int), stream));
        batched_tiledqr_kernel<<<nprob < grid_cap ? nprob : grid_cap, BQ * 32, smem, stream>>>(a);
    } else {
        // eight warps per CTA: 255 registers per thread hold the tile being worked on and the prefetched next one
        const int smem = 8 * BW_FLOATS * (int)sizeof(float);
        const int grid = nprob < sms ? nprob : sms;
        CKB(cudaFuncSetAttribute(batched_warp_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        batched_warp_kernel<8><<<grid, 8 * 32, smem, stream>>>(a);
    }
    CKB(cudaGetLastError());
    return B200QR_OK;
}

struct b200qr_tsqr_s {
    long long m;
    int n, leaf_rows, fanin;
    int *d_counter = nullptr;
    float *d_leaf_tau = nullptr;
    struct Level { long long count; int nfull, rest; float *stack = nullptr, *tau = nullptr; };
    std::vector<Level> levels;          // reduction levels: `count` R factors in, nfull + (rest > 0) out
    // streaming engine (b200qr_tsqr_create_streaming): strips of rows -> one R factor each, stacked in levels[0].stack
    // (or written to dR when there is one strip); the levels above reduce them
    int strips = 0, srb = 48;
    long long rows_per_strip = 0, blocks_per_strip = 0;
    float *d_stream_tau = nullptr, *d_rws = nullptr;
    // the strips' factors are reduced pairwise by the same kernel (TRI): level l holds `count` factors in `stack`
    struct SLevel { int count; float *stack = nullptr, *tau = nullptr; };
    std::vector<SLevel> slevels;
    bool rstack = false;                // b200qr_tsqr_create_rstack: the input IS level 0's stack of factors
    bool streaming = false;
    bool leaf_tiles = false;            // b200qr_tsqr_create_tiles: leaves by the batched tile kernel, pairwise levels above
};

extern "C" {

int b200qr_batched_workspace_floats(int m, int n, int nprob, long long *tau_floats)
{
    if (m <= 0 || n <= 0 || m % TB || n % TB || n > BQ * TB || nprob <= 0 || !tau_floats) return B200QR_EINVAL;
    const int p = m / TB, q = n / TB, K = p < q ? p : q;
    *tau_floats = (long long)nprob * K * p * TB;
    return B200QR_OK;
}

// nprob independent m x n problems (n <= 256), problem b at dA + b*strideA, in place; dTau receives
// nprob x K*p*32 floats (compact tau per problem); dCounter is one int of device scratch.
int b200qr_batched_factor(float *dA, int m, int n, int ldm, long long strideA, int nprob, float *dTau, int *dCounter,
                          void *stream)
{
    if (!dA || !dTau || !dCounter || m <= 0 || n <= 0 || m % TB || n % TB || n > BQ * TB || ldm < m || (ldm & 3) ||
        ((uintptr_t)dA & 15) || (strideA & 3) || nprob <= 0)
        return B200QR_EINVAL;
    return batched_launch(dA, m, n, ldm, strideA, nprob, dTau, dCounter, (cudaStream_t)stream);
}

int b200qr_batched_last_cuda_error(void) { return g_batched_cuda; }

// C <- Q^T C (transpose != 0) or Q C, for the Q of a matrix factored by cudaQRTask /
// b200qr_factor_device / b200qr_batched_factor: dA = factored m x n matrix (ldm), dTauCompact =
// [(k*p + i)*32 + r] (b200qr_plan_tau_device), dC = m x ncols (ldc), ncols % 32 == 0.
int b200qr_apply_q(const float *dA, int m, int n, int ldm, const float *dTauCompact, float *dC, int ldc, int ncols,
                   int transpose, void *stream)
{
    if (!dA || !dTauCompact || !dC || m <= 0 || n <= 0 || m % TB || n % TB || ncols <= 0 || ncols % TB || ldm < m ||
        ldc < m || (ldm & 3) || (ldc & 3))
        return B200QR_EINVAL;
    const int p = m / TB, q = n / TB, K = p < q ? p : q;
    const int blocks = (ncols / TB + 3) / 4;
    if (transpose) apply_q_kernel<true><<<blocks, 128, 0, (cudaStream_t)stream>>>(dA, ldm, dTauCompact, p, K, dC, ldc, ncols);
    else apply_q_kernel<false><<<blocks, 128, 0, (cudaStream_t)stream>>>(dA, ldm, dTauCompact, p, K, dC, ldc, ncols);
    CKB(cudaGetLastError());
    return B200QR_OK;
}

// ---- TSQR ------------------------------------------------------------------------------------
static int build_levels(b200qr_tsqr_s *t, long long count);
static int build_slevels(b200qr_tsqr_s *t, int count, bool borrowed);
// Level structure: `count` R factors (n x n) are stacked into a (count*n) x n matrix and factored in
// groups of `fanin` consecutive factors (one batched problem of fanin*n rows each; a shorter last
// group of >= 2 factors is its own problem, a single left-over factor passes through).  The group
// tops, all fanin*n rows apart, are the next level's R factors.
int b200qr_tsqr_create(b200qr_tsqr_t *out, long long m, int n, int leaf_rows, int fanin)
{
    if (!out || n <= 0 || n % TB || n > BQ * TB || leaf_rows < n || leaf_rows % TB || m <= 0 || m % leaf_rows ||
        fanin < 2 || fanin > 64)
        return B200QR_EINVAL;
    b200qr_tsqr_s *t = new b200qr_tsqr_s;
    t->m = m; t->n = n; t->leaf_rows = leaf_rows; t->fanin = fanin;
    const int q = n / TB;
    const long long nleaf = m / leaf_rows;
    auto fail = [&]() { b200qr_tsqr_destroy(t); return B200QR_ENOMEM; };
    if (cudaMalloc(&t->d_counter, sizeof(int)) != cudaSuccess) return fail();
    if (cudaMalloc(&t->d_leaf_tau, sizeof(float) * (size_t)nleaf * q * (leaf_rows / TB) * TB) != cudaSuccess) return fail();
    if (build_levels(t, nleaf) != B200QR_OK) return fail();
    *out = t;
    return B200QR_OK;
}

// Level structure shared by both engines: `count` R factors in, fanin at a time
static int build_levels(b200qr_tsqr_s *t, long long count)
{
    const int n = t->n, q = n / TB, fanin = t->fanin;
    while (count > 1) {
        b200qr_tsqr_s::Level L;
        L.count = count;
        L.nfull = (int)(count / fanin);
        L.rest = (int)(count % fanin);
        if (cudaMalloc(&L.stack, sizeof(float) * (size_t)count * n * n) != cudaSuccess) return B200QR_ENOMEM;
        // tau: q * (rows/32) * 32 floats per problem = q * rows floats
        size_t taufl = (size_t)L.nfull * q * fanin * n + (size_t)q * (L.rest > 0 ? L.rest : 1) * n;
        if (cudaMalloc(&L.tau, sizeof(float) * taufl) != cudaSuccess) { cudaFree(L.stack); return B200QR_ENOMEM; }
        t->levels.push_back(L);
        count = L.nfull + (L.rest > 0 ? 1 : 0);
    }
    return B200QR_OK;
}

// Pairwise reduction levels of the streaming engine: level l holds `count` factors in its stack (level 0's stack is the
// caller's matrix when `borrowed`).  tau is zeroed once: the steps a triangular block skips keep tau = 0 (H = I).
static int build_slevels(b200qr_tsqr_s *t, int count, bool borrowed)
{
    const int n = t->n;
    for (; count > 1; count = (count + 1) / 2) {
        b200qr_tsqr_s::SLevel L;
        L.count = count;
        const size_t taufl = (size_t)((count + 1) / 2) * (n / TREE_RB) * n;   // per pair: the second factor's blocks
        if (!(borrowed && t->slevels.empty()) && cudaMalloc(&L.stack, sizeof(float) * (size_t)count * n * n) != cudaSuccess)
            return B200QR_ENOMEM;
        if (cudaMalloc(&L.tau, sizeof(float) * taufl) != cudaSuccess) { cudaFree(L.stack); return B200QR_ENOMEM; }
        if (cudaMemset(L.tau, 0, sizeof(float) * taufl) != cudaSuccess) { cudaFree(L.stack); cudaFree(L.tau); return B200QR_ECUDA; }
        t->slevels.push_back(L);
    }
    return B200QR_OK;
}

// Reduction of `count` stacked n x n upper-triangular factors (a (count*n) x n column-major matrix: the exchange step
// of the multi-GPU TSQR delivers exactly this) to one, by the pairwise levels of the streaming engine.  Factor with
// b200qr_tsqr_factor(t, dStack, ld >= count*n, dR): the stack is overwritten with the tree's reflectors.
int b200qr_tsqr_create_rstack(b200qr_tsqr_t *out, int count, int n)
{
    if (!out || n <= 0 || n % TB || n > BQ * TB || count < 1) return B200QR_EINVAL;
    b200qr_tsqr_s *t = new b200qr_tsqr_s;
    t->m = (long long)count * n; t->n = n; t->leaf_rows = 0; t->fanin = 2; t->streaming = true; t->rstack = true;
    t->strips = count;
    auto fail = [&]() { b200qr_tsqr_destroy(t); return B200QR_ENOMEM; };
    if (cudaMalloc(&t->d_rws, sizeof(float) * (size_t)((count + 1) / 2) * (n * (n + 1) / 2)) != cudaSuccess) return fail();
    if (build_slevels(t, count, true) != B200QR_OK) return fail();
    *out = t;
    return B200QR_OK;
}

// Streaming engine: `strips` CTAs (<= 0: two per SM, which is what fits) each reduce a contiguous strip of rows to one
// R factor; the factors are reduced by the leaf-batch tree, four at a time.  m % 32 == 0, n % 32 == 0, n <= 256.
int b200qr_tsqr_create_streaming(b200qr_tsqr_t *out, long long m, int n, int strips)
{
    if (!out || n <= 0 || n % TB || n > BQ * TB || m <= 0 || m % TB) return B200QR_EINVAL;
    const bool strips_auto = strips <= 0;
    int sms = 0;
    if (strips_auto) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            return B200QR_ECUDA;
    }
    b200qr_tsqr_s *t = new b200qr_tsqr_s;
    t->m = m; t->n = n; t->leaf_rows = 0; t->fanin = 4; t->streaming = true;
    auto fail = [&]() { b200qr_tsqr_destroy(t); return B200QR_ENOMEM; };
    if (const char *e = getenv("QRB200_TSQR_RB")) { const int v = atoi(e); if (v == 32 || v == 48 || v == 64) t->srb = v; }
    const int SRB = t->srb;
    if (strips_auto) strips = sms * (SRB == 32 ? 3 : 2);
    long long rps = (m + strips - 1) / strips;
    rps = (rps + SRB - 1) / SRB * SRB;
    t->rows_per_strip = rps;
    t->blocks_per_strip = rps / SRB;
    t->strips = (int)((m + rps - 1) / rps);
    if (cudaMalloc(&t->d_counter, sizeof(int)) != cudaSuccess) return fail();
    if (cudaMalloc(&t->d_stream_tau, sizeof(float) * (size_t)t->strips * t->blocks_per_strip * n) != cudaSuccess) return fail();
    if (cudaMalloc(&t->d_rws, sizeof(float) * (size_t)t->strips * (n * (n + 1) / 2)) != cudaSuccess) return fail();
    if (build_slevels(t, t->strips, false) != B200QR_OK) return fail();
    *out = t;
    return B200QR_OK;
}

// Tile engine for the leaves: the rows are cut into `leaves` equal leaves (<= 0: as many as the device has warps for,
// sixteen per SM, among the divisors of m that leave whole tile rows), every leaf is one problem of the batched tile
// kernel -- one warp per leaf, a flat tree of 32 x 32 TSQRT tiles down each of its n / 32 tile columns, exactly the
// reference's chain (gridscheduler.c:136-154) per leaf -- and the leaves' R factors are reduced pairwise by the
// streaming engine's triangular levels.  Reflectors: leaves in dA in the reference's tile layout, tau in the plan.
int b200qr_tsqr_create_tiles(b200qr_tsqr_t *out, long long m, int n, int leaves)
{
    if (!out || n <= 0 || n % TB || n > BQ * TB || m <= 0 || m % TB || m < n) return B200QR_EINVAL;
    if (leaves <= 0) {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            return B200QR_ECUDA;
        for (leaves = sms * 16; leaves > 1; --leaves)
            if (m % leaves == 0 && (m / leaves) % TB == 0 && m / leaves >= n) break;
    }
    if (leaves < 1 || m % leaves || (m / leaves) % TB || m / leaves < n || m / leaves > 0x7fffffffLL) return B200QR_EINVAL;
    b200qr_tsqr_s *t = new b200qr_tsqr_s;
    t->m = m; t->n = n; t->leaf_rows = (int)(m / leaves); t->fanin = 2; t->streaming = true; t->leaf_tiles = true;
    t->strips = leaves;
    auto fail = [&]() { b200qr_tsqr_destroy(t); return B200QR_ENOMEM; };
    if (cudaMalloc(&t->d_counter, sizeof(int)) != cudaSuccess) return fail();
    if (cudaMalloc(&t->d_leaf_tau, sizeof(float) * (size_t)m * (n / TB)) != cudaSuccess) return fail();     // per leaf: q * leaf_rows
    if (cudaMalloc(&t->d_rws, sizeof(float) * (size_t)((leaves + 1) / 2) * (n * (n + 1) / 2)) != cudaSuccess) return fail();
    if (build_slevels(t, leaves, false) != B200QR_OK) return fail();
    *out = t;
    return B200QR_OK;
}

int b200qr_tsqr_leaves(b200qr_tsqr_t t) { return t ? t->strips : B200QR_EINVAL; }

int b200qr_tsqr_destroy(b200qr_tsqr_t t)
{
    if (!t) return B200QR_EINVAL;
    cudaFree(t->d_counter); cudaFree(t->d_leaf_tau);
    for (auto &L : t->levels) { cudaFree(L.stack); cudaFree(L.tau); }
    cudaFree(t->d_stream_tau); cudaFree(t->d_rws);
    for (auto &L : t->slevels) { cudaFree(L.stack); cudaFree(L.tau); }
    delete t;
    return B200QR_OK;
}

int b200qr_tsqr_stream_block_rows(b200qr_tsqr_t t) { return t && t->streaming ? t->srb : B200QR_EINVAL; }

int b200qr_tsqr_levels(b200qr_tsqr_t t) { return t ? (t->streaming ? (int)t->slevels.size() + 1 : (int)t->levels.size()) : B200QR_EINVAL; }

// Factor the m x n matrix dA (column-major, ldm) in place (leaf reflectors) and write the n x n
// upper-triangular R (column-major, leading dimension n) to dR.  Asynchronous on `stream`.
int b200qr_tsqr_factor(b200qr_tsqr_t t, float *dA, long long ldm, float *dR, void *stream_)
{
    if (!t || !dA || !dR || ldm < t->m || (ldm & 3) || ldm > 0x7fffffffLL) return B200QR_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int n = t->n, q = n / TB, fanin = t->fanin;
    int rc;
    const float *src = dA;
    long long src_stride = t->leaf_rows;
    int src_ld = (int)ldm;
    bool first_stack_filled = false;
    if (t->streaming) {
        StripArgs a;
        a.n = n; a.Rws = t->d_rws;
        if (t->leaf_tiles) {
            rc = batched_launch(dA, t->leaf_rows, n, (int)ldm, t->leaf_rows, t->strips, t->d_leaf_tau, t->d_counter, stream);
            if (rc != B200QR_OK) return rc;
            gather_r_kernel<<<1024, 256, 0, stream>>>(dA, t->leaf_rows, (int)ldm, n, t->strips, t->strips > 1 ? t->slevels[0].stack : dR);
            CKB(cudaGetLastError());
        } else if (!t->rstack) {
            CKB(cudaMemsetAsync(t->d_rws, 0, sizeof(float) * (size_t)t->strips * (n * (n + 1) / 2), stream));
            a.A = dA; a.ldm = ldm; a.rows = t->m; a.rows_per_strip = t->rows_per_strip; a.blocks_per_strip = t->blocks_per_strip;
            a.tau = t->d_stream_tau;
            a.Rout = t->strips > 1 ? t->slevels[0].stack : dR;        // the strips' triangles ARE the first level's stack
            a.ldr = t->strips > 1 ? (long long)t->strips * n : n;
            if (t->srb == 32) tsqr_strip_kernel<32, 3, false><<<t->strips, n, 0, stream>>>(a);
            else if (t->srb == 64) tsqr_strip_kernel<64, 2, false><<<t->strips, n, 0, stream>>>(a);
            else tsqr_strip_kernel<48, 2, false><<<t->strips, n, 0, stream>>>(a);
            CKB(cudaGetLastError());
        } else if (t->strips == 1) {                                  // a stack of one factor: R is that factor
            gather_r_kernel<<<64, 256, 0, stream>>>(dA, 0, (int)ldm, n, 1, dR);
            CKB(cudaGetLastError());
        }
        // pairwise reduction of the stacked factors: level l's strips write level l + 1's stack, the last one dR
        for (size_t l = 0; l < t->slevels.size(); ++l) {
            auto &L = t->slevels[l];
            const int pairs = (L.count + 1) / 2;
            const bool borrowed = t->rstack && l == 0;
            a.A = borrowed ? dA : L.stack; a.ldm = borrowed ? ldm : (long long)L.count * n; a.rows = (long long)L.count * n;
            a.rows_per_strip = 2 * n; a.blocks_per_strip = n / TREE_RB;
            a.tau = L.tau;                                           // a.Rws: the first `pairs` triangles are reused
            a.Rout = pairs > 1 ? t->slevels[l + 1].stack : dR;
            a.ldr = pairs > 1 ? (long long)pairs * n : n;
            tsqr_strip_kernel<TREE_RB, 3, true><<<pairs, n, 0, stream>>>(a);
            CKB(cudaGetLastError());
        }
        return B200QR_OK;
    } else {
        long long count = t->m / t->leaf_rows;
        rc = batched_launch(dA, t->leaf_rows, n, (int)ldm, t->leaf_rows, (int)count, t->d_leaf_tau, t->d_counter, stream);
        if (rc != B200QR_OK) return rc;
    }
    for (auto &L : t->levels) {
        const int sld = (int)(L.count * n);
        if (!first_stack_filled) {
            gather_r_kernel<<<1024, 256, 0, stream>>>(src, src_stride, src_ld, n, (int)L.count, L.stack);
            CKB(cudaGetLastError());
        }
        first_stack_filled = false;
        if (L.nfull > 0) {
            rc = batched_launch(L.stack, fanin * n, n, sld, (long long)fanin * n, L.nfull, L.tau, t->d_counter, stream);
            if (rc != B200QR_OK) return rc;
        }
        if (L.rest >= 2) {
            rc = batched_launch(L.stack + (size_t)L.nfull * fanin * n, L.rest * n, n, sld, 0, 1,
                                L.tau + (size_t)L.nfull * q * fanin * n, t->d_counter, stream);
            if (rc != B200QR_OK) return rc;
        }
        src = L.stack; src_stride = (long long)fanin * n; src_ld = sld;
    }
    gather_r_kernel<<<64, 256, 0, stream>>>(src, 0, src_ld, n, 1, dR);       // top n x n, lower triangle zeroed
    CKB(cudaGetLastError());
    return B200QR_OK;
}

}  // extern "C"


// ---------------------------------------------------------------------------------------------
// Multi-GPU TSQR behind the C ABI (BASELINE.json configs[3] at 2 / 4 / 8 GPUs): rank r owns a contiguous block
// of m_local rows.  factor(): (1) the rank reduces its block to an n x n factor with the single-GPU engine; (2) the
// exchange -- the path's one real communication step, W x n x n floats -- happens over peer memory: the rank copies its
// factor into slot `rank` of EVERY rank's stack buffer (cudaMemcpy2DAsync over NVLink, on the caller's stream), raises
// its epoch flag in every rank's flag array (system-scope release) and waits until all W flags of its own array show
// the epoch (system-scope acquire); (3) every rank reduces the stack of W factors redundantly (b200qr_tsqr_create_rstack),
// so all ranks end with the same R.  No host synchronisation, no collective library: the caller only moves the handle
// blobs once (export -> all-gather by any means -> connect).  Two stack buffers alternate by epoch, so a rank that is
// already in its next factorisation cannot overwrite a stack a slower peer is still reducing (it cannot get two epochs
// ahead: its wait needs the peer's next flag).
struct tsqr_dist_blob { cudaIpcMemHandle_t mem; int n, world, rank; long long m_local; };

struct b200qr_tsqr_dist_s {
    long long m_local;
    int n, rank, world, epoch = 0;
    b200qr_tsqr_t local = nullptr, top = nullptr;
    unsigned char *d_mem = nullptr;            // [2 stacks of world*n*n floats][flags: world ints (padded to 256 B)]
    unsigned char *peer_mem[8] = {};
    void *opened[8] = {};
    float *d_rlocal = nullptr;
    int **d_peer_flags = nullptr;              // device array of world pointers
    int connected = 0;
    size_t stack_bytes() const { return sizeof(float) * (size_t)world * n * n; }
    float *stack(unsigned char *base, int which) const { return reinterpret_cast<float *>(base + which * stack_bytes()); }
    int *flags(unsigned char *base) const { return reinterpret_cast<int *>(base + 2 * stack_bytes()); }
};

extern "C" {

int b200qr_tsqr_dist_create(b200qr_tsqr_dist_t *out, long long m_local, int n, int rank, int world, int tiles)
{
    if (!out || world < 1 || world > 8 || rank < 0 || rank >= world || n <= 0 || n % TB || n > BQ * TB || m_local < n || m_local % TB)
        return B200QR_EINVAL;
    b200qr_tsqr_dist_s *h = new b200qr_tsqr_dist_s;
    h->m_local = m_local; h->n = n; h->rank = rank; h->world = world;
    int rc = tiles ? b200qr_tsqr_create_tiles(&h->local, m_local, n, 0) : b200qr_tsqr_create_streaming(&h->local, m_local, n, 0);
    if (rc == B200QR_OK && world > 1) rc = b200qr_tsqr_create_rstack(&h->top, world, n);
    const size_t bytes = 2 * h->stack_bytes() + 256;
    if (rc == B200QR_OK && cudaMalloc(&h->d_mem, bytes) != cudaSuccess) rc = B200QR_ENOMEM;
    if (rc == B200QR_OK && cudaMalloc(&h->d_rlocal, sizeof(float) * (size_t)n * n) != cudaSuccess) rc = B200QR_ENOMEM;
    if (rc == B200QR_OK && cudaMalloc(&h->d_peer_flags, sizeof(int *) * 8) != cudaSuccess) rc = B200QR_ENOMEM;
    if (rc == B200QR_OK && cudaMemset(h->d_mem, 0, bytes) != cudaSuccess) rc = B200QR_ECUDA;
    if (rc != B200QR_OK) { b200qr_tsqr_dist_destroy(h); return rc; }
    h->peer_mem[rank] = h->d_mem;
    if (world == 1) h->connected = 1;
    *out = h;
    return B200QR_OK;
}

int b200qr_tsqr_dist_destroy(b200qr_tsqr_dist_t h)
{
    if (!h) return B200QR_EINVAL;
    for (void *p : h->opened) if (p) cudaIpcCloseMemHandle(p);
    if (h->local) b200qr_tsqr_destroy(h->local);
    if (h->top) b200qr_tsqr_destroy(h->top);
    cudaFree(h->d_mem); cudaFree(h->d_rlocal); cudaFree(h->d_peer_flags);
    delete h;
    return B200QR_OK;
}

int b200qr_tsqr_dist_handle_bytes(void) { return (int)sizeof(tsqr_dist_blob); }

int b200qr_tsqr_dist_export(b200qr_tsqr_dist_t h, void *out_bytes)
{
    if (!h || !out_bytes) return B200QR_EINVAL;
    tsqr_dist_blob b;
    memset(&b, 0, sizeof b);
    CKB(cudaIpcGetMemHandle(&b.mem, h->d_mem));
    b.n = h->n; b.world = h->world; b.rank = h->rank; b.m_local = h->m_local;
    memcpy(out_bytes, &b, sizeof b);
    return B200QR_OK;
}

static int tsqr_dist_finish_connect(b200qr_tsqr_dist_s *h)
{
    int *fl[8] = {};
    for (int r = 0; r < h->world; ++r) fl[r] = h->flags(h->peer_mem[r]);
    CKB(cudaMemcpy(h->d_peer_flags, fl, sizeof fl, cudaMemcpyHostToDevice));
    h->connected = 1;
    return B200QR_OK;
}

int b200qr_tsqr_dist_connect(b200qr_tsqr_dist_t h, const void *all_bytes)
{
    if (!h || !all_bytes) return B200QR_EINVAL;
    const tsqr_dist_blob *blobs = (const tsqr_dist_blob *)all_bytes;
    for (int r = 0; r < h->world; ++r)
        if (blobs[r].n != h->n || blobs[r].world != h->world || blobs[r].rank != r) return B200QR_EINVAL;
    for (int r = 0; r < h->world; ++r) {
        if (r == h->rank) continue;
        void *p = nullptr;
        CKB(cudaIpcOpenMemHandle(&p, blobs[r].mem, cudaIpcMemLazyEnablePeerAccess));
        h->opened[r] = p;
        h->peer_mem[r] = (unsigned char *)p;
    }
    return tsqr_dist_finish_connect(h);
}

// rank objects that live in ONE process (tests; several GPUs driven by one process): plain pointers instead of IPC handles
int b200qr_tsqr_dist_connect_local(b200qr_tsqr_dist_t *all, int world)
{
    if (!all || world < 1 || world > 8) return B200QR_EINVAL;
    for (int r = 0; r < world; ++r) {
        if (!all[r] || all[r]->world != world || all[r]->rank != r || all[r]->n != all[0]->n) return B200QR_EINVAL;
        for (int o = 0; o < world; ++o) all[r]->peer_mem[o] = all[o]->d_mem;
        int rc = tsqr_dist_finish_connect(all[r]);
        if (rc != B200QR_OK) return rc;
    }
    return B200QR_OK;
}

// split form (a single process driving several rank objects must run phase 1 of every rank before any phase 2):
// phase 1 = local reduction + exchange + signal, phase 2 = wait + top reduction.  b200qr_tsqr_dist_factor = both.
int b200qr_tsqr_dist_factor_begin(b200qr_tsqr_dist_t h, float *dA, long long ldm, void *stream_)
{
    if (!h || !h->connected || !dA) return B200QR_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int n = h->n, W = h->world;
    int rc = b200qr_tsqr_factor(h->local, dA, ldm, h->d_rlocal, stream);
    if (rc != B200QR_OK || W == 1) return rc;
    const int e = ++h->epoch;
    for (int r = 0; r < W; ++r)
        CKB(cudaMemcpy2DAsync(h->stack(h->peer_mem[r], e & 1) + (size_t)h->rank * n, sizeof(float) * (size_t)W * n, h->d_rlocal,
                              sizeof(float) * n, sizeof(float) * n, n, cudaMemcpyDeviceToDevice, stream));
    tsqr_signal_kernel<<<1, 32, 0, stream>>>(h->d_peer_flags, W, h->rank, e);
    CKB(cudaGetLastError());
    return B200QR_OK;
}

int b200qr_tsqr_dist_factor_end(b200qr_tsqr_dist_t h, float *dR, void *stream_)
{
    if (!h || !h->connected || !dR) return B200QR_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int n = h->n, W = h->world;
    if (W == 1) {
        CKB(cudaMemcpyAsync(dR, h->d_rlocal, sizeof(float) * (size_t)n * n, cudaMemcpyDeviceToDevice, stream));
        return B200QR_OK;
    }
    const char *wd = getenv("QRB200_WATCHDOG_MS");
    tsqr_wait_kernel<<<1, 32, 0, stream>>>(h->flags(h->d_mem), W, h->epoch, (unsigned long long)(wd ? atol(wd) : 60000) * 1000000ull);
    CKB(cudaGetLastError());
    return b200qr_tsqr_factor(h->top, h->stack(h->d_mem, h->epoch & 1), (long long)W * n, dR, stream);
}

int b200qr_tsqr_dist_factor(b200qr_tsqr_dist_t h, float *dA, long long ldm, float *dR, void *stream)
{
    int rc = b200qr_tsqr_dist_factor_begin(h, dA, ldm, stream);
    return rc != B200QR_OK ? rc : b200qr_tsqr_dist_factor_end(h, dR, stream);
}

const float *b200qr_tsqr_dist_local_r(b200qr_tsqr_dist_t h) { return h ? h->d_rlocal : nullptr; }

}  // extern "C"
