// qr_persistent_tc2.cuh -- second tensor-core sweep engine (QRB200_TC=2), included by qr_persistent.cu.
//
// Same mathematics as qr_persistent_tc.cuh (compact-WY form of the reference's TSMQR / UNMQR,
// qrdecomp.c:620-644, 736-762, every product as 3xTF32 on tcgen05 with fp32 accumulation in TMEM),
// restructured around what the first engine measured: a tile row of a sweep is a CHAIN (the
// resident tile row k changes with every block of reflectors), and every CTA-wide hand-over on that
// chain costs about a microsecond.  Here nothing on the chain leaves the thread:
//     z   = V^T c_bot            tcgen05.mma, issued one row ahead                    (off the chain)
//     w_c = tau_c (c_top_c + z_c - sum_{j<c} (v_j'v_c) w_j),  c_top -= w    per thread, fp32 (the chain)
//     P   = W V^T                tcgen05.mma;  c_bot -= p one row later               (off the chain)
// The A operands of both products (c_bot and W, split hi/lo) live in TENSOR MEMORY, written by the
// thread that owns the column (tcgen05.st) -- no shared-memory operand buffers, so a sweep engine
// needs only ~111 KB and half of the tensor memory.  A kernel that uses tcgen05 gets one CTA per SM
// (cudaOccupancyMaxActiveBlocksPerMultiprocessor reports 1 whatever its resources), so a CTA holds
// TWO independent engines (warps 0-7 and 8-15), each working on its own sweep ticket: while one
// waits for a product or a tile, the other computes.
// An engine's sweep is 4 tile columns = one 128-column block; 4 consumer warps (thread = matrix
// column), a helper warp (splits V and V^T, fetches V^T V from the tensor core), loader, storer, MMA
// issuer.  Panel chains run in CTAs of their own, four panel warps each (one per sub-partition).
#pragma once
#include "qr_umma.cuh"

namespace qrb200 {

template <>
struct Cfg<4> {
    static constexpr int G = 4;
    static constexpr int NS = 3;
    static constexpr int STAGE_BYTES = G * TILE_BYTES + TILE_BYTES + 1024;      // A tiles, V (hi in place), tau
    static constexpr int OFF_VX = NS * STAGE_BYTES;                             // per stage: v_lo, vt_hi, vt_lo, G (fp32)
    static constexpr int VX_BYTES = 4 * TILE_BYTES;
    static constexpr int SWEEP_END = OFF_VX + NS * VX_BYTES;
    static constexpr int ENGINE_BYTES = SWEEP_END + 1024;                       // + barriers / ticket words; a multiple of 1024 keeps engine 1's tiles swizzle-aligned
    static constexpr int PANEL_BYTES = 8 * TILE_BYTES + 5120;
    static constexpr int OFF_PANEL = 0;                                         // panel CTAs never run sweeps
    static constexpr int OFF_BAR = 2 * ENGINE_BYTES;                            // panel warps' barriers (run_panel)
    static constexpr int SMEM_BYTES = OFF_BAR + 256;
    static constexpr bool BUILD_T = false;
};

namespace tc2 {
using C = Cfg<4>;
constexpr int NCW = 4;                                   // consumer warps
constexpr int W_HELPER = 4, W_LOAD = 5, W_STORE = 6, W_MMA = 7;
constexpr int NPW4 = 4;                                  // panel warps of a panel CTA
constexpr int ENGINE_THREADS = 256;
constexpr int NTHREADS_TC2 = 2 * ENGINE_THREADS;
constexpr uint32_t TMEM_COLS = 512, ENGINE_TMEM_COLS = 256;
__device__ __forceinline__ unsigned char *eng_smem() { return g_smem + (threadIdx.x >> 8) * C::ENGINE_BYTES; }
// TMEM columns: two parity sets of (operand hi, operand lo, Z/P), then V^T V per parity
__device__ __forceinline__ uint32_t tm_ah(int par) { return (uint32_t)(par * 96); }
__device__ __forceinline__ uint32_t tm_al(int par) { return (uint32_t)(par * 96 + 32); }
__device__ __forceinline__ uint32_t tm_zp(int par) { return (uint32_t)(par * 96 + 64); }
__device__ __forceinline__ uint32_t tm_g(int) { return 192u; }          // one buffer: the helper reads it before it stages the next row
__device__ __forceinline__ uint32_t tm_ctop() { return 224u; }           // the resident tile row k, one column per thread

struct Bars {   // per engine
    static constexpr int FULL = 0, COMPUTED = 3, EMPTY = 6, GREADY = 9, SPLIT = 12, ZREADY = 14, WSTORED = 16, PREADY = 18,
                         COUNT = 20;
};
__device__ __forceinline__ uint64_t *bar(int idx) { return reinterpret_cast<uint64_t *>(eng_smem() + C::SWEEP_END) + idx; }
__device__ __forceinline__ float *a_tile(int s, int t) { return reinterpret_cast<float *>(eng_smem() + s * C::STAGE_BYTES + t * TILE_BYTES); }
__device__ __forceinline__ float *v_hi(int s) { return a_tile(s, C::G); }
__device__ __forceinline__ float *tau_s(int s) { return a_tile(s, C::G + 1); }
__device__ __forceinline__ float *vx(int s, int which) { return reinterpret_cast<float *>(eng_smem() + C::OFF_VX + s * C::VX_BYTES + which * TILE_BYTES); }
enum { V_LO = 0, VT_HI = 1, VT_LO = 2, G_F32 = 3 };

// Round / row counters of a role.  n: rounds (reflector rows + one write-back per sweep) -> ring stage n % 3;
// r: reflector rows -> TMEM parity r & 1; u[s]: reflector rounds seen on stage s (GREADY parity).
struct Ctr {
    int n = 0, r = 0;
    int u[3] = {0, 0, 0};
    __device__ __forceinline__ int stage() const { return n % 3; }
    __device__ __forceinline__ uint32_t ring_par() const { return (uint32_t)(n / 3) & 1; }
    __device__ __forceinline__ int par() const { return r & 1; }
    __device__ __forceinline__ uint32_t tm_par() const { return (uint32_t)(r >> 1) & 1; }
    __device__ __forceinline__ void reflector_done() { ++u[n % 3]; ++r; ++n; }
    __device__ __forceinline__ void writeback_done() { ++n; }
};

// x -> tensor memory as the hi / lo halves of a 3xTF32 A operand (this thread's lane, 32 columns each)
__device__ __forceinline__ void tmem_st_split(uint32_t t_hi, uint32_t t_lo, const float (&x)[TB])
{
    float t[TB];
#pragma unroll
    for (int i = 0; i < TB; ++i) t[i] = tf32_hi(x[i]);
    tmem_st_32x32(t_hi, t);
#pragma unroll
    for (int i = 0; i < TB; ++i) t[i] = x[i] - tf32_hi(x[i]);
    tmem_st_32x32(t_lo, t);
}

// ------------------------------------------------------------------ loader / storer (ring of 3 stages)
__device__ void loader(const Params &P, const CUtensorMap *mapA, const CUtensorMap *mapVs, int k, int g, Ctr &rc, const Watch &wd)
{
    const int lane = threadIdx.x & 31;
    const int p = P.p, nrows = p - k;
    const int jt = g * C::G + lane;
    const int kla = k + (lookahead(P, k, C::G) ? 1 : 0);
    const bool active = lane < C::G && jt > kla && jt < P.q;
    const unsigned amask = __ballot_sync(FULL, active);
    const uint32_t bytes = __popc(amask) * TILE_BYTES + TILE_BYTES + 128;
    const int owner = (k / C::G) % P.world;
    const bool remote = owner != P.rank;
    const CUtensorMap *mapV = mapVs + owner;
    const float *tau_src = P.peer_tauc[owner] + (size_t)k * p * TB;
    const int *flag = active ? P.prog + (size_t)jt * p : (lane == C::G ? P.peer_prog[owner] + (size_t)k * p : nullptr);
    const bool flag_remote = remote && lane == C::G;
    const int target = active ? k : k + 1;
    for (int idx = 0; idx <= nrows; ++idx, ++rc.n) {
        const int s = rc.stage();
        mbar_wait_wd(bar(Bars::EMPTY + s), rc.ring_par() ^ 1, wd);
        if (idx < nrows) {
            const int row = k + idx;
            unsigned spins = 0;
            for (;;) {
                bool ok = flag ? (flag_remote ? ld_relaxed_sys(flag + row) : ld_relaxed(flag + row)) >= target : true;
                if (__all_sync(FULL, ok)) break;
                __nanosleep(64);
                wd.check(spins);
            }
            fence_acq_rel(remote);
            fence_proxy_async();
            if (lane == 0) mbar_arrive_expect_tx(bar(Bars::FULL + s), bytes);
            __syncwarp();
            if (active) tma_load_2d(a_tile(s, lane), mapA, row * TB, jt * TB, bar(Bars::FULL + s));
            if (lane == C::G) tma_load_2d(v_hi(s), mapV, row * TB, k * TB, bar(Bars::FULL + s));
            if (lane == C::G + 1) bulk_g2s(tau_s(s), tau_src + (size_t)row * TB, 128, bar(Bars::FULL + s));
        } else if (lane == 0) {
            mbar_arrive(bar(Bars::FULL + s));
        }
    }
}

__device__ void storer(const Params &P, const CUtensorMap *mapA, int k, int g, Ctr &rc, const Watch &wd)
{
    const int lane = threadIdx.x & 31;
    const int p = P.p, nrows = p - k;
    const int jt = g * C::G + lane;
    const int kla = k + (lookahead(P, k, C::G) ? 1 : 0);
    const bool active = lane < C::G && jt > kla && jt < P.q;
    for (int idx = 0; idx <= nrows; ++idx, ++rc.n) {
        const int s = rc.stage();
        mbar_wait_wd(bar(Bars::COMPUTED + s), rc.ring_par(), wd);
        if (idx > 0) {
            const int row = idx < nrows ? k + idx : k;
            if (active) {
                tma_store_2d(mapA, row * TB, jt * TB, a_tile(s, lane));
                bulk_commit();
                bulk_wait_all();
                __threadfence();
                st_release(P.prog + (size_t)jt * p + row, k + 1);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(Bars::EMPTY + s));
    }
}

// ------------------------------------------------------------------ helper: V operands and V^T V
__device__ void helper(const Params &P, int k, uint32_t tmem, Ctr &rc, const Watch &wd)
{
    const int lane = threadIdx.x & 31;
    const int nrows = P.p - k;
    for (int idx = 0; idx < nrows; ++idx) {
        const int s = rc.stage(), par = rc.par();
        mbar_wait_wd(bar(Bars::FULL + s), rc.ring_par(), wd);
        {   // reflector column `lane`: hi in place, lo, and the transposed copies
            float v[TB];
            load_col_sw(v, v_hi(s), lane);
            if (idx == 0) {
#pragma unroll
                for (int r = 0; r < TB; ++r) v[r] = (r > lane) ? v[r] : ((r == lane) ? 1.0f : 0.f);   // unit lower V1
            }
            float *vth = vx(s, VT_HI), *vtl = vx(s, VT_LO);
#pragma unroll
            for (int r = 0; r < TB; ++r) {
                const float h = tf32_hi(v[r]);
                const int off = r * 32 + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3));
                vth[off] = h;
                vtl[off] = v[r] - h;
            }
            tc::store_line_split(v_hi(s), vx(s, V_LO), lane, v);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar(Bars::SPLIT + par));
        // V^T V comes back from the tensor core together with Z of this row
        mbar_wait_wd(bar(Bars::ZREADY + par), rc.tm_par(), wd);
        tc_fence_after();
        float grow[TB];
        tmem_ld_32x32(grow, tmem + tm_g(par));                       // lane j: (V^T V)[j][0..31]
        const float tj = tau_s(s)[lane];
#pragma unroll
        for (int c = 0; c < TB; ++c) grow[c] *= tj;                  // tau_j folded in: the chain below is FMAs only
        store_col_smem(vx(s, G_F32) + lane * TB, grow);             // row-major: consumers read row j as a broadcast
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(Bars::GREADY + s));
        rc.reflector_done();
    }
    rc.writeback_done();
}

// ------------------------------------------------------------------ MMA issuer
// 3xTF32: D = Ah*Bh + Ah*Bl + Al*Bh, four K = 8 steps each.  Operand descriptors are passed as their low
// words (umma_desc_lo): the issue loop is 12 x (two adds + one tcgen05.mma).
__device__ __forceinline__ void mma3_ss(uint32_t d, uint32_t a_h, uint32_t a_l, uint32_t b_h, uint32_t b_l, uint32_t idesc)
{
    const uint32_t As[3] = {a_h, a_h, a_l}, Bs[3] = {b_h, b_l, b_h};
    bool acc = false;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            umma_tf32(d, umma_desc_from(As[pass] + 2 * kk), umma_desc_from(Bs[pass] + 2 * kk), idesc, acc);
            acc = true;
        }
}
__device__ __forceinline__ void mma3_ts(uint32_t d, uint32_t a_h, uint32_t a_l, uint32_t b_h, uint32_t b_l, uint32_t idesc)
{
    const uint32_t As[3] = {a_h, a_h, a_l}, Bs[3] = {b_h, b_l, b_h};
    bool acc = false;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            umma_tf32_ts(d, As[pass] + 8 * kk, umma_desc_from(Bs[pass] + 2 * kk), idesc, acc);
            acc = true;
        }
}

__device__ void mma_issuer(const Params &P, int k, uint32_t tmem, Ctr &rc, const Watch &wd)
{
    const int nrows = P.p - k;
    const bool me = (threadIdx.x & 31) == 0;
    const uint32_t idesc = umma_idesc_tf32(128, 32, false, false);
    auto issue_gz = [&](int s, int par, uint32_t tpar) {     // V^T V and Z of the row staged in s
        const uint32_t vh = umma_desc_lo(v_hi(s)), vl = umma_desc_lo(vx(s, V_LO));
        mbar_wait_wd(bar(Bars::SPLIT + par), tpar, wd);
        tc_fence_after();
        if (me) {
            mma3_ts(tmem + tm_zp(par), tmem + tm_ah(par), tmem + tm_al(par), vh, vl, idesc);     // Z first: the chain waits for it
            mma3_ss(tmem + tm_g(par), vh, vl, vh, vl, idesc);                                     // rows >= 32 of D unused
            umma_commit(bar(Bars::ZREADY + par));
        }
        __syncwarp();
    };
    issue_gz(rc.stage(), rc.par(), rc.tm_par());
    for (int idx = 0; idx < nrows; ++idx) {
        const int s = rc.stage(), par = rc.par();
        const uint32_t th = umma_desc_lo(vx(s, VT_HI)), tl = umma_desc_lo(vx(s, VT_LO));
        // the next row's first products are staged a chain step ahead (same order as the consumers' arrivals)
        if (idx + 1 < nrows) issue_gz((rc.n + 1) % 3, par ^ 1, (uint32_t)((rc.r + 1) >> 1) & 1);
        // P = W V^T (W in the operand columns of this parity, written by the consumers)
        mbar_wait_wd(bar(Bars::WSTORED + par), rc.tm_par(), wd);
        tc_fence_after();
        if (me) {
            mma3_ts(tmem + tm_zp(par), tmem + tm_ah(par), tmem + tm_al(par), th, tl, idesc);
            umma_commit(bar(Bars::PREADY + par));
        }
        __syncwarp();
        rc.reflector_done();
    }
    rc.writeback_done();
}

// ------------------------------------------------------------------ consumer: one matrix column per thread
__device__ void consumer(const Params &P, int k, int g, uint32_t tmem, Ctr &rc, const Watch &wd)
{
    const int w = (threadIdx.x >> 5) & 7, lane = threadIdx.x & 31;
    const int nrows = P.p - k;
    const int jt = g * C::G + w;
    const bool act = jt > k + (lookahead(P, k, C::G) ? 1 : 0) && jt < P.q;
    const uint32_t lane_base = (uint32_t)(w * 32) << 16;
    // optional phase timers (opts.trace == 3): nanoseconds waited per barrier kind, summed over the run by
    // lane 0 of consumer warp 0 of every engine into ctr[32..]: FULL, ZREADY, GREADY, PREADY, total, rows
    const bool prof = P.trace_mode == 3 && w == 0 && lane == 0;
    unsigned long long t_full = 0, t_z = 0, t_g = 0, t_p = 0, t_begin = prof ? gtimer() : 0;
#define TIMED(acc, stmt) do { unsigned long long t0_ = prof ? gtimer() : 0; stmt; if (prof) acc += gtimer() - t0_; } while (0)

    auto split_row = [&](int s, uint32_t ring_par, int par) {    // operands of Z = V^T c_bot into tensor memory
        TIMED(t_full, mbar_wait_wd(bar(Bars::FULL + s), ring_par, wd));
        if (act) {
            float x[TB];
            load_col_sw(x, a_tile(s, w), lane);
            tmem_st_split(tmem + lane_base + tm_ah(par), tmem + lane_base + tm_al(par), x);
        }
        tc_fence_before();
        mbar_arrive(bar(Bars::SPLIT + par));
    };
    // finish row `prev`: c_bot -= p (or, for the UNMQR row, the result becomes the resident tile row)
    auto finish_row = [&](int s_prev, int par_prev, uint32_t tpar_prev, bool prev_is_diag) {
        TIMED(t_p, mbar_wait_wd(bar(Bars::PREADY + par_prev), tpar_prev, wd));
        tc_fence_after();
        if (act) {
            float pr[TB], x[TB];
            tmem_ld_32x32(pr, tmem + lane_base + tm_zp(par_prev));
            load_col_sw(x, a_tile(s_prev, w), lane);
            if (prev_is_diag) {
#pragma unroll
                for (int r = 0; r < TB; ++r) x[r] -= pr[r];
                tmem_st_32x32(tmem + lane_base + tm_ctop(), x);      // the UNMQR result is the resident tile row
            } else {
#pragma unroll
                for (int r = 0; r < TB; ++r) x[r] -= pr[r];
                store_col_sw(a_tile(s_prev, w), lane, x);
            }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(bar(Bars::COMPUTED + s_prev));
    };

    split_row(rc.stage(), rc.ring_par(), rc.par());
    bool have_prev = false;
    int s_prev = 0, par_prev = 0;
    uint32_t tpar_prev = 0;
    for (int idx = 0; idx < nrows; ++idx) {
        const int s = rc.stage(), par = rc.par();
        const uint32_t tpar = rc.tm_par();
        // previous row: subtract its product (for the UNMQR row the result becomes the resident tile row),
        // release its stage; then stage the next row's operands -- a whole chain step ahead of their use
        if (have_prev) finish_row(s_prev, par_prev, tpar_prev, idx == 1);
        if (idx + 1 < nrows) split_row((rc.n + 1) % 3, (uint32_t)((rc.n + 1) / 3) & 1, par ^ 1);
        // z = V^T c_bot
        TIMED(t_z, mbar_wait_wd(bar(Bars::ZREADY + par), tpar, wd));
        tc_fence_after();
        float sv[TB];
        if (act) {
            tmem_ld_32x32(sv, tmem + lane_base + tm_zp(par));
            if (idx > 0) {
                float ct[TB];
                tmem_ld_32x32(ct, tmem + lane_base + tm_ctop());
#pragma unroll
                for (int r = 0; r < TB; ++r) sv[r] += ct[r];
            }
        }
        // w_c = tau_c (s_c - sum_{j<c} G[j][c] w_j): the reference's reflector-by-reflector recurrence
        TIMED(t_g, mbar_wait_wd(bar(Bars::GREADY + s), rc.u[s] & 1, wd));
        if (act) {
            // sv_c -= (tau_j G[j][c]) sv_j for j < c, in order; then w_j = tau_j sv_j
            const float *gm = vx(s, G_F32);
#pragma unroll
            for (int j = 0; j < TB - 1; ++j) {
                const float sj = sv[j];
#pragma unroll
                for (int c4 = (j + 1) & ~3; c4 < TB; c4 += 4) {
                    const float4 g4 = *reinterpret_cast<const float4 *>(gm + j * TB + c4);
                    if (c4 + 0 > j) sv[c4 + 0] = fmaf(-g4.x, sj, sv[c4 + 0]);
                    if (c4 + 1 > j) sv[c4 + 1] = fmaf(-g4.y, sj, sv[c4 + 1]);
                    if (c4 + 2 > j) sv[c4 + 2] = fmaf(-g4.z, sj, sv[c4 + 2]);
                    if (c4 + 3 > j) sv[c4 + 3] = fmaf(-g4.w, sj, sv[c4 + 3]);
                }
            }
            {
                float tv[TB];
                load_col_smem(tv, tau_s(s));
#pragma unroll
                for (int r = 0; r < TB; ++r) sv[r] *= tv[r];
            }
            if (idx > 0) {
                float ct[TB];
                tmem_ld_32x32(ct, tmem + lane_base + tm_ctop());
#pragma unroll
                for (int r = 0; r < TB; ++r) ct[r] -= sv[r];
                tmem_st_32x32(tmem + lane_base + tm_ctop(), ct);
            }
            // W becomes the A operand of P = W V^T (the c_bot operands of this parity are dead: Z is complete)
            tmem_st_split(tmem + lane_base + tm_ah(par), tmem + lane_base + tm_al(par), sv);
        }
        tc_fence_before();
        mbar_arrive(bar(Bars::WSTORED + par));
        have_prev = true; s_prev = s; par_prev = par; tpar_prev = tpar;
        rc.reflector_done();
    }
    if (have_prev) finish_row(s_prev, par_prev, tpar_prev, nrows == 1);
    // write-back round: the resident tile row k goes out through the next stage
    {
        const int s = rc.stage();
        mbar_wait_wd(bar(Bars::FULL + s), rc.ring_par(), wd);
        if (act) {
            float ct[TB];
            tmem_ld_32x32(ct, tmem + lane_base + tm_ctop());
            store_col_sw(a_tile(s, w), lane, ct);
        }
        fence_proxy_async_smem();
        mbar_arrive(bar(Bars::COMPUTED + s));
        rc.writeback_done();
    }
    if (prof) {
        unsigned long long *dbg = reinterpret_cast<unsigned long long *>(P.ctr + 32);
        atomicAdd(dbg + 0, t_full); atomicAdd(dbg + 1, t_z); atomicAdd(dbg + 2, t_g); atomicAdd(dbg + 3, t_p);
        atomicAdd(dbg + 4, gtimer() - t_begin); atomicAdd(dbg + 5, (unsigned long long)nrows);
    }
#undef TIMED
}

}  // namespace tc2

__global__ void __launch_bounds__(tc2::NTHREADS_TC2, 1) tiledqr_persistent_tc2(const __grid_constant__ KernelArgs KA)
{
    using namespace tc2;
    const Params &P = KA.rc[0].P;
    const CUtensorMap &mapA = KA.rc[0].mapA;
    const CUtensorMap *mapVs = KA.rc[0].mapV;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int eng = warp >> 3, lw = warp & 7;                     // engine, warp inside the engine
    const Watch wd{gtimer(), P.watchdog_ns};
    const bool panel_cta = (int)blockIdx.x < P.npanel_ctas;
    int *s_task = reinterpret_cast<int *>(bar(Bars::COUNT));
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(g_smem + C::OFF_BAR + 192);

    if ((threadIdx.x & (ENGINE_THREADS - 1)) == 0) {             // one thread per engine
        for (int s = 0; s < 3; ++s) {
            mbar_init(bar(Bars::FULL + s), 1);
            mbar_init(bar(Bars::COMPUTED + s), NCW * 32);
            mbar_init(bar(Bars::EMPTY + s), 1);
            mbar_init(bar(Bars::GREADY + s), 1);
        }
        for (int par = 0; par < 2; ++par) {
            mbar_init(bar(Bars::SPLIT + par), NCW * 32 + 32);
            mbar_init(bar(Bars::ZREADY + par), 1);
            mbar_init(bar(Bars::WSTORED + par), NCW * 32);
            mbar_init(bar(Bars::PREADY + par), 1);
        }
        if (threadIdx.x == 0) {
            uint64_t *pf = reinterpret_cast<uint64_t *>(g_smem + C::OFF_BAR) + 3 * C::NS;
            for (int i = 0; i < 2 * NPW4; ++i) mbar_init(pf + i, 1);
            tma_prefetch_desc(&mapA);
            for (int r = 0; r < P.world; ++r) tma_prefetch_desc(&mapVs[r]);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && !panel_cta) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (panel_cta) {
        if (warp >= NPW4) return;
        uint32_t phbits = 0;
        for (;;) {
            int t = 0;
            if (lane == 0) t = atomicAdd(&P.ctr[2], 1);
            t = __shfl_sync(FULL, t, 0);
            if (t >= P.npanels) break;
            run_panel<4>(P, &mapA, P.panel_tab[t], warp, phbits, wd);
        }
        return;
    }

    const uint32_t tmem = *tmem_slot + (uint32_t)eng * ENGINE_TMEM_COLS;
    Ctr rc;
    if (P.trace_mode == 7 && eng == 1) goto done;                // debugging aid: one engine only
    for (;;) {
        named_bar_sync(1 + eng, ENGINE_THREADS);
        if (lw == W_LOAD && lane == 0) {
            int t = atomicAdd(&P.ctr[0], 1);
            int2 kg = t < P.nsweeps ? P.sweep_tab[t] : make_int2(-1, 0);
            s_task[0] = kg.x; s_task[1] = kg.y;
        }
        named_bar_sync(1 + eng, ENGINE_THREADS);
        const int k = s_task[0], g = s_task[1];
        if (k < 0) break;
        {
            const int j_lo = max(g * C::G, k + 1 + (lookahead(P, k, C::G) ? 1 : 0));
            const int j_hi = min((g + 1) * C::G, P.q);
            if (j_lo >= j_hi) continue;
        }
        if (lw == W_LOAD) loader(P, &mapA, mapVs, k, g, rc, wd);
        else if (lw == W_STORE) storer(P, &mapA, k, g, rc, wd);
        else if (lw == W_HELPER) helper(P, k, tmem, rc, wd);
        else if (lw == W_MMA) mma_issuer(P, k, tmem, rc, wd);
        else consumer(P, k, g, tmem, rc, wd);
    }
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TMEM_COLS>(*tmem_slot);
}

}  // namespace qrb200
