// qr_tiles.cu -- the tiled QR at tile sizes 32 / 64 / 128 (BASELINE.json configs[1]: "8192 x 8192 ... tile-size sweep").
//
// The reference bakes tile size 32 into both paths (gpucalc.cu:1582 `b = 32`; qrdecomp.c:506 literal 32), and the
// drop-in output layout is defined on 32 x 32 tiles, so the persistent engines of qr_persistent.cu are 32-only.  This
// file is the documented NON-DEFAULT mode for other tile sizes: the same four tile kernels with the tile size B a
// template parameter -- GEQRT (qrdecomp.c:532-557), TSQRT (689-721), UNMQR (620-644), TSMQR (736-762), same operation
// sequence and sign conventions -- in the reference's flat-tree order, host-ordered like cudaQRFull (one panel launch and
// two update launches per step, the update of the next panel's column on the panel stream so that panel k + 1 overlaps
// the rest of update k).  Output layout = the reference's with b -> B: R in the global upper triangle, GEQRT reflectors
// below the diagonal of diagonal tiles, a full B x B block V2 in every tile below; tau compact [(k*p + i)*B + r].
//
// Mapping: a CTA of B threads owns one tile column of width B, thread = matrix column, its B rows of the current tile in
// registers; the reflector of the step is broadcast through shared memory, one CTA barrier per reflector.  What it costs
// (bench.py sweep, 8192^2): the chain per tile grows with B (B reflectors of B + 1 entries, strictly serial) while the
// number of independent tile columns shrinks (8192 / 128 = 64 CTAs for 148 SMs), which is why 32 is the size the fast
// engines are built around.
#include "qr_device.cuh"
#include "../../include/qrb200.h"

#include <cstdint>
#include <cstdlib>

namespace qrb200 {

template <int B>
__device__ __forceinline__ void ld_col(float (&x)[B], const float *col)
{
#pragma unroll
    for (int r = 0; r < B; r += 4) {
        const float4 t = __ldcg(reinterpret_cast<const float4 *>(col + r));
        x[r] = t.x; x[r + 1] = t.y; x[r + 2] = t.z; x[r + 3] = t.w;
    }
}
template <int B>
__device__ __forceinline__ void st_col(float *col, const float (&x)[B])
{
#pragma unroll
    for (int r = 0; r < B; r += 4) *reinterpret_cast<float4 *>(col + r) = make_float4(x[r], x[r + 1], x[r + 2], x[r + 3]);
}
template <int B>
__device__ __forceinline__ float sumsq_col(const float (&x)[B])
{
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int r = 0; r < B; r += 4) {
        s0 = fmaf(x[r], x[r], s0); s1 = fmaf(x[r + 1], x[r + 1], s1);
        s2 = fmaf(x[r + 2], x[r + 2], s2); s3 = fmaf(x[r + 3], x[r + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}
// dot of the register column with a shared-memory vector (128-bit broadcast loads)
template <int B>
__device__ __forceinline__ float dot_sm(const float (&b)[B], const float *x)
{
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int r = 0; r < B; r += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(x + r);
        s0 = fmaf(b[r], v.x, s0); s1 = fmaf(b[r + 1], v.y, s1); s2 = fmaf(b[r + 2], v.z, s2); s3 = fmaf(b[r + 3], v.w, s3);
    }
    return (s0 + s1) + (s2 + s3);
}
template <int B>
__device__ __forceinline__ void axpy_sm(float (&b)[B], float g, const float *x)
{
#pragma unroll
    for (int r = 0; r < B; r += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(x + r);
        b[r] = fmaf(g, v.x, b[r]); b[r + 1] = fmaf(g, v.y, b[r + 1]); b[r + 2] = fmaf(g, v.z, b[r + 2]); b[r + 3] = fmaf(g, v.w, b[r + 3]);
    }
}

// Shared memory of the panel kernel: R (B rows of B + 1 floats), two broadcast slots, two {1/piv, tau} slots.
template <int B> constexpr int panel_smem_floats() { return B * (B + 1) + 2 * B + 8; }

// P(k): GEQRT of tile (k, k), then the TSQRT chain down tile column k.  One CTA of B threads.
template <int B>
__global__ void __launch_bounds__(B) tiles_panel_kernel(float *A, int ldm, int p, int k, float *tauc)
{
    extern __shared__ __align__(16) float tsm[];
    float *Rs = tsm;                                        // Rs[r * (B + 1) + j]
    float *xs = tsm + B * (B + 1);                          // two slots of B floats (16-byte aligned: B*(B+1) % 4 == 0)
    float *hs = xs + 2 * B;
    const int j = threadIdx.x;
    float *colk = A + (size_t)(k * B + j) * ldm;
    float a[B];
    ld_col<B>(a, colk + (size_t)k * B);
    // ---- GEQRT (qrdecomp.c:532-557, 1201-1227, 489-530): reflector c acts on rows c.. of columns c..
#pragma unroll 1
    for (int c = 0; c < B; ++c) {
        float *x = xs + (c & 1) * B;
        float *h2 = hs + (c & 1) * 2;
        if (j == c) {
            st_col<B>(x, a);
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int r = 0; r < B; r += 2) { s0 = r > c ? fmaf(a[r], a[r], s0) : s0; s1 = r + 1 > c ? fmaf(a[r + 1], a[r + 1], s1) : s1; }
            float x0 = 0.f;
#pragma unroll
            for (int r = 0; r < B; ++r) x0 = r == c ? a[r] : x0;
            const HouseScalars h = house_scalars(x0, s0 + s1);
            h2[0] = h.inv; h2[1] = h.tau;
            tauc[((size_t)k * p + k) * B + c] = h.tau;
        }
        __syncthreads();
        const float inv = h2[0], tau = h2[1];
        if (j >= c) {
            // f = tau (a_c + (1 / piv) sum_{r > c} a_r x_r);  a_c -= f;  a_r -= f x_r / piv;  owner keeps v = x / piv
            float d0 = 0.f, d1 = 0.f, ac = 0.f;
#pragma unroll
            for (int r = 0; r < B; r += 2) {
                d0 = r > c ? fmaf(a[r], x[r], d0) : d0;
                d1 = r + 1 > c ? fmaf(a[r + 1], x[r + 1], d1) : d1;
                ac = r == c ? a[r] : (r + 1 == c ? a[r + 1] : ac);
            }
            const float f = tau * fmaf(inv, d0 + d1, ac);
            const float g = j > c ? -f * inv : 0.f;
#pragma unroll
            for (int r = 0; r < B; ++r) {
                const float upd = j == c ? x[r] * inv : fmaf(g, x[r], a[r]);
                a[r] = r > c ? upd : (r == c ? a[r] - f : a[r]);
            }
        }
    }
    st_col<B>(colk + (size_t)k * B, a);                    // R above / on the diagonal (final unless a chain follows), V below
#pragma unroll
    for (int r = 0; r < B; ++r) Rs[r * (B + 1) + j] = a[r];
    __syncthreads();
    // ---- TSQRT chain (qrdecomp.c:689-721, 1243-1272, 647-687): [R; tile (i, k)], R stays in shared memory
    for (int i = k + 1; i < p; ++i) {
        float b[B];
        ld_col<B>(b, colk + (size_t)i * B);
        float ssq = sumsq_col<B>(b);
        float my_tau = 0.f;
#pragma unroll 1
        for (int c = 0; c < B; ++c) {
            float *x = xs + (c & 1) * B;
            float *h2 = hs + (c & 1) * 2;
            if (j == c) {
                st_col<B>(x, b);
                const HouseScalars h = house_scalars(Rs[c * (B + 1) + c], ssq);
                h2[0] = h.inv; h2[1] = h.tau;
                my_tau = h.tau;
            }
            __syncthreads();
            const float inv = h2[0], tau = h2[1];
            if (j >= c) {
                const float rc = Rs[c * (B + 1) + j];
                const float beta = tau * fmaf(inv, dot_sm<B>(b, x), rc);
                Rs[c * (B + 1) + j] = rc - beta;
                if (j > c) {
                    axpy_sm<B>(b, -beta * inv, x);
                    if (j == c + 1) ssq = sumsq_col<B>(b);
                } else {
#pragma unroll
                    for (int r = 0; r < B; ++r) b[r] *= inv;               // the owner's column becomes V2(:, c)
                }
            }
        }
        st_col<B>(colk + (size_t)i * B, b);
        tauc[((size_t)k * p + i) * B + j] = my_tau;
        __syncthreads();                                    // the slots: the next tile starts at c = 0 again
    }
    // the final R of the diagonal tile
    float *rcol = colk + (size_t)k * B;
    for (int r = 0; r <= j; ++r) rcol[r] = Rs[r * (B + 1) + j];
}

// Shared memory of the update / apply kernels: the reflector tile (column-major, B x B), the resident top tile
// (top[c * (B + 1) + j]), tau.
template <int B> constexpr int update_smem_floats() { return B * B + B * (B + 1) + B; }

template <int B>
__device__ __forceinline__ void stage_tile(float *Vt, float *ts, const float *A, int ldm, int i, int k, int p, const float *tauc, int j)
{
    const float *src = A + (size_t)(k * B + j) * ldm + (size_t)i * B;
#pragma unroll 4                                            // (not the whole column at once: the caller's column is live too)
    for (int r = 0; r < B; r += 4) *reinterpret_cast<float4 *>(Vt + j * B + r) = __ldcg(reinterpret_cast<const float4 *>(src + r));
    ts[j] = tauc[((size_t)k * p + i) * B + j];
}

// A TSQRT block of B reflectors, v_c = [e_c; V2(:, c)], applied to (top[c], b[]) -- thread = column, b[] in registers.
// REV: reflectors last-to-first (Q instead of Q^T).
template <int B, bool REV>
__device__ __forceinline__ void apply_tile(float (&b)[B], float *top, const float *Vt, const float *ts, int j)
{
#pragma unroll 1
    for (int cc = 0; cc < B; ++cc) {
        const int c = REV ? B - 1 - cc : cc;
        const float *v = Vt + c * B;
        const float a0 = top[c * (B + 1) + j];
        const float beta = ts[c] * (a0 + dot_sm<B>(b, v));
        top[c * (B + 1) + j] = a0 - beta;
        axpy_sm<B>(b, -beta, v);
    }
}
// A GEQRT block (unit lower triangular V, the tile's upper part is R and is skipped) applied to this thread's column of
// the resident top tile, in shared memory: once per tile-column chain, so it runs out of shared memory and leaves the
// registers to the chain above.
template <int B, bool REV>
__device__ __forceinline__ void apply_diag_smem(float *top, const float *Vt, const float *ts, int j)
{
#pragma unroll 1
    for (int cc = 0; cc < B; ++cc) {
        const int c = REV ? B - 1 - cc : cc;
        const float *v = Vt + c * B;
        float d = top[c * (B + 1) + j];
        for (int r = c + 1; r < B; ++r) d = fmaf(top[r * (B + 1) + j], v[r], d);
        const float beta = ts[c] * d;
        top[c * (B + 1) + j] -= beta;
        for (int r = c + 1; r < B; ++r) top[r * (B + 1) + j] = fmaf(-beta, v[r], top[r * (B + 1) + j]);
    }
}
template <int B>
__device__ __forceinline__ void col_to_top(float *top, const float *col, int j)
{
#pragma unroll 4
    for (int r = 0; r < B; r += 4) {
        const float4 t = __ldcg(reinterpret_cast<const float4 *>(col + r));
        top[r * (B + 1) + j] = t.x; top[(r + 1) * (B + 1) + j] = t.y; top[(r + 2) * (B + 1) + j] = t.z; top[(r + 3) * (B + 1) + j] = t.w;
    }
}
template <int B>
__device__ __forceinline__ void top_to_col(float *col, const float *top, int j)
{
#pragma unroll 4
    for (int r = 0; r < B; r += 4)
        *reinterpret_cast<float4 *>(col + r) = make_float4(top[r * (B + 1) + j], top[(r + 1) * (B + 1) + j], top[(r + 2) * (B + 1) + j],
                                                           top[(r + 3) * (B + 1) + j]);
}

// S(k, J): UNMQR on tile (k, J), then the TSMQR chain down tile column J; one CTA of B threads per tile column,
// J = j_first + blockIdx.x, tile row k resident in shared memory.
template <int B>
__global__ void __launch_bounds__(B) tiles_update_kernel(float *A, int ldm, int p, int k, int j_first, const float *tauc)
{
    extern __shared__ __align__(16) float tsm[];
    float *Vt = tsm, *top = tsm + B * B, *ts = top + B * (B + 1);
    const int j = threadIdx.x, J = j_first + blockIdx.x;
    float *col = A + (size_t)(J * B + j) * ldm;
    stage_tile<B>(Vt, ts, A, ldm, k, k, p, tauc, j);
    col_to_top<B>(top, col + (size_t)k * B, j);
    __syncthreads();
    apply_diag_smem<B, false>(top, Vt, ts, j);
    for (int i = k + 1; i < p; ++i) {
        __syncthreads();
        stage_tile<B>(Vt, ts, A, ldm, i, k, p, tauc, j);
        __syncthreads();
        float b[B];
        ld_col<B>(b, col + (size_t)i * B);
        apply_tile<B, false>(b, top, Vt, ts, j);
        st_col<B>(col + (size_t)i * B, b);
    }
    top_to_col<B>(col + (size_t)k * B, top, j);
}

// C <- Q^T C (FWD) or Q C: one CTA of B threads per B columns of C (thread = column of C)
template <int B, bool FWD>
__global__ void __launch_bounds__(B) tiles_apply_q_kernel(const float *A, int ldm, int p, int K, const float *tauc, float *Cm, int ldc)
{
    extern __shared__ __align__(16) float tsm[];
    float *Vt = tsm, *top = tsm + B * B, *ts = top + B * (B + 1);
    const int j = threadIdx.x;
    float *col = Cm + (size_t)(blockIdx.x * B + j) * ldc;
    for (int kk = 0; kk < K; ++kk) {
        const int k = FWD ? kk : K - 1 - kk;
        __syncthreads();
        col_to_top<B>(top, col + (size_t)k * B, j);
        if (FWD) {
            stage_tile<B>(Vt, ts, A, ldm, k, k, p, tauc, j);
            __syncthreads();
            apply_diag_smem<B, false>(top, Vt, ts, j);
        }
        for (int ll = k + 1; ll < p; ++ll) {
            const int l = FWD ? ll : p - 1 - (ll - k - 1);
            __syncthreads();
            stage_tile<B>(Vt, ts, A, ldm, l, k, p, tauc, j);
            __syncthreads();
            float b[B];
            ld_col<B>(b, col + (size_t)l * B);
            apply_tile<B, !FWD>(b, top, Vt, ts, j);
            st_col<B>(col + (size_t)l * B, b);
        }
        if (!FWD) {
            __syncthreads();
            stage_tile<B>(Vt, ts, A, ldm, k, k, p, tauc, j);
            __syncthreads();
            apply_diag_smem<B, true>(top, Vt, ts, j);
        }
        top_to_col<B>(col + (size_t)k * B, top, j);
    }
}

}  // namespace qrb200

using namespace qrb200;

static thread_local int g_tiles_cuda = 0;
#define CKT(call)                                                          \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) { g_tiles_cuda = (int)e_; return B200QR_ECUDA; } \
    } while (0)

struct b200qr_tileplan_s {
    int m, n, b, p, q, K;
    float *d_tau = nullptr;
    cudaStream_t s0 = nullptr, s1 = nullptr;
    cudaEvent_t e_in = nullptr, e_s0 = nullptr, e_s1 = nullptr;
    cudaEvent_t *ep = nullptr, *eu = nullptr;       // per step: panel done, rest-of-update done
};

template <int B>
static int tiles_factor_t(b200qr_tileplan_s *pl, float *dA, int ldm, cudaStream_t stream)
{
    const int p = pl->p, q = pl->q, K = pl->K;
    const int psm = panel_smem_floats<B>() * (int)sizeof(float), usm = update_smem_floats<B>() * (int)sizeof(float);
    CKT(cudaFuncSetAttribute(tiles_panel_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm));
    CKT(cudaFuncSetAttribute(tiles_update_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, usm));
    CKT(cudaEventRecord(pl->e_in, stream));
    CKT(cudaStreamWaitEvent(pl->s0, pl->e_in, 0));
    CKT(cudaStreamWaitEvent(pl->s1, pl->e_in, 0));
    for (int k = 0; k < K; ++k) {
        // s0: P(k), then the update of the NEXT panel's column; s1: the rest of update k, overlapping P(k + 1)
        tiles_panel_kernel<B><<<1, B, psm, pl->s0>>>(dA, ldm, p, k, pl->d_tau);
        CKT(cudaEventRecord(pl->ep[k], pl->s0));
        if (k + 1 < q) {
            if (k > 0) CKT(cudaStreamWaitEvent(pl->s0, pl->eu[k - 1], 0));       // column k + 1 has seen step k - 1
            tiles_update_kernel<B><<<1, B, usm, pl->s0>>>(dA, ldm, p, k, k + 1, pl->d_tau);
        }
        if (k + 2 < q) {
            CKT(cudaStreamWaitEvent(pl->s1, pl->ep[k], 0));
            tiles_update_kernel<B><<<q - k - 2, B, usm, pl->s1>>>(dA, ldm, p, k, k + 2, pl->d_tau);
        }
        CKT(cudaEventRecord(pl->eu[k], pl->s1));
    }
    CKT(cudaGetLastError());
    CKT(cudaEventRecord(pl->e_s0, pl->s0));
    CKT(cudaEventRecord(pl->e_s1, pl->s1));
    CKT(cudaStreamWaitEvent(stream, pl->e_s0, 0));
    CKT(cudaStreamWaitEvent(stream, pl->e_s1, 0));
    return B200QR_OK;
}

template <int B>
static int tiles_apply_t(b200qr_tileplan_s *pl, const float *dA, int ldm, float *dC, int ldc, int ncols, int transpose, cudaStream_t stream)
{
    const int usm = update_smem_floats<B>() * (int)sizeof(float);
    if (transpose) {
        CKT(cudaFuncSetAttribute(tiles_apply_q_kernel<B, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, usm));
        tiles_apply_q_kernel<B, true><<<ncols / B, B, usm, stream>>>(dA, ldm, pl->p, pl->K, pl->d_tau, dC, ldc);
    } else {
        CKT(cudaFuncSetAttribute(tiles_apply_q_kernel<B, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, usm));
        tiles_apply_q_kernel<B, false><<<ncols / B, B, usm, stream>>>(dA, ldm, pl->p, pl->K, pl->d_tau, dC, ldc);
    }
    CKT(cudaGetLastError());
    return B200QR_OK;
}

extern "C" {

int b200qr_tiles_last_cuda_error(void) { return g_tiles_cuda; }

int b200qr_tiles_create(b200qr_tileplan_t *out, int m, int n, int b)
{
    if (!out || (b != 32 && b != 64 && b != 128) || m <= 0 || n <= 0 || m % b || n % b) return B200QR_EINVAL;
    b200qr_tileplan_s *pl = new b200qr_tileplan_s;
    pl->m = m; pl->n = n; pl->b = b; pl->p = m / b; pl->q = n / b; pl->K = pl->p < pl->q ? pl->p : pl->q;
    bool ok = cudaMalloc(&pl->d_tau, sizeof(float) * (size_t)pl->K * pl->p * b) == cudaSuccess &&
              cudaMemset(pl->d_tau, 0, sizeof(float) * (size_t)pl->K * pl->p * b) == cudaSuccess &&
              cudaStreamCreateWithFlags(&pl->s0, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&pl->s1, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&pl->e_in, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&pl->e_s0, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&pl->e_s1, cudaEventDisableTiming) == cudaSuccess;
    if (ok) {
        pl->ep = new cudaEvent_t[pl->K]();
        pl->eu = new cudaEvent_t[pl->K]();
        for (int k = 0; k < pl->K && ok; ++k)
            ok = cudaEventCreateWithFlags(&pl->ep[k], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&pl->eu[k], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) { (void)cudaGetLastError(); b200qr_tiles_destroy(pl); return B200QR_ENOMEM; }
    *out = pl;
    return B200QR_OK;
}

int b200qr_tiles_destroy(b200qr_tileplan_t pl)
{
    if (!pl) return B200QR_EINVAL;
    cudaFree(pl->d_tau);
    if (pl->ep) for (int k = 0; k < pl->K; ++k) if (pl->ep[k]) cudaEventDestroy(pl->ep[k]);
    if (pl->eu) for (int k = 0; k < pl->K; ++k) if (pl->eu[k]) cudaEventDestroy(pl->eu[k]);
    delete[] pl->ep; delete[] pl->eu;
    if (pl->e_in) cudaEventDestroy(pl->e_in);
    if (pl->e_s0) cudaEventDestroy(pl->e_s0);
    if (pl->e_s1) cudaEventDestroy(pl->e_s1);
    if (pl->s0) cudaStreamDestroy(pl->s0);
    if (pl->s1) cudaStreamDestroy(pl->s1);
    delete pl;
    return B200QR_OK;
}

int b200qr_tiles_tile(b200qr_tileplan_t pl) { return pl ? pl->b : B200QR_EINVAL; }
const float *b200qr_tiles_tau(b200qr_tileplan_t pl) { return pl ? pl->d_tau : nullptr; }

int b200qr_tiles_factor(b200qr_tileplan_t pl, float *dA, int ldm, void *stream)
{
    if (!pl || !dA || ldm < pl->m || (ldm & 3) || ((uintptr_t)dA & 15)) return B200QR_EINVAL;
    if (pl->b == 32) return tiles_factor_t<32>(pl, dA, ldm, (cudaStream_t)stream);
    if (pl->b == 64) return tiles_factor_t<64>(pl, dA, ldm, (cudaStream_t)stream);
    return tiles_factor_t<128>(pl, dA, ldm, (cudaStream_t)stream);
}

int b200qr_tiles_apply_q(b200qr_tileplan_t pl, const float *dA, int ldm, float *dC, int ldc, int ncols, int transpose, void *stream)
{
    if (!pl || !dA || !dC || ncols <= 0 || ncols % pl->b || ldm < pl->m || ldc < pl->m || (ldm & 3) || (ldc & 3)) return B200QR_EINVAL;
    if (pl->b == 32) return tiles_apply_t<32>(pl, dA, ldm, dC, ldc, ncols, transpose, (cudaStream_t)stream);
    if (pl->b == 64) return tiles_apply_t<64>(pl, dA, ldm, dC, ldc, ncols, transpose, (cudaStream_t)stream);
    return tiles_apply_t<128>(pl, dA, ldm, dC, ldc, ncols, transpose, (cudaStream_t)stream);
}

}  // extern "C"
