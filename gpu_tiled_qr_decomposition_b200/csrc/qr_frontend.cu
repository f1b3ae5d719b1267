// qr_frontend.cu -- consumers of "R plus the compact Householder vectors"  and the arbitrary-size
// front-end (8f rank 4).  The reference has none of these: it never forms or applies Q (qrdecomp.c:1292-1311 multAB is
// an unused stub) and silently truncates m, n to multiples of 32 (qrdecomp.c:162, gpucalc.cu:1582).
//
//   b200qr_form_q      Q(:, 0:ncols) = Q I(:, 0:ncols)                 (identity + b200qr_apply_q)
//   b200qr_trsm_upper  X <- R^-1 X, R = upper triangle of a factored matrix (blocked back substitution)
//   b200qr_solve_ls    min |A x - b|: B <- Q^T B, X = R^-1 B(0:n, :)
//   b200qr_any_*       m, n NOT multiples of 32: the matrix is embedded in a zero-padded (mp x np) device buffer.
//                      QR of [A 0; 0 0] has the same R and the same reflectors on the real rows (padded rows carry zero
//                      reflector entries; a zero column gives the reference's nrm == 0 case, qrdecomp.c:1219: v
//                      unscaled = 0, tau = 2, R entry 0), so the m x n part of the result is the factorisation of A.
#include "qr_device.cuh"
#include "../../include/qrb200.h"

#include <cstdint>
#include <vector>

namespace qrb200 {

__global__ void identity_cols_kernel(float *Q, int ldq, int m, int ncols)
{
    const long long total = (long long)m * ncols;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e % m), c = (int)(e / m);
        Q[(size_t)c * ldq + r] = r == c ? 1.0f : 0.f;
    }
}

// Back substitution R X = B for one block of 32 right-hand sides per CTA (thread = row of the current update,
// warp 0 = the 32 x 32 diagonal solve).  R: n x n upper triangle, column-major (ldr); B: n x nrhs (ldb), overwritten.
// Columns of R are walked right to left in blocks of 32: solve the diagonal block for all 32 right-hand sides (lane =
// right-hand side, 32 serial steps in registers), then B(0:j0, :) -= R(0:j0, j0:j0+32) X(j0:j0+32, :) by all threads
// (thread = row: R is column-major, so consecutive threads read consecutive addresses).
__global__ void __launch_bounds__(256) trsm_upper_kernel(const float *R, int ldr, int n, float *B, int ldb, int nrhs)
{
    __shared__ float xs[TB][TB + 1];                        // xs[k][rhs]: solved block
    __shared__ float rs[TB][TB + 1];                        // rs[i][k] = R(j0 + i, j0 + k)
    const int rhs0 = blockIdx.x * TB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nr = nrhs - rhs0 < TB ? nrhs - rhs0 : TB;
    for (int j1 = n; j1 > 0; j1 -= TB) {
        const int j0 = j1 - TB < 0 ? 0 : j1 - TB, w = j1 - j0;
        for (int e = tid; e < TB * TB; e += blockDim.x) {
            const int i = e & 31, k = e >> 5;
            rs[i][k] = (i < w && k < w && i <= k) ? R[(size_t)(j0 + k) * ldr + j0 + i] : (i == k ? 1.0f : 0.f);
        }
        __syncthreads();
        if (warp == 0) {
            float y[TB];
#pragma unroll
            for (int i = 0; i < TB; ++i) y[i] = (i < w && lane < nr) ? B[(size_t)(rhs0 + lane) * ldb + j0 + i] : 0.f;
#pragma unroll
            for (int k = TB - 1; k >= 0; --k) {
                y[k] = y[k] / rs[k][k];
#pragma unroll
                for (int i = 0; i < TB; ++i)
                    if (i < k) y[i] = fmaf(-rs[i][k], y[k], y[i]);
            }
#pragma unroll
            for (int i = 0; i < TB; ++i) {
                xs[i][lane] = y[i];
                if (i < w && lane < nr) B[(size_t)(rhs0 + lane) * ldb + j0 + i] = y[i];
            }
        }
        __syncthreads();
        // trailing update of the rows above the block
        for (int r = tid; r < j0; r += blockDim.x) {
            float acc[TB];
#pragma unroll
            for (int c = 0; c < TB; ++c) acc[c] = 0.f;
            for (int k = 0; k < w; ++k) {
                const float rk = R[(size_t)(j0 + k) * ldr + r];
#pragma unroll
                for (int c = 0; c < TB; ++c) acc[c] = fmaf(rk, xs[k][c], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < TB; ++c)
                if (c < nr) B[(size_t)(rhs0 + c) * ldb + r] -= acc[c];
        }
        __syncthreads();
    }
}

// dst (rows x cols, ld) <- upper triangle of src, zero below
__global__ void copy_upper_kernel(const float *src, int lds, float *dst, int ldd, int rows, int cols)
{
    const long long total = (long long)rows * cols;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e % rows), c = (int)(e / rows);
        dst[(size_t)c * ldd + r] = r <= c ? src[(size_t)c * lds + r] : 0.f;
    }
}

}  // namespace qrb200

using namespace qrb200;

static thread_local int g_fe_cuda = 0;
#define CKF(call)                                                          \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) { g_fe_cuda = (int)e_; return B200QR_ECUDA; } \
    } while (0)

struct b200qr_any_s {
    int m, n, mp, np;
    b200qr_plan_t plan = nullptr;
    float *dA = nullptr;                 // mp x np, ld = mp
    float *dC = nullptr;                 // scratch for apply / solve: mp x cap_cols
    int cap_cols = 0;
};

extern "C" {

int b200qr_frontend_last_cuda_error(void) { return g_fe_cuda; }

int b200qr_form_q(const float *dA, int m, int n, int ldm, const float *dTauCompact, float *dQ, int ldq, int ncols, void *stream)
{
    if (!dA || !dTauCompact || !dQ || ncols <= 0 || ncols % TB || ncols > m || ldq < m || (ldq & 3)) return B200QR_EINVAL;
    identity_cols_kernel<<<1024, 256, 0, (cudaStream_t)stream>>>(dQ, ldq, m, ncols);
    CKF(cudaGetLastError());
    return b200qr_apply_q(dA, m, n, ldm, dTauCompact, dQ, ldq, ncols, 0, stream);
}

int b200qr_trsm_upper(const float *dR, int n, int ldr, float *dB, int ldb, int nrhs, void *stream)
{
    if (!dR || !dB || n <= 0 || nrhs <= 0 || ldr < n || ldb < n) return B200QR_EINVAL;
    trsm_upper_kernel<<<(nrhs + TB - 1) / TB, 256, 0, (cudaStream_t)stream>>>(dR, ldr, n, dB, ldb, nrhs);
    CKF(cudaGetLastError());
    return B200QR_OK;
}

int b200qr_solve_ls(const float *dA, int m, int n, int ldm, const float *dTauCompact, float *dB, int ldb, int nrhs, void *stream)
{
    if (m < n || nrhs <= 0 || nrhs % TB) return B200QR_EINVAL;
    int rc = b200qr_apply_q(dA, m, n, ldm, dTauCompact, dB, ldb, nrhs, 1, stream);
    if (rc != B200QR_OK) return rc;
    return b200qr_trsm_upper(dA, n, ldm, dB, ldb, nrhs, stream);
}

// ---- arbitrary sizes
int b200qr_any_create(b200qr_any_t *out, int m, int n, const b200qr_opts *opts)
{
    if (!out || m <= 0 || n <= 0) return B200QR_EINVAL;
    b200qr_any_s *h = new b200qr_any_s;
    h->m = m; h->n = n;
    h->mp = (m + TB - 1) / TB * TB; h->np = (n + TB - 1) / TB * TB;
    int rc = b200qr_plan_create(&h->plan, h->mp, h->np, opts);
    if (rc == B200QR_OK && cudaMalloc(&h->dA, sizeof(float) * (size_t)h->mp * h->np) != cudaSuccess) rc = B200QR_ENOMEM;
    if (rc != B200QR_OK) { b200qr_any_destroy(h); return rc; }
    *out = h;
    return B200QR_OK;
}

int b200qr_any_destroy(b200qr_any_t h)
{
    if (!h) return B200QR_EINVAL;
    if (h->plan) b200qr_plan_destroy(h->plan);
    cudaFree(h->dA); cudaFree(h->dC);
    delete h;
    return B200QR_OK;
}

int b200qr_any_padded_rows(b200qr_any_t h) { return h ? h->mp : B200QR_EINVAL; }
int b200qr_any_padded_cols(b200qr_any_t h) { return h ? h->np : B200QR_EINVAL; }
const float *b200qr_any_factored(b200qr_any_t h) { return h ? h->dA : nullptr; }
const float *b200qr_any_tau(b200qr_any_t h) { return h ? b200qr_plan_tau_device(h->plan) : nullptr; }

// dA: m x n device matrix (column-major, ldm >= m, any alignment); it is not modified.
int b200qr_any_factor(b200qr_any_t h, const float *dA, int ldm, void *stream_)
{
    if (!h || !dA || ldm < h->m) return B200QR_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    CKF(cudaMemsetAsync(h->dA, 0, sizeof(float) * (size_t)h->mp * h->np, stream));
    CKF(cudaMemcpy2DAsync(h->dA, sizeof(float) * h->mp, dA, sizeof(float) * ldm, sizeof(float) * h->m, h->n,
                          cudaMemcpyDeviceToDevice, stream));
    return b200qr_factor_device(h->plan, h->dA, h->mp, stream);
}

// R: min(m, n) x n upper trapezoid (column-major, ldr >= min(m, n))
int b200qr_any_get_r(b200qr_any_t h, float *dR, int ldr, void *stream)
{
    const int k = h ? (h->m < h->n ? h->m : h->n) : 0;
    if (!h || !dR || ldr < k) return B200QR_EINVAL;
    copy_upper_kernel<<<1024, 256, 0, (cudaStream_t)stream>>>(h->dA, h->mp, dR, ldr, k, h->n);
    CKF(cudaGetLastError());
    return B200QR_OK;
}

static int any_scratch(b200qr_any_s *h, int cols)
{
    if (cols <= h->cap_cols) return B200QR_OK;
    cudaFree(h->dC);
    h->dC = nullptr; h->cap_cols = 0;
    if (cudaMalloc(&h->dC, sizeof(float) * (size_t)h->mp * cols) != cudaSuccess) { (void)cudaGetLastError(); return B200QR_ENOMEM; }
    h->cap_cols = cols;
    return B200QR_OK;
}

// C (m x ncols, ldc) <- Q^T C (transpose != 0) or Q C, Q = the m x m orthogonal factor
int b200qr_any_apply_q(b200qr_any_t h, float *dC, int ldc, int ncols, int transpose, void *stream_)
{
    if (!h || !dC || ncols <= 0 || ldc < h->m) return B200QR_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int cp = (ncols + TB - 1) / TB * TB;
    int rc = any_scratch(h, cp);
    if (rc != B200QR_OK) return rc;
    CKF(cudaMemsetAsync(h->dC, 0, sizeof(float) * (size_t)h->mp * cp, stream));
    CKF(cudaMemcpy2DAsync(h->dC, sizeof(float) * h->mp, dC, sizeof(float) * ldc, sizeof(float) * h->m, ncols,
                          cudaMemcpyDeviceToDevice, stream));
    rc = b200qr_apply_q(h->dA, h->mp, h->np, h->mp, b200qr_plan_tau_device(h->plan), h->dC, h->mp, cp, transpose, stream);
    if (rc != B200QR_OK) return rc;
    CKF(cudaMemcpy2DAsync(dC, sizeof(float) * ldc, h->dC, sizeof(float) * h->mp, sizeof(float) * h->m, ncols,
                          cudaMemcpyDeviceToDevice, stream));
    return B200QR_OK;
}

// Least squares min |A x - b| for m >= n (A of full column rank): dB m x nrhs (ldb) is not modified, dX n x nrhs (ldx)
int b200qr_any_solve_ls(b200qr_any_t h, const float *dB, int ldb, int nrhs, float *dX, int ldx, void *stream_)
{
    if (!h || !dB || !dX || nrhs <= 0 || h->m < h->n || ldb < h->m || ldx < h->n) return B200QR_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int cp = (nrhs + TB - 1) / TB * TB;
    int rc = any_scratch(h, cp);
    if (rc != B200QR_OK) return rc;
    CKF(cudaMemsetAsync(h->dC, 0, sizeof(float) * (size_t)h->mp * cp, stream));
    CKF(cudaMemcpy2DAsync(h->dC, sizeof(float) * h->mp, dB, sizeof(float) * ldb, sizeof(float) * h->m, nrhs,
                          cudaMemcpyDeviceToDevice, stream));
    rc = b200qr_apply_q(h->dA, h->mp, h->np, h->mp, b200qr_plan_tau_device(h->plan), h->dC, h->mp, cp, 1, stream);
    if (rc != B200QR_OK) return rc;
    rc = b200qr_trsm_upper(h->dA, h->n, h->mp, h->dC, h->mp, nrhs, stream);      // the real n x n triangle only
    if (rc != B200QR_OK) return rc;
    CKF(cudaMemcpy2DAsync(dX, sizeof(float) * ldx, h->dC, sizeof(float) * h->mp, sizeof(float) * h->n, nrhs,
                          cudaMemcpyDeviceToDevice, stream));
    return B200QR_OK;
}

}  // extern "C"
