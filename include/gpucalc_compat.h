/*
 * gpucalc_compat.h -- the reference's GPU entry points, as exported by libqrb200.so.
 *
 * A caller written against the reference's include/gpucalc.h (lines 3-6) links against
 * libqrb200.so unchanged: same symbol names, same argument lists, C linkage, void
 * returns.  Behavioural notes a maintainer should know are in INTEGRATION.md.
 */
#ifndef QRB200_GPUCALC_COMPAT_H
#define QRB200_GPUCALC_COMPAT_H

#ifdef __cplusplus
extern "C" {
#endif

/* Replaces gpucalc.cu:1579-1686.  Host matrix `mat` (column-major, leading dimension
 * ldm, m and n multiples of 32) is factored in place: R above/on the diagonal, the
 * Householder vectors below.  Prints "(m,n)\n<T> tasks\n" and "GPU: <ms> ms\n" like the
 * reference unless QRB200_QUIET=1.  `maxblocks` capped the reference's persistent grid
 * at a Fermi-era co-residency limit; here the grid is sized from the device and
 * maxblocks is only honoured when QRB200_HONOR_MAXBLOCKS=1.  The workspace is kept between calls
 * of the same shape (b200qr_release_cache() in qrb200.h frees it).  cudaDeviceReset() is only
 * issued when QRB200_COMPAT_RESET=1 (the reference always resets, gpucalc.cu:1685). */
void cudaQRTask(float *mat, int m, int n, int ldm, int maxblocks);

/* Declared at gpucalc.h:4 but left undefined by the reference (body commented out,
 * gpucalc.cu:1801-1877).  Here: host-ordered launches of the same four tile kernels
 * in the closed-form serial order, ldm = m.  Same output as cudaQRTask. */
void cudaQRFull(float *mat, int m, int n);

/* Replaces gpucalc.cu:1706-1774: times the TSMQR ("DAPP") kernel on nblocks independent
 * 64x32 column pairs, n repetitions; timings[t] = elapsed_ms * nblocks. */
void testDAPP(float *timings, int n, int nblocks);

/* Replaces gpucalc.cu:1776-1799: one DAPP on a 64x64 column-major host matrix (ldm 64):
 * tau = mat[0..31], V2 = tile(1,0), A = tile(0,1), B = tile(1,1); in place. */
void doCUDADAPP(float *mat);

#ifdef __cplusplus
}
#endif
#endif
