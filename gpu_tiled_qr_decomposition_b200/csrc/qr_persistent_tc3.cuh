// qr_persistent_tc3.cuh -- third tensor-core sweep engine (QRB200_TC=3), included by qr_persistent.cu.
//
// Same mathematics as the other two tensor-core engines -- the reference's TSMQR / UNMQR
// (qrdecomp.c:620-644, 736-762) in compact-WY form, every product 3xTF32 on tcgen05 with fp32
// accumulation in tensor memory -- rebuilt around what the first two measured: a tile row of a sweep
// is a CHAIN through the resident tile row k (c_top), and the only way to make the chain short is to
// take everything else off it.  For a block of 32 reflectors (V, T) and 128 matrix columns:
//     Y = C_bot^T V                  off the chain, issued rows ahead                (issuer A)
//     S = c_top + Y                  per thread                                        (chain warps)
//     W = S T                        tcgen05.mma, K = 32                               (issuer B)
//     c_top -= W                     per thread                                        (chain warps)
//     C_bot^T -= W V^T               off the chain: accumulator preloaded with C_bot^T (issuer B)
// so one chain step is: TMEM load, 32 adds, split, TMEM store, 12 small MMAs.  T is NOT rebuilt by
// every sweep: the panel warp that factors a tile also emits its T factor and the transposed block V^T
// (build_t_warp, qr_device.cuh) into two workspace matrices with the layout of A, and the sweeps
// TMA-load them next to V.  Warp roles of a sweep CTA (13 warps, one CTA per SM):
//     loader   polls the progress words, TMA-loads the row's 4 tiles into the C ring (8 stages) and
//              V, T, V^T into the operand ring (4 stages)
//     bulk x4  thread = matrix column: preloads the product-3 accumulator with the column and writes
//              its hi / lo halves as the A operand of product 1 (all in tensor memory, 3 buffers);
//              three rows later reads the finished column back into the C stage
//     chain x4 thread = matrix column: the chain above; c_top stays in registers
//     storer   TMA-stores finished rows (two stores in flight), publishes their progress words
//     issuers A / B / C   one thread each issues the MMAs of product 1 / 2 / 3
// Panel chains run in CTAs of their own (four panel warps each), as in engine 2.
#pragma once
#include "qr_umma.cuh"

namespace qrb200 {

template <>
struct Cfg<5> {
    static constexpr int G = 4;
    static constexpr int NS = 3;                                   // only places the panel warps' barriers (run_panel)
    static constexpr int NCI = 3;                                  // IN ring: a tile row of the sweep as loaded
    static constexpr int NCO = 5;                                  // OUT ring: a finished tile row for the storer
    static constexpr int NOS = 4;                                  // operand rings (V, T, V^T: hi and lo tile each)
    static constexpr int CSTAGE_BYTES = G * TILE_BYTES;
    static constexpr int OSTAGE_BYTES = 2 * TILE_BYTES;
    static constexpr int OFF_OUT = NCI * CSTAGE_BYTES;
    static constexpr int OFF_V = OFF_OUT + NCO * CSTAGE_BYTES;
    static constexpr int OFF_T = OFF_V + NOS * OSTAGE_BYTES;
    static constexpr int OFF_X = OFF_T + NOS * OSTAGE_BYTES;
    static constexpr int SWEEP_END = OFF_X + NOS * OSTAGE_BYTES;
    static constexpr int PANEL_BYTES = 11 * TILE_BYTES + 5120;     // in 2, out 2, R / scratch, six operand tiles, Gram scratch
    static constexpr int OFF_PANEL = 0;                            // panel CTAs never run sweeps
    static constexpr int OFF_BAR = SWEEP_END;
    static constexpr int SMEM_BYTES = OFF_BAR + 1024;
    static constexpr bool BUILD_T = true;
};

#define TC3_NS tc3
#define TC3_PROF 0
#define TC3_KERNEL tiledqr_persistent_tc3
#include "qr_persistent_tc3_body.inc"
#undef TC3_NS
#undef TC3_PROF
#undef TC3_KERNEL
#define TC3_NS tc3p
#define TC3_PROF 1
#define TC3_KERNEL tiledqr_persistent_tc3_prof
#include "qr_persistent_tc3_body.inc"
#undef TC3_NS
#undef TC3_PROF
#undef TC3_KERNEL

}  // namespace qrb200
