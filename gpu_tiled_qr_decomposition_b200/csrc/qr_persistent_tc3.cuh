// qr_persistent_tc3.cuh -- third tensor-core sweep engine (QRB200_TC=3; the default for large matrices), included by
// qr_persistent.cu.  Design, measurements and what bounds it: DESIGN.md section 6, profiles/r1_measurements.md.
//
// Same mathematics as the other two tensor-core engines -- the reference's TSMQR / UNMQR (qrdecomp.c:620-644, 736-762)
// in compact-WY form, every product 3xTF32 on tcgen05 with fp32 accumulation in tensor memory -- rebuilt around what
// the first two measured: a tile row of a sweep is a CHAIN through the resident tile row k (c_top), and the only way to
// make the chain short is to take everything else off it.  For a block of 32 reflectors (V, T) and 128 matrix columns:
//     Y = C_bot^T V                  product 1: off the chain, issued rows ahead       (issuer A)
//     S = c_top + Y                  per thread                                          (chain warps)
//     W = S T                        product 2: the only MMA on the chain                (issuer B)
//     c_top -= W                     per thread                                          (chain warps)
//     C_bot^T -= W V^T               product 3: off the chain, accumulator preloaded     (issuer C)
// T is NOT rebuilt by every sweep: a panel CTA runs pairs of warps -- the factor warp runs GEQRT / the TSQRT chain, its
// builder warp forms T and the six operand tiles already split into 3xTF32 halves (build_t_warp, qr_device.cuh), stores
// them into every rank's workspace and publishes the tile (factor_main / builder_main in the body file).
// Warp roles of a sweep CTA (13 warps, one CTA per SM, each role a function with a register allocation of its own):
//     loader   polls the progress words four rows ahead; one TMA box per tile row (IN ring) and per operand (V / T /
//              V^T rings), every stream gated by its own ring
//     bulk x4  thread = matrix column: preloads the product-3 accumulator with the column and writes its hi / lo halves
//              as the A operand of product 1 (tensor memory); three rows later reads the finished column back into an
//              OUT stage
//     chain x4 thread = matrix column: the chain above; c_top stays in registers
//     storer   TMA-stores finished rows, publishes their progress words two rows per fence
//     issuers A / B / C   one thread each issues the MMAs of product 1 / 2 / 3 (a thread issues one tcgen05.mma of this
//              size per ~50 cycles; the pipe takes several threads at once)
#pragma once
#include "qr_umma.cuh"

namespace qrb200 {

// Ring depths, overridable for A/B builds (-DTC3_NCI=4 ...).  Round-2 A/B at 16384^2 / 32768^2 (IN, OUT, operands):
// (3,5,4) 132.6 / 1017 ms -- (4,5,3) 134.8 / 1035 -- (4,4,4) 144.8 / 1091 -- (5,4,3) 148.3 / 1108 -- (3,6,3) 134.4 / 1034:
// the OUT ring is the one that matters and five stages are enough; releasing an OUT stage as soon as the TMA store has
// READ it (wait_group.read) instead of at the row's publication changed nothing (132.4 / 1015).
#ifndef TC3_NCI
#define TC3_NCI 3
#endif
#ifndef TC3_NCO
#define TC3_NCO 5
#endif
#ifndef TC3_NOS
#define TC3_NOS 4
#endif
template <>
struct Cfg<5> {
    static constexpr int G = 4;
    static constexpr int NS = 3;                                   // only places the panel warps' barriers (run_panel)
    static constexpr int NCI = TC3_NCI;                                  // IN ring: a tile row of the sweep as loaded
    static constexpr int NCO = TC3_NCO;                                  // OUT ring: a finished tile row for the storer
    static constexpr int NOS = TC3_NOS;                                  // operand rings (V, T, V^T: hi and lo tile each)
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
