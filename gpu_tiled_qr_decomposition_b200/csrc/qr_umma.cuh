// qr_umma.cuh -- tcgen05 (5th-generation tensor core) building blocks for sm_100a.
//
// Thin inline-PTX wrappers (TMEM allocation, shared-memory matrix descriptors, the instruction
// descriptor for kind::tf32, MMA issue / commit, TMEM loads) and the 3xTF32 operand split that
// keeps float32 accuracy: a = hi + lo with hi = tf32(a), lo = a - hi, and
// a*b ~= hi_a*hi_b + hi_a*lo_b + lo_a*hi_b (the dropped lo*lo term is ~2^-22 relative).
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix / instruction descriptor" tables.
#pragma once
#include "qr_device.cuh"

namespace qrb200 {

// ---- TMEM allocation (one warp, .sync.aligned) ------------------------------------------------
// The column count is an immediate: with a register operand the toolchain has to assume the
// whole tensor memory and limits the kernel to one CTA per SM.
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_smem)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle, 8-row groups 1024 bytes apart (the layout a
// TMA load with CU_TENSOR_MAP_SWIZZLE_128B produces for 128-byte lines).  Valid for a K-major
// operand (K contiguous inside the 128-byte line) and for an MN-major operand whose MN extent is one
// 128-byte line (32 floats): in both cases the second dimension advances in 1024-byte groups of 8.
__device__ __forceinline__ uint64_t umma_desc_sw128(const void *smem, uint32_t byte_offset)
{
    const uint32_t addr = smem_u32(smem) + byte_offset;
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);            // start address, 16-byte units
    d |= (uint64_t)1 << 16;                           // leading byte offset (16 B; unused within one swizzle atom)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset: next group of 8 lines
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}
// Instruction descriptor for kind::tf32, fp32 accumulate, dense.  a_mn / b_mn: operand is MN-major.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, bool a_mn, bool b_mn, bool neg_a = false)
{
    return (1u << 4)            // D format: F32
           | (2u << 7)          // A format: TF32
           | (2u << 10)         // B format: TF32
           | ((neg_a ? 1u : 0u) << 13)
           | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16)
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// same with the A operand in tensor memory (128 lanes = rows, K = 8 consecutive 32-bit columns)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// registers -> TMEM: warp w of the warpgroup writes lanes 32w..32w+31, 32 consecutive columns each
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const float (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])),
        "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])),
        "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])),
        "r"(__float_as_uint(v[31]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// Cheap descriptor arithmetic for issue loops: the high word of a SW128 descriptor is constant, the low word
// is (address >> 4) | LBO; a K step of 8 floats (32 bytes) adds 2 to it.
constexpr uint32_t UMMA_DESC_HI_SW128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(const void *smem) { return ((smem_u32(smem) >> 4) & 0x3fff) | (1u << 16); }
__device__ __forceinline__ uint64_t umma_desc_from(uint32_t lo)
{
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(UMMA_DESC_HI_SW128));
    return d;
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMEM -> registers: warp w of the warpgroup reads lanes 32w..32w+31, 32 consecutive columns each
__device__ __forceinline__ void tmem_ld_32x32(float (&v)[32], uint32_t taddr)
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- 3xTF32 split --------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_hi(float a)
{
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(a));
    return __uint_as_float(h);
}
// Split the 128-byte line `lane` of a swizzled tile in place into hi (stays) and lo (other tile).
__device__ __forceinline__ void split_line_sw(float *tile_hi, float *tile_lo, int line)
{
    char *bh = reinterpret_cast<char *>(tile_hi) + line * 128;
    char *bl = reinterpret_cast<char *>(tile_lo) + line * 128;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4 a = *reinterpret_cast<float4 *>(bh + i * 16);       // position-wise: the swizzle permutes chunks only
        float4 h = make_float4(tf32_hi(a.x), tf32_hi(a.y), tf32_hi(a.z), tf32_hi(a.w));
        *reinterpret_cast<float4 *>(bh + i * 16) = h;
        *reinterpret_cast<float4 *>(bl + i * 16) = make_float4(a.x - h.x, a.y - h.y, a.z - h.z, a.w - h.w);
    }
}

}  // namespace qrb200
