// tc_common.cuh — PTX wrappers shared by the tcgen05 GEMM kernels (mlp_tc2.cu): mbarrier, TMA (cp.async.bulk[.tensor]),
// UMMA descriptors, TMEM loads.  Layout facts used below were pinned on a B200 by scripts/umma_probe.cu
// (profiles/r02_umma_layout_probe.txt):
//   * K-major operand  <- TMA tile with CU_TENSOR_MAP_SWIZZLE_128B, descriptor layout type 2, SBO = 1024 B (8 rows of
//     128 B), the leading-byte offset is ignored; a k-step of 8 tf32 advances the start address by 32 B.
//   * MN-major tf32 operand <- TMA tile with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (32-byte atoms: the only MN-major
//     layout kind::tf32 has), descriptor layout type 1, LBO = byte stride between 32-channel groups, SBO = 512 B
//     (4 k-rows of 128 B); a k-step of 8 rows advances the start address by 1024 B.
//   * tiles may be rewritten in place through the generic proxy (fence.proxy.async before the consumer is signalled).
//   * TMA tile stores from a SWIZZLE_128B staging tile clip rows / columns beyond the tensor-map extents.
#pragma once
#include "common.cuh"
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)

namespace rsbtc {

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rsb_smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rsb_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rsb_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = rsb_smem_addr(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(0x989680u)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 1-D bulk copy global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(rsb_smem_addr(bar))
                 : "memory");
}
// 2-D tensor tile global -> shared (c0 = inner / column coordinate, c1 = row coordinate; out-of-range parts arrive as zeros)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(rsb_smem_addr(bar))
                 : "memory");
}
// four rows r0..r3 of a 2-D tensor (tensor map with a ONE-row box of `box cols` columns starting at column c0) -> four
// consecutive box rows at dst; with SWIZZLE_128B the 16-byte chunks are XOR-ed with the shared-memory line index exactly as a
// plain tile load does (hardware probe: scripts/gather4_probe.cu, profiles/r02_gather4_probe.txt)
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, int c0, int r0, int r1, int r2, int r3, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(rsb_smem_addr(bar))
                 : "memory");
}
// 2-D tensor tile shared -> global (clipped to the tensor-map extents), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, uint32_t src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"((uint64_t)map), "r"(c0), "r"(c1),
                 "r"(src)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): [0,14) start >> 4, [16,30) LBO >> 4,
// [32,46) SBO >> 4, [46,48) version = 1, [61,64) layout type (0 none, 1 128B with 32-byte atoms, 2 128B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}
// instruction descriptor, kind::tf32, fp32 accumulate; mn_major sets the major-ness of BOTH operands
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, bool mn_major)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(rsb_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v)
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// round-to-nearest (ties away) to tf32 for finite inputs: add half a tf32 ulp to the magnitude, clear 13 mantissa bits
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
// a = hi + lo with hi = tf32(a) (rounded) and lo = a - hi: the 3xTF32 split.  The residual is handed to the tensor core as
// a plain fp32 value: kind::tf32 reads the upper 19 bits of its operands, i.e. truncates lo to tf32 itself (representation
// error <= 2^-22 |a|, against 2^-23 with an explicit rounding that costs two more integer instructions per element).
// NaN: to_tf32 maps the canonical NaN 0x7fffffff to -0, but lo = NaN - hi stays NaN, so a diverged activation still
// reaches the accumulator as NaN (torch semantics).  -DRSB_LO_RAW=0 restores the explicit rounding of lo.
#ifndef RSB_LO_RAW
#define RSB_LO_RAW 1
#endif
__device__ __forceinline__ void split4(const float4 v, float4 &hi, float4 &lo)
{
    hi.x = to_tf32(v.x); hi.y = to_tf32(v.y); hi.z = to_tf32(v.z); hi.w = to_tf32(v.w);
#if RSB_LO_RAW
    lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
#else
    lo.x = to_tf32(v.x - hi.x); lo.y = to_tf32(v.y - hi.y); lo.z = to_tf32(v.z - hi.z); lo.w = to_tf32(v.w - hi.w);
#endif
}
__device__ __forceinline__ float relu_nan(float x) { return rsb_relu(x); }

}  // namespace rsbtc
