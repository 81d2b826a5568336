// mlp_tc.cu — the shared-MLP GEMMs of RepSurf-U on the 5th-gen tensor cores (tcgen05 + TMEM), fp32-faithful.
//
// Two kernels cover forward, input-gradient and weight-gradient of every 1x1-conv / linear layer on the path:
//
//   gemm_rows_kernel :  Y[r, n]  = sum_k A(r, k) * W[n, k]        (rows x K) @ (N x K)^T     forward + dgrad
//   gemm_wgrad_kernel:  dW[m, n] += sum_r G(r, m) * X(r, n)       (rows x M)^T @ (rows x N)  wgrad (operands
//                       transposed into K-major tiles while they are staged)
//
// They replace the reference's library calls nn.Conv2d/Conv1d(1x1)/nn.Linear (+ autograd) on the RepSurf path
//   classification/modules/repsurface_utils.py:236-243, segmentation/modules/repsurface_utils.py:220-227, 267-282
// and FUSE what the reference runs as separate cuDNN / elementwise passes around them.  Operands are never
// materialised in HBM in their "activated" form: A(r,k), G(r,m), X(r,n) are computed while the tile is staged
// (`Opnd`): train-mode BatchNorm + ReLU of the previous layer, the channel-de-differentiated first layer
// relu(bn_l(y_l) + bn_f(y_f)), the BatchNorm-backward affine dY = a*dZ + b*Y + d, and the max-pool backward
// selection.  Epilogues (`Epi`) add the bias, apply the ReLU mask of the layer below (dgrad) and accumulate the
// per-channel statistics the NEXT BatchNorm step needs (forward: sum, sum^2; backward: sum dZ, sum dZ*xhat) in fp64.
//
// Numerics: tcgen05 has no fp32 x fp32 MMA.  Every operand is split a = hi + lo with hi = tf32(a),
// lo = tf32(a - hi); three kind::tf32 MMAs (lo*hi + hi*lo + hi*hi) accumulate in fp32 in TMEM ("3xTF32"):
// ~2^-21 relative per product, i.e. fp32-level, which the 1e-5 parity bar needs (plain TF32 gives 1e-3).
//
// Structure of both kernels (one persistent CTA per SM, 9 warps):
//   warps 0-3  producers: evaluate the operand transform, split hi/lo, write both tiles to shared memory in the
//              UMMA canonical no-swizzle layout (8 x 16 B core matrices); gemm_rows also fetches the matching
//              pre-split weight chunk with cp.async.bulk (TMA 1-D) onto the same mbarrier.
//   warp  4    MMA issuer: one elected lane issues 4 k-steps x 3 tcgen05.mma per 32-deep chunk into a
//              double-buffered TMEM accumulator (128 lanes x <=256 columns), tcgen05.commit -> mbarriers.
//   warps 5-8  epilogue: tcgen05.ld 32x32b -> registers -> epilogue transform -> global (+ fp64 statistics).
#include "common.cuh"
#include "mlp_tc.h"

namespace {

constexpr int TM = 128;          // UMMA M
constexpr int KC = 32;           // reduction elements per pipeline chunk (4 UMMA k-steps of 8)
constexpr int STAGES_MAX = 4;     // pipeline depth is chosen per launch from the shared-memory budget
constexpr int NT_MAX = 256;      // UMMA N <= 256
constexpr int A_TILE_BYTES = TM * KC * 4;           // 16 KB (one of hi / lo), K-major tiles
constexpr int PROD_WARPS = 8;                      // operand-staging warps: two groups of 4, alternating chunks
constexpr int GROUP_THREADS = 128;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int MMA_WARP = PROD_WARPS;               // warp index of the MMA issuer
constexpr int THREADS = (PROD_WARPS + 1 + 4) * 32; // + 4 epilogue warps

typedef rsb_opnd_t Opnd;
typedef rsb_epi_t Epi;

struct RowsParams {
    Opnd A;
    Epi E;
    const float *Wp;
    long rows;
    int N, NT, n_tiles, k_chunks, stages;
    int raw_depth;   // > 0: operand pieces are prefetched with cp.async into a per-thread ring of this depth
};

struct WgradParams {
    Opnd G, X;
    float *dW;
    int ldw;
    long rows;
    int M, N, NT, m_tiles, n_tiles, stages;
    int raw_slots;       // > 0: operands are prefetched with cp.async into a ring of this many raw slabs
    int raw_slot_bytes;
};

// ---- PTX wrappers --------------------------------------------------------------------------------------
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
            : "r"(addr), "r"(parity), "r"(0x989680u)   // suspend-time hint: the warp sleeps in hardware instead of spinning
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     rsb_smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(rsb_smem_addr(bar))
                 : "memory");
}

// UMMA shared-memory descriptor, no swizzle (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   [0,14) start>>4, [16,30) leading-dim byte offset>>4, [32,46) stride-dim byte offset>>4, [46,48) version = 1,
//   [61,64) layout = 0 (INTERLEAVE).  K-major:  LBO = between core matrices along K, SBO = between 8-row groups.
//   MN-major: LBO = between 8-deep groups along K, SBO = between 4-channel groups along MN.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// instruction descriptor, kind::tf32, fp32 accumulate (InstrDescriptor bit layout); mn_major sets both operands
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, bool mn_major)
{
    return (1u << 4) /* D = f32 */ | (2u << 7) /* A = tf32 */ | (2u << 10) /* B = tf32 */ |
           (mn_major ? ((1u << 15) | (1u << 16)) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(rsb_smem_addr(bar))
                 : "memory");
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
// round-to-nearest (ties away) to tf32, as cvt.rna.tf32.f32 does for every finite input: add half a tf32 ulp to the
// magnitude, clear the 13 low mantissa bits.  The PTX instruction compiles to the same two integer operations PLUS
// an |x| < inf test and a select per value (8 conversions per staged float4: a fifth of the split's instructions);
// the guard only matters for NaN payloads, which never enter a GEMM here (check_nan_umb repairs the descriptors).
__device__ __forceinline__ float to_tf32(float x)
{
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// ---- operand transform ------------------------------------------------------------------------------------
// Per-channel coefficients are hoisted into registers once per (thread, chunk); loads are 128-bit when the
// operand's row pitch and channel offset allow it.
struct Coef4 {
    float a[4], b[4], d[4], a2[4], d2[4];
};

__device__ __forceinline__ void coef_load(const Opnd &O, int k, int nv, Coef4 &c)
{
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const bool ok = e < nv;
        c.a[e] = (ok && O.a) ? __ldg(O.a + k + e) : 0.f;
        c.d[e] = (ok && O.d) ? __ldg(O.d + k + e) : 0.f;
        c.b[e] = (ok && O.b) ? __ldg(O.b + k + e) : 0.f;
        if (O.kind == RSB_OPND_DUAL_BN_RELU) {
            c.a2[e] = ok ? __ldg(O.a + O.ku + k + e) : 0.f;
            c.d2[e] = ok ? __ldg(O.d + O.ku + k + e) : 0.f;
        }
    }
}

__device__ __forceinline__ void ld4(const float *p, bool vec, int nv, float *v)
{
    if (vec) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = e < nv ? __ldg(p + e) : 0.f;
    }
}

struct OpndFlags {
    bool vecU, vecV, vecU2;
};

__device__ __forceinline__ OpndFlags opnd_flags(const Opnd &O)
{
    OpndFlags f;
    const bool k0ok = (O.k0 & 3) == 0;
    f.vecU = k0ok && (O.ldu & 3) == 0 && ((uintptr_t)O.U & 15) == 0 && (O.kind != RSB_OPND_AFFINE2 || (O.ku & 3) == 0);
    f.vecU2 = f.vecU && (O.ku & 3) == 0;
    f.vecV = k0ok && O.V && (O.ldv & 3) == 0 && ((uintptr_t)O.V & 15) == 0;
    return f;
}

// Math of the operand transform on already-fetched quads: u = first piece, w = second piece
// (DUAL: the second half of U; AFFINE2 / POOLED: V).  POOLED additionally reads the pooled gradient.
__device__ __forceinline__ void opnd_apply4(const Opnd &O, const Coef4 &c, const float *u, const float *w, int k, int nv,
                                            long g, int sidx, float *v)
{
    switch (O.kind) {
    case RSB_OPND_RAW:
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = u[e];
        break;
    case RSB_OPND_BN_RELU:
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = rsb_relu(fmaf(u[e], c.a[e], c.d[e]));
        break;
    case RSB_OPND_DUAL_BN_RELU:
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = rsb_relu(fmaf(u[e], c.a[e], c.d[e]) + fmaf(w[e], c.a2[e], c.d2[e]));
        break;
    case RSB_OPND_AFFINE2:
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = fmaf(c.a[e], u[e], fmaf(c.b[e], w[e], c.d[e]));
        break;
    default:  // RSB_OPND_POOLED: dZ is nonzero only on the arg-max sample of its (group, channel)
#pragma unroll
        for (int e = 0; e < 4; e++) {
            float dz = 0.f;
            if (e < nv && __ldg(O.arg + (size_t)g * O.ldu + k + e) == sidx) dz = __ldg(O.U + (size_t)g * O.ldu + k + e);
            v[e] = fmaf(c.a[e], dz, fmaf(c.b[e], w[e], c.d[e]));
        }
        break;
    }
#pragma unroll
    for (int e = 0; e < 4; e++) v[e] = e < nv ? v[e] : 0.f;
}

// global addresses of the (up to) two 16-byte pieces a quad needs
__device__ __forceinline__ void opnd_pieces(const Opnd &O, long r, int k, const float *&p0, const float *&p1)
{
    p0 = p1 = nullptr;
    switch (O.kind) {
    case RSB_OPND_RAW:
    case RSB_OPND_BN_RELU:
        p0 = O.U + (size_t)r * O.ldu + k;
        break;
    case RSB_OPND_DUAL_BN_RELU:
        p0 = O.U + (size_t)r * O.ldu + k;
        p1 = p0 + O.ku;
        break;
    case RSB_OPND_AFFINE2:
        p0 = O.U + (size_t)r * O.ldu + (k % O.ku);
        p1 = O.V + (size_t)r * O.ldv + k;
        break;
    default:
        p1 = O.V + (size_t)r * O.ldv + k;
        break;
    }
}

// 4 consecutive channels k..k+3 (k includes k0) of row r; nv = valid channels (synchronous fetch).
__device__ __forceinline__ void opnd_eval4(const Opnd &O, const OpndFlags &F, const Coef4 &c, long r, int k, int nv,
                                           long g, int sidx, float *v)
{
    const float *p0, *p1;
    opnd_pieces(O, r, k, p0, p1);
    float u[4] = {0.f, 0.f, 0.f, 0.f}, w[4] = {0.f, 0.f, 0.f, 0.f};
    if (p0) ld4(p0, F.vecU, nv, u);
    if (p1) ld4(p1, O.kind == RSB_OPND_DUAL_BN_RELU ? F.vecU2 : F.vecV, nv, w);
    opnd_apply4(O, c, u, w, k, nv, g, sidx, v);
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void split4(const float *v, float4 &hi, float4 &lo)
{
    hi.x = to_tf32(v[0]); hi.y = to_tf32(v[1]); hi.z = to_tf32(v[2]); hi.w = to_tf32(v[3]);
    // the residual goes to the tensor core unrounded (kind::tf32 truncates it; NaN survives): see tc_common.cuh split4
#ifndef RSB_LO_RAW
#define RSB_LO_RAW 1
#endif
#if RSB_LO_RAW
    lo.x = v[0] - hi.x; lo.y = v[1] - hi.y; lo.z = v[2] - hi.z; lo.w = v[3] - hi.w;
#else
    lo.x = to_tf32(v[0] - hi.x); lo.y = to_tf32(v[1] - hi.y); lo.z = to_tf32(v[2] - hi.z); lo.w = to_tf32(v[3] - hi.w);
#endif
}

// column sums of a warp's 32x32 register tile through a padded shared tile; result for column `lane`
__device__ __forceinline__ float warp_colsum(float *tile, const float *v, int lane)
{
#pragma unroll
    for (int j = 0; j < 32; j++) tile[lane * 33 + j] = v[j];
    __syncwarp();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i++) s += tile[i * 33 + lane];
    __syncwarp();
    return s;
}

struct alignas(128) Barriers {
    uint64_t full[STAGES_MAX], empty[STAGES_MAX], acc_full[2], acc_empty[2];
    uint32_t tmem_slot, pad;
};

__device__ __forceinline__ uint32_t cta_prologue(Barriers *B, int tid, int warp, int full_count = GROUP_THREADS + 1)
{
    if (tid == 0) {
        for (int s = 0; s < STAGES_MAX; s++) { mbar_init(&B->full[s], full_count); mbar_init(&B->empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&B->acc_full[a], 1); mbar_init(&B->acc_empty[a], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(rsb_smem_addr(&B->tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return B->tmem_slot;
}

__device__ __forceinline__ void cta_epilogue(uint32_t tmem_base, int warp)
{
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

// ============================================================================================================
// Y = A @ W^T  (forward and input-gradient)
// ============================================================================================================
__global__ void __launch_bounds__(THREADS, 1) gemm_rows_kernel(const __grid_constant__ RowsParams P)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NT = P.NT;
    const size_t b_bytes = (size_t)NT * KC * 4;  // one of hi / lo
    const size_t stage_bytes = 2 * A_TILE_BYTES + 2 * b_bytes;
    const int STAGES = P.stages;
    Barriers *B = reinterpret_cast<Barriers *>(smem + STAGES * stage_bytes);
    float *stat_tile = reinterpret_cast<float *>(B + 1);  // [4 warps][32][33], then the cp.async ring
    const uint32_t tmem_base = cta_prologue(B, tid, warp, P.raw_depth > 0 ? PROD_THREADS + 1 : GROUP_THREADS + 1);

    const long n_row_tiles = (P.rows + TM - 1) / TM;
    const int kc_count = P.k_chunks;
    const Opnd &A = P.A;
    const Epi &E = P.E;

    if (warp < PROD_WARPS) {
        if (P.raw_depth > 0) {
            // ===== producers, asynchronous path: every thread owns 4 rows x one channel quad of each chunk and
            // prefetches its own 16-byte pieces RD chunks ahead with cp.async into a private shared-memory ring
            // (no registers held, no cross-thread hand-off), so ~RD * 32 KB per SM are always in flight.
            // Everything that does not change per chunk (shared-memory offsets, row pointers, validity) is hoisted:
            // the per-chunk work is address adds, the transform itself and the hi/lo split.
            const int RD = P.raw_depth;
            // thread -> (channel quad k4, rows rsub + 32*j).  The 8 lanes of a quarter-warp take the 8 rows of ONE core
            // matrix (same k4), so a warp-wide 128-bit tile store covers whole 128-byte core matrices: conflict-free.
            // (With k4 = tid & 7 the 8 lanes hit the same 4 banks, 128 B apart: ncu counted 37 M store bank conflicts in
            // 45 M store wavefronts and the LSU data pipe at 65-73 % of peak - the limiter of this kernel.)
            const int k4 = (tid >> 3) & 7, rsub = (tid >> 6) * 8 + (tid & 7);
            constexpr uint32_t PIECE_STRIDE = PROD_THREADS * 16;                       // bytes between (piece, j) planes
            constexpr uint32_t SLOT_BYTES = 2 * 4 * PIECE_STRIDE;
            const uint32_t ring0 = rsb_smem_addr(stat_tile + 4 * 32 * 33) + (uint32_t)tid * 16;
            uint32_t st_off[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int r = rsub + 32 * j;
                st_off[j] = (uint32_t)(((r >> 3) * (KC / 4) * 32 + k4 * 32 + (r & 7) * 4) * 4);
            }
            const bool has0 = A.kind != RSB_OPND_POOLED;
            const bool has1 = A.kind == RSB_OPND_DUAL_BN_RELU || A.kind == RSB_OPND_AFFINE2 || A.kind == RSB_OPND_POOLED;
            const float *b0 = A.U, *b1 = A.kind == RSB_OPND_DUAL_BN_RELU ? A.U + A.ku : A.V;
            const long ld0 = A.ldu, ld1 = A.kind == RSB_OPND_DUAL_BN_RELU ? A.ldu : A.ldv;
            const bool coef_vec = (A.k0 & 3) == 0 && (A.ku & 3) == 0 && (!A.a || ((uintptr_t)A.a & 15) == 0) &&
                                  (!A.d || ((uintptr_t)A.d & 15) == 0) && (!A.b || ((uintptr_t)A.b & 15) == 0);
            const long my_tiles = n_row_tiles > blockIdx.x ? (n_row_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
            const long total = my_tiles * P.n_tiles * kc_count;

            // ---- issue side (runs RD chunks ahead) ----
            // row pointers advance by a constant per tile (no 64-bit multiplies in the loop); rows past the end are
            // zero-filled (src-size 0) from the operand's base address
            long iq = 0, i_row = (long)blockIdx.x * TM + rsub;       // i_row: this thread's first row of the tile in flight
            int i_nt = 0, i_kc = 0, i_slot = 0;
            const float *ip0[4], *ip1[4];
            const long step0 = (long)gridDim.x * TM * ld0, step1 = (long)gridDim.x * TM * ld1;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                ip0[j] = b0 + (i_row + 32 * j) * ld0;
                ip1[j] = has1 ? b1 + (i_row + 32 * j) * ld1 : nullptr;
            }
            auto issue = [&]() {
                if (iq < total) {
                    const int kl = i_kc * KC + k4 * 4;
                    const int nv = max(0, min(4, A.K - kl));
                    const int ks = nv > 0 ? A.k0 + kl : A.k0;
                    const int k_u = A.kind == RSB_OPND_AFFINE2 ? ks % A.ku : ks;
                    const uint32_t slot = ring0 + (uint32_t)i_slot * SLOT_BYTES;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const bool ok = i_row + 32 * j < P.rows;
                        const int bytes = ok ? nv * 4 : 0;
                        if (has0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(slot + (uint32_t)j * PIECE_STRIDE), "l"(ok ? ip0[j] + k_u : b0), "r"(bytes) : "memory");
                        if (has1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(slot + (uint32_t)(4 + j) * PIECE_STRIDE), "l"(ok ? ip1[j] + ks : b1), "r"(bytes) : "memory");
                    }
                    if (++i_kc == kc_count) {
                        i_kc = 0;
                        if (++i_nt == P.n_tiles) {
                            i_nt = 0;
                            i_row += (long)gridDim.x * TM;
#pragma unroll
                            for (int j = 0; j < 4; j++) { ip0[j] += step0; if (has1) ip1[j] += step1; }
                        }
                    }
                }
                cp_async_commit();
                iq++;
                if (++i_slot == RD) i_slot = 0;
            };
            for (int p = 0; p < RD; p++) issue();

            // ---- consume side ----
            // per-channel coefficients of this thread's quad: loaded once when the operand has a single K chunk
            Coef4 cf;
            auto load_coef = [&](int kl, int nv, int k) {
                if (nv == 4 && coef_vec) {
                    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 a4 = A.a ? __ldg(reinterpret_cast<const float4 *>(A.a + k)) : z4;
                    const float4 d4 = A.d ? __ldg(reinterpret_cast<const float4 *>(A.d + k)) : z4;
                    const float4 b4 = A.b ? __ldg(reinterpret_cast<const float4 *>(A.b + k)) : z4;
                    cf.a[0] = a4.x; cf.a[1] = a4.y; cf.a[2] = a4.z; cf.a[3] = a4.w;
                    cf.d[0] = d4.x; cf.d[1] = d4.y; cf.d[2] = d4.z; cf.d[3] = d4.w;
                    cf.b[0] = b4.x; cf.b[1] = b4.y; cf.b[2] = b4.z; cf.b[3] = b4.w;
                    if (A.kind == RSB_OPND_DUAL_BN_RELU) {
                        const float4 a24 = __ldg(reinterpret_cast<const float4 *>(A.a + A.ku + k));
                        const float4 d24 = __ldg(reinterpret_cast<const float4 *>(A.d + A.ku + k));
                        cf.a2[0] = a24.x; cf.a2[1] = a24.y; cf.a2[2] = a24.z; cf.a2[3] = a24.w;
                        cf.d2[0] = d24.x; cf.d2[1] = d24.y; cf.d2[2] = d24.z; cf.d2[3] = d24.w;
                    }
                } else if (nv > 0) {
                    coef_load(A, k, nv, cf);
                }
            };
            if (kc_count == 1) load_coef(k4 * 4, max(0, min(4, A.K - k4 * 4)), A.k0 + k4 * 4);
            uint32_t it = 0;
            int c_slot = 0;
            for (long tile = blockIdx.x; tile < n_row_tiles; tile += gridDim.x) {
                const int nrows = (int)min((long)TM, P.rows - tile * TM);
                for (int nt = 0; nt < P.n_tiles; nt++) {
                    for (int kc = 0; kc < kc_count; kc++, it++) {
                        const int s = it % STAGES;
                        if (RD == 3) cp_async_wait<2>(); else cp_async_wait<1>();
                        mbar_wait(&B->empty[s], ((it / STAGES) & 1) ^ 1);
                        unsigned char *st = smem + (size_t)s * stage_bytes;
                        if (tid == 0) {
                            mbar_arrive_expect_tx(&B->full[s], (uint32_t)(2 * b_bytes));
                            const float *src = P.Wp + ((size_t)nt * kc_count + kc) * (2 * (size_t)NT * KC);
                            bulk_g2s(st + 2 * A_TILE_BYTES, src, (uint32_t)(2 * b_bytes), &B->full[s]);
                        }
                        const uint32_t a_hi = rsb_smem_addr(st), a_lo = a_hi + A_TILE_BYTES;
                        const int kl = kc * KC + k4 * 4;
                        const int nv = max(0, min(4, A.K - kl));
                        const int k = A.k0 + kl;
                        if (kc_count > 1) load_coef(kl, nv, k);
                        const uint32_t slot = ring0 + (uint32_t)c_slot * SLOT_BYTES;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int r = rsub + 32 * j;
                            float v[4] = {0.f, 0.f, 0.f, 0.f};
                            if (nv > 0 && r < nrows) {
                                float u[4], w[4];
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(u[0]), "=f"(u[1]), "=f"(u[2]), "=f"(u[3]) : "r"(slot + (uint32_t)j * PIECE_STRIDE));
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w[0]), "=f"(w[1]), "=f"(w[2]), "=f"(w[3]) : "r"(slot + (uint32_t)(4 + j) * PIECE_STRIDE));
                                long g = 0;
                                int sidx = 0;
                                if (A.kind == RSB_OPND_POOLED) { const long row = tile * TM + r; g = row / A.ns; sidx = (int)(row - g * A.ns); }
                                opnd_apply4(A, cf, u, w, k, nv, g, sidx, v);
                            }
                            float4 hi, lo;
                            split4(v, hi, lo);
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + st_off[j]), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + st_off[j]), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w) : "memory");
                        }
                        fence_proxy_async();
                        mbar_arrive(&B->full[s]);
                        issue();
                        if (++c_slot == RD) c_slot = 0;
                    }
                }
            }
            cp_async_wait<0>();
        } else {
        // ===== producers: two groups of 4 warps take alternate chunks, so two stages are being loaded at any time.
        // Inside a group: thread = (channel quad k4, row sub-index), 8 rows per thread per chunk, coefficients hoisted.
        const int grp = warp >> 2, gt = tid & (GROUP_THREADS - 1);
        const int k4 = (gt >> 3) & 7, rsub = (gt >> 6) * 8 + (gt & 7);   // rows rsub + 16*j; quarter-warp = one core matrix
        const OpndFlags F = opnd_flags(A);
        uint32_t it = 0;
        for (long tile = blockIdx.x; tile < n_row_tiles; tile += gridDim.x) {
            const long row_base = tile * TM;
            for (int nt = 0; nt < P.n_tiles; nt++) {
                for (int kc = 0; kc < kc_count; kc++, it++) {
                    if ((it & 1) != (uint32_t)grp) continue;
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&B->empty[s], ph ^ 1);
                    unsigned char *st = smem + (size_t)s * stage_bytes;
                    if (gt == 0) {
                        mbar_arrive_expect_tx(&B->full[s], (uint32_t)(2 * b_bytes));
                        const float *src = P.Wp + ((size_t)nt * kc_count + kc) * (2 * (size_t)NT * KC);
                        bulk_g2s(st + 2 * A_TILE_BYTES, src, (uint32_t)(2 * b_bytes), &B->full[s]);
                    }
                    float *a_hi = reinterpret_cast<float *>(st);
                    float *a_lo = reinterpret_cast<float *>(st + A_TILE_BYTES);
                    const int kl = kc * KC + k4 * 4;                 // logical channel (without k0)
                    const int nv = min(4, A.K - kl);                 // valid channels of this quad (<= 0: none)
                    Coef4 cf;
                    if (nv > 0) coef_load(A, A.k0 + kl, nv, cf);
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        float v[4][4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const long row = row_base + rsub + 16 * (half * 4 + j);
                            if (nv > 0 && row < P.rows) {
                                long g = 0;
                                int sidx = 0;
                                if (A.kind == RSB_OPND_POOLED) { g = row / A.ns; sidx = (int)(row - g * A.ns); }
                                opnd_eval4(A, F, cf, row, A.k0 + kl, nv, g, sidx, v[j]);
                            } else {
                                v[j][0] = v[j][1] = v[j][2] = v[j][3] = 0.f;
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int r = rsub + 16 * (half * 4 + j);
                            float4 hi, lo;
                            split4(v[j], hi, lo);
                            const int off = (r >> 3) * (KC / 4) * 32 + k4 * 32 + (r & 7) * 4;   // floats
                            *reinterpret_cast<float4 *>(a_hi + off) = hi;
                            *reinterpret_cast<float4 *>(a_lo + off) = lo;
                        }
                    }
                    fence_proxy_async();
                    mbar_arrive(&B->full[s]);
                }
            }
        }
        }
    } else if (warp == MMA_WARP) {
        // =============================== MMA issuer ===============================
        const uint32_t idesc = umma_idesc_tf32(TM, NT, false);
        uint32_t it = 0, acc_it = 0;
        for (long tile = blockIdx.x; tile < n_row_tiles; tile += gridDim.x) {
            for (int nt = 0; nt < P.n_tiles; nt++, acc_it++) {
                const int ab = acc_it & 1;
                mbar_wait(&B->acc_empty[ab], ((acc_it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(ab * NT_MAX);
                for (int kc = 0; kc < kc_count; kc++, it++) {
                    const int s = it % STAGES;
                    mbar_wait(&B->full[s], (it / STAGES) & 1);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t st = rsb_smem_addr(smem + (size_t)s * stage_bytes);
                        const uint32_t a_hi = st, a_lo = st + A_TILE_BYTES;
                        const uint32_t b_hi = st + 2 * A_TILE_BYTES, b_lo = b_hi + (uint32_t)b_bytes;
                        const uint32_t SBO = (KC / 4) * 128, LBO = 128;
#pragma unroll
                        for (int ks = 0; ks < KC / 8; ks++) {
                            const uint32_t koff = ks * 2 * 128;  // two core matrices (8 tf32) along K per UMMA
                            const uint64_t dah = umma_desc(a_hi + koff, LBO, SBO), dal = umma_desc(a_lo + koff, LBO, SBO);
                            const uint64_t dbh = umma_desc(b_hi + koff, LBO, SBO), dbl = umma_desc(b_lo + koff, LBO, SBO);
                            umma_tf32(tmem_d, dal, dbh, idesc, (kc | ks) ? 1u : 0u);   // small terms first
                            umma_tf32(tmem_d, dah, dbl, idesc, 1u);
                            umma_tf32(tmem_d, dah, dbh, idesc, 1u);
                        }
                        umma_commit(&B->empty[s]);
                        if (kc == kc_count - 1) umma_commit(&B->acc_full[ab]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // =============================== epilogue ===============================
        const int q = warp & 3;                 // TMEM lane quadrant this warp may access
        const int r = q * 32 + lane;
        float *my_tile = stat_tile + (warp - MMA_WARP - 1) * 32 * 33;
        // forward statistics live in registers for the whole kernel (column n0 + lane of every 32-column block of
        // every N tile this thread sees); one fp64 atomic per (CTA warp, column) at the end instead of per tile
        constexpr int ACC_BLOCKS = NT_MAX / 32;
        double acc1[ACC_BLOCKS], acc2[ACC_BLOCKS];
#pragma unroll
        for (int i = 0; i < ACC_BLOCKS; i++) acc1[i] = acc2[i] = 0.0;
        const bool reg_stats = E.stats && E.kind == RSB_EPI_BIAS_STATS && P.n_tiles == 1;
        uint32_t acc_it = 0;
        for (long tile = blockIdx.x; tile < n_row_tiles; tile += gridDim.x) {
            const long row = tile * TM + r;
            const bool row_ok = row < P.rows;
            for (int nt = 0; nt < P.n_tiles; nt++, acc_it++) {
                const int ab = acc_it & 1;
                mbar_wait(&B->acc_full[ab], (acc_it >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * NT_MAX);
                const int ncols = min(NT, P.N - nt * NT);
                for (int c0 = 0; c0 < ncols; c0 += 32) {
                    float v[32];
                    tmem_ld32(taddr + c0, v);
                    const int n0 = nt * NT + c0;
                    const float *yl = E.kind == RSB_EPI_RELU_MASK ? E.Yl + (size_t)(row_ok ? row : 0) * E.ldl : nullptr;
                    if (E.kind == RSB_EPI_BIAS_STATS) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const int n = n0 + j;
                            const float y = v[j] + ((E.bias && n < P.N) ? __ldg(E.bias + n) : 0.f);
                            v[j] = (row_ok && n < P.N) ? y : 0.f;
                        }
                    } else {
                        // dgrad: ReLU mask of the layer below, recomputed from its stored pre-BN output
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const int n = n0 + j;
                            float out = 0.f;
                            if (row_ok && n < P.N) {
                                float z = fmaf(__ldg(yl + n), __ldg(E.sc + n), __ldg(E.sh + n));
                                if (E.dual) z += fmaf(__ldg(yl + P.N + n), __ldg(E.sc + P.N + n), __ldg(E.sh + P.N + n));
                                out = z > 0.f ? v[j] : 0.f;
                            }
                            v[j] = out;
                        }
                    }
                    if (E.kind == RSB_EPI_BIAS_STATS) {
                        // Transpose the warp's 32x32 block through shared memory: lanes become COLUMNS, so every
                        // store instruction writes one contiguous 128-byte row segment (the TMEM layout gives each
                        // lane a whole row, which would scatter 32 partial lines per instruction), and the column
                        // statistics fall out of the same pass.
#pragma unroll
                        for (int j = 0; j < 32; j++) my_tile[lane * 33 + j] = v[j];
                        __syncwarp();
                        const int n = n0 + lane;
                        const long rb = tile * TM + q * 32;
                        float s1 = 0.f, s2 = 0.f;
                        const int nrows = (E.Y && n < P.N) ? (int)max(0L, min(32L, P.rows - rb)) : 0;
                        float *ycol = E.Y + (size_t)rb * E.ldy + n;
                        const float *tcol = my_tile + lane;
                        if (nrows == 32) {
#pragma unroll
                            for (int i = 0; i < 32; i++) {
                                const float t = tcol[i * 33];
                                s1 += t;
                                s2 = fmaf(t, t, s2);
                                *ycol = t;
                                ycol += E.ldy;
                            }
                        } else {
#pragma unroll 8
                            for (int i = 0; i < 32; i++) {
                                const float t = tcol[i * 33];
                                s1 += t;
                                s2 = fmaf(t, t, s2);
                                if (i < nrows) ycol[(size_t)i * E.ldy] = t;
                            }
                        }
                        __syncwarp();
                        if (E.stats) {
                            if (reg_stats) {
#pragma unroll
                                for (int bi = 0; bi < ACC_BLOCKS; bi++)
                                    if (bi == (c0 >> 5)) { acc1[bi] += (double)s1; acc2[bi] += (double)s2; }
                            } else if (n < P.N) {
                                atomicAdd(E.stats + n, (double)s1);
                                atomicAdd(E.stats + P.N + n, (double)s2);
                            }
                        }
                        continue;
                    }
                    if (row_ok && E.Y) {
                        float *yrow = E.Y + (size_t)row * E.ldy + n0;
                        if (n0 + 32 <= P.N && ((uintptr_t)yrow & 15) == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4 *>(yrow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
                            for (int j = 0; j < 32 && n0 + j < P.N; j++) yrow[j] = v[j];
                        }
                    }
                    if (E.stats) {
                        // column sums over this warp's 32 rows through a padded shared tile, then fp64 atomics
                        const int n = n0 + lane;
                        const float s0 = warp_colsum(my_tile, v, lane);
                        if (reg_stats) {
#pragma unroll
                            for (int j = 0; j < 32; j++) my_tile[lane * 33 + j] = v[j] * v[j];
                            __syncwarp();
                            float sq = 0.f;
#pragma unroll
                            for (int i = 0; i < 32; i++) sq += my_tile[i * 33 + lane];
                            __syncwarp();
#pragma unroll
                            for (int bi = 0; bi < ACC_BLOCKS; bi++)
                                if (bi == (c0 >> 5)) { acc1[bi] += (double)s0; acc2[bi] += (double)sq; }
                            continue;
                        }
                        if (n < P.N) atomicAdd(E.stats + n, (double)s0);
                        if (E.kind == RSB_EPI_BIAS_STATS) {
#pragma unroll
                            for (int j = 0; j < 32; j++) my_tile[lane * 33 + j] = v[j] * v[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j++) {
                                const int nn = n0 + j;
                                const float h = (row_ok && nn < P.N) ? (__ldg(yl + nn) - __ldg(E.mu + nn)) * __ldg(E.inv + nn) : 0.f;
                                my_tile[lane * 33 + j] = v[j] * h;
                            }
                        }
                        __syncwarp();
                        float s1 = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; i++) s1 += my_tile[i * 33 + lane];
                        __syncwarp();
                        if (n < P.N) atomicAdd(E.stats + P.N + n, (double)s1);
                        if (E.kind == RSB_EPI_RELU_MASK && E.dual) {
#pragma unroll
                            for (int j = 0; j < 32; j++) {
                                const int nn = n0 + j;
                                const float h = (row_ok && nn < P.N) ? (__ldg(yl + P.N + nn) - __ldg(E.mu + P.N + nn)) * __ldg(E.inv + P.N + nn) : 0.f;
                                my_tile[lane * 33 + j] = v[j] * h;
                            }
                            __syncwarp();
                            float s2 = 0.f;
#pragma unroll
                            for (int i = 0; i < 32; i++) s2 += my_tile[i * 33 + lane];
                            __syncwarp();
                            if (n < P.N) atomicAdd(E.stats + 2 * P.N + n, (double)s2);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&B->acc_empty[ab]);
            }
        }
        if (reg_stats) {
#pragma unroll
            for (int bi = 0; bi < ACC_BLOCKS; bi++) {
                const int n = bi * 32 + lane;
                if (n < P.N && (acc1[bi] != 0.0 || acc2[bi] != 0.0)) {
                    atomicAdd(E.stats + n, acc1[bi]);
                    atomicAdd(E.stats + P.N + n, acc2[bi]);
                }
            }
        }
    }
    cta_epilogue(tmem_base, warp);
}

// ============================================================================================================
// dW[m, n] += sum_r G(r, m) * X(r, n)   (weight gradient; reduction over rows)
// ============================================================================================================
// ---- wgrad staging -------------------------------------------------------------------------------------------
// The TRANSPOSE of a [32 rows x width channels] slab is staged as a K-major tile (channels = tile rows, the 32
// reduction rows = K): lanes run along channels (coalesced global reads of one slab row), each thread packs 4
// consecutive reduction rows of its channel into one 16 B core-matrix row.  Same canonical layout and descriptors
// as gemm_rows (LBO 128 B, SBO 1024 B).  A "unit" = (operand, 32-channel block, row quad); a group of 4 warps
// stages one chunk: warp w4 owns the row quads w4 and w4+4.  Units are processed in batches whose global loads
// are all issued before any is consumed (the kernel is load-latency bound, see profiles/).
struct WUnit {
    float u[4], w[4];        // raw pieces of the 4 rows (first / second tensor)
    float dz[4];             // POOLED: selected pooled gradient per row
    float a, b, d, a2, d2;   // channel coefficients
    int ct, k4, which;       // channel inside the tile, row quad, operand (0 = G, 1 = X)
    bool live;               // channel valid
};

__device__ __forceinline__ void wunit_load(WUnit &U, const Opnd &O, int c_base, int cb, int k4, long row0, long rows, int lane)
{
    U.k4 = k4;
    U.ct = cb + lane;
    const int c = c_base + U.ct;
    U.live = c < O.K;
#pragma unroll
    for (int e = 0; e < 4; e++) { U.u[e] = 0.f; U.w[e] = 0.f; U.dz[e] = 0.f; }
    U.a = U.b = U.d = U.a2 = U.d2 = 0.f;
    if (!U.live) return;
    const int k = O.k0 + c;
    const long r0 = row0 + k4 * 4;
    const int nr = (int)min(4L, rows - r0);          // valid rows of this quad (may be <= 0)
    if (O.a) U.a = __ldg(O.a + k);
    if (O.d) U.d = __ldg(O.d + k);
    if (O.b) U.b = __ldg(O.b + k);
    switch (O.kind) {
    case RSB_OPND_RAW:
    case RSB_OPND_BN_RELU: {
        const float *pu = O.U + (size_t)r0 * O.ldu + k;
#pragma unroll
        for (int e = 0; e < 4; e++) if (e < nr) U.u[e] = __ldg(pu + (size_t)e * O.ldu);
        break;
    }
    case RSB_OPND_DUAL_BN_RELU: {
        U.a2 = __ldg(O.a + O.ku + k);
        U.d2 = __ldg(O.d + O.ku + k);
        const float *pu = O.U + (size_t)r0 * O.ldu + k;
#pragma unroll
        for (int e = 0; e < 4; e++) if (e < nr) { U.u[e] = __ldg(pu + (size_t)e * O.ldu); U.w[e] = __ldg(pu + (size_t)e * O.ldu + O.ku); }
        break;
    }
    case RSB_OPND_AFFINE2: {
        const float *pu = O.U + (size_t)r0 * O.ldu + (k % O.ku);
        const float *pv = O.V + (size_t)r0 * O.ldv + k;
#pragma unroll
        for (int e = 0; e < 4; e++) if (e < nr) { U.u[e] = __ldg(pu + (size_t)e * O.ldu); U.w[e] = __ldg(pv + (size_t)e * O.ldv); }
        break;
    }
    default: {
        const float *pv = O.V + (size_t)r0 * O.ldv + k;
        // rows of a quad share their pooling group whenever ns % 4 == 0 (r0 is a multiple of 4)
        const bool same = (O.ns & 3) == 0;
        const int g0 = (int)(r0 / O.ns), s0 = (int)(r0 - (long)g0 * O.ns);
        int arg0 = 0;
        float dz0 = 0.f;
        if (same && nr > 0) { arg0 = __ldg(O.arg + (size_t)g0 * O.ldu + k); dz0 = __ldg(O.U + (size_t)g0 * O.ldu + k); }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (e >= nr) continue;
            U.w[e] = __ldg(pv + (size_t)e * O.ldv);
            if (same) {
                U.u[e] = __int_as_float(arg0 == s0 + e ? 1 : 0);
                U.dz[e] = dz0;
            } else {
                const long r = r0 + e;
                const long g = r / O.ns;
                U.u[e] = __int_as_float(__ldg(O.arg + (size_t)g * O.ldu + k) == (int)(r - g * O.ns) ? 1 : 0);
                U.dz[e] = __ldg(O.U + (size_t)g * O.ldu + k);
            }
        }
        break;
    }
    }
}

__device__ __forceinline__ void wunit_store(const WUnit &U, const Opnd &O, long row0, long rows, float *hi_base, float *lo_base)
{
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (U.live) {
        const long r0 = row0 + U.k4 * 4;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (r0 + e >= rows) continue;
            switch (O.kind) {
            case RSB_OPND_RAW: v[e] = U.u[e]; break;
            case RSB_OPND_BN_RELU: v[e] = rsb_relu(fmaf(U.u[e], U.a, U.d)); break;
            case RSB_OPND_DUAL_BN_RELU: v[e] = rsb_relu(fmaf(U.u[e], U.a, U.d) + fmaf(U.w[e], U.a2, U.d2)); break;
            case RSB_OPND_AFFINE2: v[e] = fmaf(U.a, U.u[e], fmaf(U.b, U.w[e], U.d)); break;
            default: v[e] = fmaf(U.a, __float_as_int(U.u[e]) ? U.dz[e] : 0.f, fmaf(U.b, U.w[e], U.d)); break;
            }
        }
    }
    float4 hi, lo;
    split4(v, hi, lo);
    const int off = (U.ct >> 3) * (KC / 4) * 32 + U.k4 * 32 + (U.ct & 7) * 4;   // floats
    *reinterpret_cast<float4 *>(hi_base + off) = hi;
    *reinterpret_cast<float4 *>(lo_base + off) = lo;
}

__global__ void __launch_bounds__(THREADS, 1) gemm_wgrad_kernel(const __grid_constant__ WgradParams P)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NT = P.NT;
    const size_t a_bytes = A_TILE_BYTES;                  // one of hi / lo of the G^T tile (128 channels x 32 rows)
    const size_t b_bytes = (size_t)NT * KC * 4;           // one of hi / lo of the X^T tile (NT channels x 32 rows)
    const size_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const int STAGES = P.stages;
    Barriers *B = reinterpret_cast<Barriers *>(smem + STAGES * stage_bytes);
    // single-tile problems never restage a channel block that holds no valid channel: zero it once
    const bool skip_empty = P.m_tiles == 1 && P.n_tiles == 1;
    if (skip_empty)
        for (size_t i = tid; i < STAGES * stage_bytes / 16; i += THREADS) reinterpret_cast<float4 *>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
    const uint32_t tmem_base = cta_prologue(B, tid, warp, P.raw_slots > 0 ? PROD_THREADS + 1 : GROUP_THREADS + 1);

    const long n_chunks = (P.rows + KC - 1) / KC;
    // chunks of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const long my_chunks = n_chunks > blockIdx.x ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (my_chunks > 0) {
        if (warp < PROD_WARPS && P.raw_slots > 0) {
            // ===== producers, asynchronous path: the raw [32 rows x channels] slabs of both operands are copied
            // with 16-byte cp.async (a warp per slab row: fully coalesced) into a ring of raw slabs, RD chunks
            // ahead; the transposing transform then reads shared memory only.
            const int RS = P.raw_slots, RD = RS - 1;             // RD chunks in flight, one slab being read
            // per-channel coefficients of both operand tiles live in shared memory for a whole (mt, nt) pass:
            // [G: a, b, d, a2, d2 | X: a, b, d, a2, d2] x (TM + NT_MAX) channels, filled cooperatively below
            float *ctab = reinterpret_cast<float *>(B + 1);
            int *itab = reinterpret_cast<int *>(ctab + 5 * (TM + NT_MAX));        // AFFINE2 wrap: U index per channel
            unsigned char *ring = reinterpret_cast<unsigned char *>(itab + (TM + NT_MAX));
            struct Region { const float *base; long ld; int quads, last_bytes, off, shift; };   // off: float4 index in the slab
            Region Rg[4];
            int idxmod_g = 0, idxmod_x = 0;                        // AFFINE2 wrap: U index = (k) % ku
            auto setup = [&](const Opnd &O, int c_base, int width, Region &ra, Region &rb, int &off, int &idxmod) {
                const int nvalid = max(0, min(width, O.K - c_base));
                const int quads = (nvalid + 3) / 4, last = (nvalid - 4 * (quads - 1)) * 4;
                const int k = O.k0 + c_base;
                ra.base = rb.base = nullptr; ra.quads = rb.quads = 0; ra.ld = rb.ld = 0; ra.last_bytes = rb.last_bytes = 16; ra.off = rb.off = 0;
                idxmod = 0;
                if (O.kind == RSB_OPND_AFFINE2) {
                    const int kk0 = k % O.ku;
                    const bool wrap = kk0 + nvalid > O.ku;
                    const int ustart = wrap ? 0 : kk0, uw = wrap ? O.ku : nvalid;
                    ra.base = O.U + ustart; ra.ld = O.ldu; ra.quads = (uw + 3) / 4; ra.last_bytes = (uw - 4 * (ra.quads - 1)) * 4;
                    idxmod = wrap ? O.ku : 0;
                    rb.base = O.V + k; rb.ld = O.ldv; rb.quads = quads; rb.last_bytes = last;
                } else {
                    ra.base = O.U + k; ra.ld = O.ldu; ra.quads = quads; ra.last_bytes = last;
                    if (O.kind == RSB_OPND_DUAL_BN_RELU) { rb.base = O.U + O.ku + k; rb.ld = O.ldu; rb.quads = quads; rb.last_bytes = last; }
                }
                ra.off = off; off += KC * ra.quads;
                rb.off = off; off += KC * rb.quads;
                // lanes of a warp cover (row, quad) pairs: 2^shift lanes per slab row, 32 >> shift rows per pass, so a
                // narrow slab (8 quads = 32 channels) still issues full-width cp.async instructions
                auto lanes_per_row = [](int quads) { int sh = 0; while ((1 << sh) < quads && sh < 5) sh++; return sh; };
                ra.shift = lanes_per_row(ra.quads);
                rb.shift = lanes_per_row(rb.quads);
            };
            auto copy_chunk = [&](long ci, int slot) {
                if (ci < my_chunks) {
                    const long row0 = (blockIdx.x + ci * gridDim.x) * (long)KC;
                    const uint32_t sb = rsb_smem_addr(ring + (size_t)slot * P.raw_slot_bytes);
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const Region &R = Rg[r];
                        if (R.quads == 0) continue;
                        const int rpp = 32 >> R.shift, qpl = 1 << R.shift;
                        for (int rr = warp * rpp + (lane >> R.shift); rr < KC; rr += PROD_WARPS * rpp) {
                            const long row = row0 + rr;
                            const bool ok = row < P.rows;
                            const float *src = R.base + (ok ? row : 0) * R.ld;
                            for (int q = lane & (qpl - 1); q < R.quads; q += qpl) {
                                const int bytes = ok ? (q == R.quads - 1 ? R.last_bytes : 16) : 0;
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sb + (uint32_t)(R.off + rr * R.quads + q) * 16), "l"(src + q * 4), "r"(bytes) : "memory");
                            }
                        }
                    }
                }
                cp_async_commit();
            };
            auto fill_tab = [&](const Opnd &O, int c_base, int width, int idxmod, int t0) {
                for (int ct = tid; ct < width; ct += PROD_THREADS) {
                    const int k = O.k0 + c_base + ct;
                    const bool ok = c_base + ct < O.K;
                    ctab[0 * (TM + NT_MAX) + t0 + ct] = (ok && O.a) ? __ldg(O.a + k) : 0.f;
                    ctab[1 * (TM + NT_MAX) + t0 + ct] = (ok && O.b) ? __ldg(O.b + k) : 0.f;
                    ctab[2 * (TM + NT_MAX) + t0 + ct] = (ok && O.d) ? __ldg(O.d + k) : 0.f;
                    ctab[3 * (TM + NT_MAX) + t0 + ct] = (ok && O.kind == RSB_OPND_DUAL_BN_RELU) ? __ldg(O.a + O.ku + k) : 0.f;
                    ctab[4 * (TM + NT_MAX) + t0 + ct] = (ok && O.kind == RSB_OPND_DUAL_BN_RELU) ? __ldg(O.d + O.ku + k) : 0.f;
                    itab[t0 + ct] = idxmod ? (k % idxmod) : ct;
                }
            };
            // lim = rows of this thread's quad that exist (ragged last chunk): the rest is staged as zeros, whatever the
            // transform makes of a zero-filled slab row
            auto stage = [&](const Opnd &O, const Region &ra, const Region &rb, int t0, int c_base, int width,
                             const float *slab, float *hi_base, float *lo_base, int lim) {
                const int k4 = warp;                                  // PROD_WARPS == KC / 4 row quads
                const int nvalid = max(0, min(width, O.K - c_base));
                const int nb = skip_empty ? (nvalid + 31) / 32 : (width + 31) / 32;
                for (int cbk = 0; cbk < nb; cbk++) {
                    const int ct = cbk * 32 + lane;
                    if (ct >= width) break;
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    if (ct < nvalid) {
                        const float a = ctab[0 * (TM + NT_MAX) + t0 + ct], b = ctab[1 * (TM + NT_MAX) + t0 + ct];
                        const float d = ctab[2 * (TM + NT_MAX) + t0 + ct];
                        const float a2 = ctab[3 * (TM + NT_MAX) + t0 + ct], d2 = ctab[4 * (TM + NT_MAX) + t0 + ct];
                        const int iu = itab[t0 + ct];
                        const float *pa = slab + (size_t)ra.off * 4 + (size_t)(k4 * 4) * ra.quads * 4 + iu;
                        const float *pb = slab + (size_t)rb.off * 4 + (size_t)(k4 * 4) * rb.quads * 4 + ct;
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const float u = pa[(size_t)e * ra.quads * 4];
                            const float w = rb.quads ? pb[(size_t)e * rb.quads * 4] : 0.f;
                            switch (O.kind) {
                            case RSB_OPND_RAW: v[e] = u; break;
                            case RSB_OPND_BN_RELU: v[e] = rsb_relu(fmaf(u, a, d)); break;
                            case RSB_OPND_DUAL_BN_RELU: v[e] = rsb_relu(fmaf(u, a, d) + fmaf(w, a2, d2)); break;
                            default: v[e] = fmaf(a, u, fmaf(b, w, d)); break;
                            }
                            if (e >= lim) v[e] = 0.f;
                        }
                    }
                    float4 hi, lo;
                    split4(v, hi, lo);
                    const int off = (ct >> 3) * (KC / 4) * 32 + k4 * 32 + (ct & 7) * 4;   // floats
                    *reinterpret_cast<float4 *>(hi_base + off) = hi;
                    *reinterpret_cast<float4 *>(lo_base + off) = lo;
                }
            };
            uint32_t it = 0;
            for (int mt = 0; mt < P.m_tiles; mt++)
                for (int nt = 0; nt < P.n_tiles; nt++) {
                    int off = 0;
                    setup(P.G, mt * TM, TM, Rg[0], Rg[1], off, idxmod_g);
                    setup(P.X, nt * NT, NT, Rg[2], Rg[3], off, idxmod_x);
                    fill_tab(P.G, mt * TM, TM, idxmod_g, 0);
                    fill_tab(P.X, nt * NT, NT, idxmod_x, TM);
                    asm volatile("bar.sync 1, %0;" ::"n"(PROD_THREADS) : "memory");
                    int islot = 0, cslot = 0;
                    for (int p = 0; p < RD; p++) { copy_chunk(p, islot); if (++islot == RS) islot = 0; }
                    for (long ci = 0; ci < my_chunks; ci++, it++) {
                        switch (RD) {   // wait until at most RD-1 younger slabs are still in flight
                        case 7: cp_async_wait<6>(); break;
                        case 6: cp_async_wait<5>(); break;
                        case 5: cp_async_wait<4>(); break;
                        case 4: cp_async_wait<3>(); break;
                        case 3: cp_async_wait<2>(); break;
                        case 2: cp_async_wait<1>(); break;
                        default: cp_async_wait<0>(); break;
                        }
                        asm volatile("bar.sync 1, %0;" ::"n"(PROD_THREADS) : "memory");   // slab ci complete; slab ci-1 no longer read
                        copy_chunk(ci + RD, islot);
                        if (++islot == RS) islot = 0;
                        const int s = it % STAGES;
                        mbar_wait(&B->empty[s], ((it / STAGES) & 1) ^ 1);
                        unsigned char *st = smem + (size_t)s * stage_bytes;
                        const float *slab = reinterpret_cast<const float *>(ring + (size_t)cslot * P.raw_slot_bytes);
                        const int lim = (int)min((long)KC, P.rows - (blockIdx.x + ci * gridDim.x) * (long)KC) - warp * 4;
                        stage(P.G, Rg[0], Rg[1], 0, mt * TM, TM, slab, reinterpret_cast<float *>(st), reinterpret_cast<float *>(st + a_bytes), lim);
                        stage(P.X, Rg[2], Rg[3], TM, nt * NT, NT, slab, reinterpret_cast<float *>(st + 2 * a_bytes),
                              reinterpret_cast<float *>(st + 2 * a_bytes + b_bytes), lim);
                        fence_proxy_async();
                        mbar_arrive(&B->full[s]);
                        if (tid == 0) mbar_arrive(&B->full[s]);
                        if (++cslot == RS) cslot = 0;
                    }
                    cp_async_wait<0>();
                    asm volatile("bar.sync 1, %0;" ::"n"(PROD_THREADS) : "memory");
                }
        } else if (warp < PROD_WARPS) {
            const int grp = warp >> 2, w4 = warp & 3;
            uint32_t it = 0;
            for (int mt = 0; mt < P.m_tiles; mt++)
                for (int nt = 0; nt < P.n_tiles; nt++)
                    for (long ci = 0; ci < my_chunks; ci++, it++) {
                        if ((it & 1) != (uint32_t)grp) continue;
                        const int s = it % STAGES;
                        mbar_wait(&B->empty[s], ((it / STAGES) & 1) ^ 1);
                        unsigned char *st = smem + (size_t)s * stage_bytes;
                        const long row0 = (blockIdx.x + ci * gridDim.x) * (long)KC;
                        // units of this warp for this chunk: 2 row quads x (channel blocks of G + channel blocks of X)
                        const int vg = min(TM, P.M - mt * TM), vx = min(NT, P.N - nt * NT);
                        const int nbg = skip_empty ? (vg + 31) / 32 : (TM + 31) / 32;
                        const int nbx = skip_empty ? (vx + 31) / 32 : (NT + 31) / 32;
                        const int per_h = nbg + nbx, n_units = 2 * per_h;
                        float *g_hi = reinterpret_cast<float *>(st), *g_lo = reinterpret_cast<float *>(st + a_bytes);
                        float *x_hi = reinterpret_cast<float *>(st + 2 * a_bytes), *x_lo = reinterpret_cast<float *>(st + 2 * a_bytes + b_bytes);
                        constexpr int UB = 4;
                        for (int u0 = 0; u0 < n_units; u0 += UB) {
                            WUnit U[UB];
#pragma unroll
                            for (int i = 0; i < UB; i++) {
                                const int u = u0 + i;
                                U[i].which = -1;
                                if (u < n_units) {
                                    const int h = u / per_h, rem = u - h * per_h;
                                    const bool isg = rem < nbg;
                                    const int cb = (isg ? rem : rem - nbg) * 32;
                                    if (cb + lane < (isg ? TM : NT)) {   // tiles are multiples of 16 channels, not of 32
                                        U[i].which = isg ? 0 : 1;
                                        wunit_load(U[i], isg ? P.G : P.X, isg ? mt * TM : nt * NT, cb, w4 + 4 * h, row0, P.rows, lane);
                                    }
                                }
                            }
#pragma unroll
                            for (int i = 0; i < UB; i++) {
                                if (U[i].which == 0) wunit_store(U[i], P.G, row0, P.rows, g_hi, g_lo);
                                else if (U[i].which == 1) wunit_store(U[i], P.X, row0, P.rows, x_hi, x_lo);
                            }
                        }
                        fence_proxy_async();
                        mbar_arrive(&B->full[s]);
                        if ((tid & (GROUP_THREADS - 1)) == 0) mbar_arrive(&B->full[s]);   // barrier counts GROUP_THREADS + 1
                    }
        } else if (warp == MMA_WARP) {
            const uint32_t idesc = umma_idesc_tf32(TM, NT, false);
            uint32_t it = 0, acc_it = 0;
            for (int mt = 0; mt < P.m_tiles; mt++)
                for (int nt = 0; nt < P.n_tiles; nt++, acc_it++) {
                    const int ab = acc_it & 1;
                    mbar_wait(&B->acc_empty[ab], ((acc_it >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(ab * NT_MAX);
                    for (long ci = 0; ci < my_chunks; ci++, it++) {
                        const int s = it % STAGES;
                        mbar_wait(&B->full[s], (it / STAGES) & 1);
                        tc_fence_after();
                        if (lane == 0) {
                            const uint32_t st = rsb_smem_addr(smem + (size_t)s * stage_bytes);
                            const uint32_t a_hi = st, a_lo = st + (uint32_t)a_bytes;
                            const uint32_t b_hi = st + 2 * (uint32_t)a_bytes, b_lo = b_hi + (uint32_t)b_bytes;
#pragma unroll
                            for (int ks = 0; ks < KC / 8; ks++) {
                                const uint32_t koff = ks * 2 * 128, SBO = (KC / 4) * 128, LBO = 128;
                                const uint64_t dah = umma_desc(a_hi + koff, LBO, SBO), dal = umma_desc(a_lo + koff, LBO, SBO);
                                const uint64_t dbh = umma_desc(b_hi + koff, LBO, SBO), dbl = umma_desc(b_lo + koff, LBO, SBO);
                                umma_tf32(tmem_d, dal, dbh, idesc, (ci | ks) ? 1u : 0u);
                                umma_tf32(tmem_d, dah, dbl, idesc, 1u);
                                umma_tf32(tmem_d, dah, dbh, idesc, 1u);
                            }
                            umma_commit(&B->empty[s]);
                            if (ci == my_chunks - 1) umma_commit(&B->acc_full[ab]);
                        }
                        __syncwarp();
                    }
                }
        } else {
            const int q = warp & 3;
            const int r = q * 32 + lane;            // output channel inside the M tile
            uint32_t acc_it = 0;
            for (int mt = 0; mt < P.m_tiles; mt++)
                for (int nt = 0; nt < P.n_tiles; nt++, acc_it++) {
                    const int ab = acc_it & 1;
                    mbar_wait(&B->acc_full[ab], (acc_it >> 1) & 1);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * NT_MAX);
                    const int m = mt * TM + r;
                    const int ncols = min(NT, P.N - nt * NT);
                    for (int c0 = 0; c0 < ncols; c0 += 32) {
                        float v[32];
                        tmem_ld32(taddr + c0, v);
                        if (m < P.M) {
                            float *dst = P.dW + (size_t)m * P.ldw + nt * NT + c0;
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                if (c0 + j < ncols) atomicAdd(dst + j, v[j]);
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(&B->acc_empty[ab]);
                }
        }
    }
    cta_epilogue(tmem_base, warp);
}

// W [N, K] (row-major, optionally transposed source) -> pre-split canonical K-major chunks
__global__ void weight_prep_kernel(const float *__restrict__ W, int N, int K, int ldw, int transposed, int NT,
                                   int n_tiles, int k_chunks, float *__restrict__ Wp)
{
    const long total = (long)n_tiles * k_chunks * NT * KC;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int kk = (int)(i % KC);
        const int nn = (int)((i / KC) % NT);
        const int kc = (int)((i / ((long)KC * NT)) % k_chunks);
        const int nt = (int)(i / ((long)KC * NT * k_chunks));
        const int n = nt * NT + nn, k = kc * KC + kk;
        float w = 0.f;
        if (n < N && k < K) w = transposed ? W[(size_t)k * ldw + n] : W[(size_t)n * ldw + k];
        const float hi = to_tf32(w), lo = to_tf32(w - hi);
        // canonical K-major no-swizzle: core matrix (nn/8, kk/4) at (nn/8)*SBO + (kk/4)*128 B, row (nn%8)*16 B
        const size_t off = (size_t)(nn >> 3) * ((KC / 4) * 32) + (size_t)(kk >> 2) * 32 + (nn & 7) * 4 + (kk & 3);
        float *blk = Wp + ((size_t)nt * k_chunks + kc) * (2 * (size_t)NT * KC);
        blk[off] = hi;
        blk[(size_t)NT * KC + off] = lo;
    }
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
// N tile width: the fewest tiles of <= 256 columns, balanced (N = 272 -> 160 + 112, not 256 + 16) and, when there are
// several, a multiple of 32 so that the 32-column store boxes of one tile never reach into the next; the same rule
// lays out the pre-split weights (weight_prep_kernel) and is applied by both kernel generations
inline int pick_nt(int N) { const int t = (N + NT_MAX - 1) / NT_MAX; return round_up((N + t - 1) / t, t > 1 ? 32 : 16); }

int check_opnd(const Opnd &O, const char *what)
{
    if (O.K < 1 || O.kind < 0 || O.kind > RSB_OPND_GATHER) { rsb_set_error("%s: bad operand descriptor", what); return 1; }
    if (O.kind == RSB_OPND_GATHER) return 0;    // checked by the TMA-fed launcher, the only implementation
    if (!O.U) { rsb_set_error("%s: operand U is null", what); return 1; }
    if (O.kind != RSB_OPND_RAW && (!O.a || !O.d)) { rsb_set_error("%s: operand needs coefficient vectors", what); return 1; }
    if ((O.kind == RSB_OPND_AFFINE2 || O.kind == RSB_OPND_POOLED) && (!O.V || !O.b || O.ku < 1)) { rsb_set_error("%s: affine operand needs V, b, ku", what); return 1; }
    if (O.kind == RSB_OPND_POOLED && (!O.arg || O.ns < 1)) { rsb_set_error("%s: pooled operand needs arg, ns", what); return 1; }
    return 0;
}

}  // namespace

RSB_EXPORT long rsb_linear_tc_weight_floats(int N, int K)
{
    const int NT = pick_nt(N);
    return (long)((N + NT - 1) / NT) * ((K + KC - 1) / KC) * 2 * NT * KC;
}

RSB_EXPORT int rsb_linear_tc_prep_weight(int N, int K, const float *W, int ldw, int transposed, float *Wp,
                                         cudaStream_t stream)
{
    RSB_REQUIRE(N >= 1 && K >= 1, "bad sizes");
    const int NT = pick_nt(N);
    const int n_tiles = (N + NT - 1) / NT, k_chunks = (K + KC - 1) / KC;
    const long total = (long)n_tiles * k_chunks * NT * KC;
    const int grid = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    weight_prep_kernel<<<grid, 256, 0, stream>>>(W, N, K, ldw, transposed, NT, n_tiles, k_chunks, Wp);
    RSB_CHECK_LAUNCH("weight_prep_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

// RSB_TC_TRACE=1: report every launch that falls back to the first-generation (SIMT-staged) kernels
static const int g_tc_sync = getenv("RSB_TC_SYNC") ? 1 : 0;
static const int g_tc_trace = [] { const char *v = getenv("RSB_TC_TRACE"); return (v && v[0] && v[0] != '0') ? 1 : 0; }();

RSB_EXPORT int rsb_gemm_rows(long rows, int N, const rsb_opnd_t *A, const float *Wp, const rsb_epi_t *E,
                             cudaStream_t stream)
{
    RSB_REQUIRE(rows >= 0 && N >= 1 && A && E && Wp, "bad arguments");
    if (check_opnd(*A, "rsb_gemm_rows")) return (int)cudaErrorInvalidValue;
    RSB_REQUIRE(E->kind == RSB_EPI_BIAS_STATS || (E->Yl && E->sc && E->sh && E->mu && E->inv), "dgrad epilogue needs Yl/sc/sh/mu/inv");
    if (rows == 0) return 0;
    {   // TMA-fed kernel (mlp_tc2.cu) whenever the operands meet its alignment rules
        const int r2 = rsb_gemm_rows2_launch(rows, N, A, Wp, E, stream);
        if (r2 >= 0) return r2;
        if (A->kind == RSB_OPND_GATHER || E->scatter) {
            rsb_set_error("rsb_gemm_rows: gathered operands / scattered results need the TMA-fed kernel (16-byte aligned table, pitch %% 4 == 0)");
            return (int)cudaErrorInvalidValue;
        }
        if (g_tc_trace)
            fprintf(stderr, "[rsb] gemm_rows -> first-generation kernel: rows %ld N %d K %d kind %d k0 %d ku %d ldu %d ldv %d ldy %d epi %d\n", rows, N,
                    A->K, A->kind, A->k0, A->ku, A->ldu, A->ldv, E->ldy, E->kind);
    }
    RowsParams P;
    P.A = *A; P.E = *E; P.Wp = Wp; P.rows = rows; P.N = N;
    P.NT = pick_nt(N);
    P.n_tiles = (N + P.NT - 1) / P.NT;
    P.k_chunks = (A->K + KC - 1) / KC;
    const size_t stage_b = 2 * (size_t)A_TILE_BYTES + 2 * (size_t)P.NT * KC * 4;
    const size_t fixed_b = sizeof(Barriers) + 4 * 32 * 33 * 4 + 1024;
    const size_t budget = 227 * 1024 - fixed_b;
    // asynchronous operand prefetch needs 16-byte aligned pieces
    const bool al_u = (A->ldu % 4 == 0) && ((uintptr_t)A->U % 16 == 0) && (A->k0 % 4 == 0) &&
                      (A->kind != RSB_OPND_DUAL_BN_RELU || A->ku % 4 == 0) && (A->kind != RSB_OPND_AFFINE2 || A->ku % 4 == 0);
    const bool al_v = !A->V || ((A->ldv % 4 == 0) && ((uintptr_t)A->V % 16 == 0));
    const size_t ring_b = 2 * 4 * (size_t)PROD_THREADS * 16;      // one ring slot: 2 pieces x 4 rows x 256 threads x 16 B
    P.raw_depth = 0;
    P.stages = (int)(budget / stage_b);
    if (al_u && al_v && !g_tc_sync) {
        for (int rd = 3; rd >= 2 && P.raw_depth == 0; rd--)
            if (budget >= 2 * stage_b + rd * ring_b) { P.raw_depth = rd; P.stages = (int)((budget - rd * ring_b) / stage_b); }
    }
    if (P.stages > STAGES_MAX) P.stages = STAGES_MAX;
    RSB_REQUIRE(P.stages >= 2, "tile does not fit");
    const size_t smem = P.stages * stage_b + fixed_b + P.raw_depth * ring_b;
    static bool attr_set = false;
    if (!attr_set) {
        RSB_CUDA(cudaFuncSetAttribute(gemm_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const long n_row_tiles = (rows + TM - 1) / TM;
    const int grid = (int)(n_row_tiles < rsb_sm_count() ? n_row_tiles : rsb_sm_count());
    gemm_rows_kernel<<<grid, THREADS, smem, stream>>>(P);
    RSB_CHECK_LAUNCH("gemm_rows_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_gemm_wgrad(long rows, const rsb_opnd_t *G, const rsb_opnd_t *X, float *dW, int ldw,
                              cudaStream_t stream)
{
    RSB_REQUIRE(rows >= 0 && G && X && dW, "bad arguments");
    if (check_opnd(*G, "rsb_gemm_wgrad(G)") || check_opnd(*X, "rsb_gemm_wgrad(X)")) return (int)cudaErrorInvalidValue;
    if (rows == 0) return 0;
    {
        const int r2 = rsb_gemm_wgrad2_launch(rows, G, X, dW, ldw, stream);
        if (r2 >= 0) return r2;
        if (G->kind == RSB_OPND_GATHER || X->kind == RSB_OPND_GATHER) {
            rsb_set_error("rsb_gemm_wgrad: gathered operands need the TMA-fed kernel (16-byte aligned table, pitch %% 4 == 0)");
            return (int)cudaErrorInvalidValue;
        }
        if (g_tc_trace)
            fprintf(stderr, "[rsb] gemm_wgrad -> first-generation kernel: rows %ld G(K %d kind %d k0 %d ku %d ldu %d ldv %d) X(K %d kind %d k0 %d ku %d ldu %d ldv %d)\n",
                    rows, G->K, G->kind, G->k0, G->ku, G->ldu, G->ldv, X->K, X->kind, X->k0, X->ku, X->ldu, X->ldv);
    }
    WgradParams P;
    P.G = *G; P.X = *X; P.dW = dW; P.ldw = ldw; P.rows = rows;
    P.M = G->K; P.N = X->K;
    P.NT = pick_nt(P.N);
    P.m_tiles = (P.M + TM - 1) / TM;
    P.n_tiles = (P.N + P.NT - 1) / P.NT;
    const size_t stage_b = 2 * (size_t)A_TILE_BYTES + 2 * (size_t)P.NT * KC * 4;
    const size_t tab_b = 6 * (size_t)(TM + NT_MAX) * 4;      // coefficient / index tables of the asynchronous path
    const size_t budget = 227 * 1024 - sizeof(Barriers) - 1024 - tab_b;
    // asynchronous prefetch: every piece must be 16-byte aligned; POOLED operands stay on the synchronous path
    auto aligned = [](const rsb_opnd_t *O) {
        if (O->kind == RSB_OPND_POOLED) return false;
        if ((O->ldu % 4) || ((uintptr_t)O->U % 16) || (O->k0 % 4)) return false;
        if ((O->kind == RSB_OPND_DUAL_BN_RELU || O->kind == RSB_OPND_AFFINE2) && (O->ku % 4)) return false;
        if (O->kind == RSB_OPND_AFFINE2 && ((O->ldv % 4) || ((uintptr_t)O->V % 16))) return false;
        return true;
    };
    // float4 per slab row of one operand tile — the same arithmetic as the kernel's setup()
    auto tile_quads = [](const rsb_opnd_t *O, int c_base, int width) {
        int nvalid = O->K - c_base;
        if (nvalid > width) nvalid = width;
        if (nvalid < 0) nvalid = 0;
        const int quads = (nvalid + 3) / 4;
        if (O->kind == RSB_OPND_DUAL_BN_RELU) return 2 * quads;
        if (O->kind == RSB_OPND_AFFINE2) {
            const int kk0 = (O->k0 + c_base) % O->ku;
            const int uw = (kk0 + nvalid > O->ku) ? O->ku : nvalid;
            return quads + (uw + 3) / 4;
        }
        return quads;
    };
    int max_quads = 0;
    for (int mt = 0; mt < P.m_tiles; mt++)
        for (int nt = 0; nt < P.n_tiles; nt++) {
            const int q = tile_quads(G, mt * TM, TM) + tile_quads(X, nt * P.NT, P.NT);
            if (q > max_quads) max_quads = q;
        }
    P.raw_slots = 0;
    P.raw_slot_bytes = 0;
    P.stages = (int)(budget / stage_b);
    if (aligned(G) && aligned(X) && !g_tc_sync) {
        const size_t slot = (size_t)KC * 16 * max_quads;
        // two operand stages are enough once loads are decoupled; spend the rest of shared memory on ring depth
        if (budget >= 2 * stage_b + 3 * slot) {
            int rs = (int)((budget - 2 * stage_b) / slot);
            if (rs > 8) rs = 8;
            P.raw_slots = rs;
            P.raw_slot_bytes = (int)slot;
            P.stages = (int)((budget - (size_t)rs * slot) / stage_b);
        }
    }
    if (P.stages > STAGES_MAX) P.stages = STAGES_MAX;
    RSB_REQUIRE(P.stages >= 2, "tile does not fit");
    const size_t smem = P.stages * stage_b + sizeof(Barriers) + 1024 + tab_b + (size_t)P.raw_slots * P.raw_slot_bytes;
    static bool attr_set = false;
    if (!attr_set) {
        RSB_CUDA(cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const long n_chunks = (rows + KC - 1) / KC;
    // every CTA flushes its whole dW tile with atomics: give each at least 2 chunks of reduction first
    // (swept 1/2/4/8/16 on the S3DIS step: 12.14 / 12.06 / 12.16 / 12.58 / 14.25 ms of wgrad time)
    static const int min_chunks = getenv("RSB_WGRAD_MINCHUNKS") ? atoi(getenv("RSB_WGRAD_MINCHUNKS")) : 2;
    long want = (n_chunks + min_chunks - 1) / min_chunks;
    if (want < 1) want = 1;
    const int grid = (int)(want < rsb_sm_count() ? want : rsb_sm_count());
    gemm_wgrad_kernel<<<grid, THREADS, smem, stream>>>(P);
    RSB_CHECK_LAUNCH("gemm_wgrad_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

// convenience entry kept from the first bring-up: Y = act(X) @ W^T + bias with modes 0/1/2 (see header)
RSB_EXPORT int rsb_linear_tc_forward(long rows, int K, int N, const float *X, int ldx, const float *Wp,
                                     const float *bias, int mode, const float *sc, const float *sh, float *Y,
                                     double *stats, cudaStream_t stream)
{
    RSB_REQUIRE(mode >= 0 && mode <= 2, "bad mode");
    rsb_opnd_t A = {};
    A.U = X; A.ldu = ldx; A.K = K; A.ku = K; A.a = sc; A.d = sh;
    A.kind = mode == 0 ? RSB_OPND_RAW : (mode == 1 ? RSB_OPND_BN_RELU : RSB_OPND_DUAL_BN_RELU);
    rsb_epi_t E = {};
    E.Y = Y; E.ldy = N; E.bias = bias; E.stats = stats; E.kind = RSB_EPI_BIAS_STATS;
    return rsb_gemm_rows(rows, N, &A, Wp, &E, stream);
}
