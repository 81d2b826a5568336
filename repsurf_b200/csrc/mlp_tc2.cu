// mlp_tc2.cu — TMA-fed versions of the shared-MLP GEMMs (same contract as mlp_tc.cu's rsb_gemm_rows / rsb_gemm_wgrad,
// which dispatch here whenever the operands satisfy TMA's 16-byte alignment rules).
//
// Why a second generation: the first-generation kernels staged every operand element through SIMT code three times
// (cp.async issue into a private ring, ring -> registers -> canonical tile, per-element transposition for wgrad); ncu's
// source view (profiles/r01_ncu_source_gemm.md) showed 500-700 SASS instructions per warp and chunk, issue slots 27-45 %
// busy, DRAM 14-30 %: a dependent-issue chain, not a bandwidth or tensor kernel.  Here
//   * one elected thread moves the RAW operand tiles with TMA (cp.async.bulk.tensor.2d) straight into the swizzled
//     layout the tensor core reads: K-major SWIZZLE_128B tiles for Y = A W^T, MN-major 32-byte-atom tiles for
//     dW = G^T X — the transposition of the weight-gradient operands is done by the MMA's major-ness bit, not by code;
//   * 16 transform warps rewrite those tiles IN PLACE: one LDS.128 (two for two-tensor operands), the BatchNorm /
//     ReLU / BatchNorm-backward transform, the hi/lo tf32 split, two STS.128 — ~40 instructions per 16-byte piece,
//     conflict-free because a warp touches 512 contiguous bytes.  Two-tensor operands (dual first layer, BatchNorm
//     backward) need no extra buffer: "hi" replaces the first raw tile, "lo" the second;
//   * weights small enough (<= 64 KB pre-split) stay resident in shared memory for the whole kernel;
//   * the epilogue leaves through a swizzled staging tile and TMA stores (clipped at the ragged edge by the tensor
//     map), column statistics are read back from the same tile and accumulated in fp64 in shared memory;
//   * the input-gradient GEMM applies the ReLU mask of the layer below and accumulates that layer's BatchNorm-backward
//     statistics in its epilogue (no separate streaming pass over dZ);
//   * the weight-gradient grid is (M tile) x (N tile) x (row split): every CTA owns one dW tile and flushes it once
//     with vector reductions, instead of every CTA cycling through all tiles.
// Numerics are unchanged: 3xTF32 (lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM).
#include "tc_common.cuh"
#include "mlp_tc.h"
#include <stdlib.h>

using namespace rsbtc;

namespace {

typedef rsb_opnd_t Opnd;
typedef rsb_epi_t Epi;

constexpr int TM = 128;                 // UMMA M
constexpr int KC = 32;                  // reduction depth of a pipeline stage (4 k-steps of 8 tf32)
constexpr int A_TILE = TM * KC * 4;     // 16 KB: one 128 x 32 fp32 tile
constexpr int BOX = 32 * 32 * 4;        // 4 KB: one 32-row x 32-channel box of a weight-gradient operand
constexpr int XF_WARPS = 16;            // transform warps
constexpr int XF_THREADS = XF_WARPS * 32;
constexpr int LOAD_WARP = XF_WARPS;     // TMA issuer
constexpr int MMA_WARP = XF_WARPS + 1;  // tcgen05.mma issuer (also owns the TMEM allocation)
constexpr int EPI_WARP0 = XF_WARPS + 2; // 4 epilogue warps
constexpr int THREADS2 = (XF_WARPS + 2 + 4) * 32;
// gemm_rows2 warp layouts.  The epilogue of a tile is a latency chain (tcgen05.ld -> [mask operand loads] -> staging -> TMA store
// -> statistics read-back) of 2.5-6 us, longer than a tile's share of the HBM stream when K is small (ncu, profiles/r02_ncu_gemm.md:
// issue slots 17-35 % busy, long-scoreboard stalls 9-25 per issue): there TWO epilogue warp sets alternate over the two TMEM
// accumulator buffers and 8 warps suffice for the in-place operand transform.  With many K chunks per tile the transform is the
// busier side: 16 transform warps, one epilogue set.  <XFW transform warps, ES epilogue sets of 4 warps>
template <int XFW, int ES>
struct RowsCfg {
    static constexpr int XF_W = XFW, XF_T = XFW * 32, LOAD_W = XFW, MMA_W = XFW + 1, EPI_W0 = XFW + 2, EPI_N = 4 * ES,
                         THREADS = (XFW + 2 + 4 * ES) * 32, PIECES = (TM * KC * 4 / 16) / (XFW * 32);
};
constexpr int STAGES_MAX = 4;           // gemm_rows2
constexpr int W_STAGES_MAX = 10;        // gemm_wgrad2 (small stages: depth hides the TMA latency)
constexpr int W_GROUPS = 2;             // gemm_wgrad2: transform warp groups, each owning every W_GROUPS-th chunk
constexpr int W_GT = XF_THREADS / W_GROUPS;    // threads of a group: 256 = the 16-byte pieces of one 32 x 32 box
constexpr int SMEM_MAX = 227 * 1024;

// ---- host: tensor maps ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encoder()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// [rows, cols] fp32 view with row pitch ld (floats); cols beyond the extent and rows beyond `rows` read as zero
int make_map(CUtensorMap *m, const float *base, long cols, long rows, long ld, int box_cols, int box_rows, CUtensorMapSwizzle sw)
{
    EncodeTiledFn fn = encoder();
    if (!fn) { rsb_set_error("cuTensorMapEncodeTiled is not available from this driver"); return 1; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rsb_set_error("cuTensorMapEncodeTiled failed (%d): base %p cols %ld rows %ld ld %ld box %dx%d", (int)r, (const void *)base, cols, rows, ld,
                      box_cols, box_rows);
        return 1;
    }
    return 0;
}

inline bool al16(const void *p) { return ((uintptr_t)p & 15) == 0; }
inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---- device: operand transform -------------------------------------------------------------------------------
struct Coef {
    float4 a, d, b, a2, d2;
};

__device__ __forceinline__ float4 ld_coef4(const float *p, int k, int nv)
{
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!p || nv <= 0) return r;
    if (nv == 4 && (((uintptr_t)(p + k)) & 15) == 0) return __ldg(reinterpret_cast<const float4 *>(p + k));
    r.x = __ldg(p + k);
    if (nv > 1) r.y = __ldg(p + k + 1);
    if (nv > 2) r.z = __ldg(p + k + 2);
    if (nv > 3) r.w = __ldg(p + k + 3);
    return r;
}

// coefficients of channels k .. k+3 (k includes k0); channels past the operand's width get zeros, which makes every
// transform return 0 there (the raw values are zero-filled by TMA)
__device__ __forceinline__ void load_coef(const Opnd &O, int k, int nv, Coef &c)
{
    c.a = ld_coef4(O.a, k, nv);
    c.d = ld_coef4(O.d, k, nv);
    c.b = ld_coef4(O.b, k, nv);
    if (O.kind == RSB_OPND_DUAL_BN_RELU) {
        c.a2 = ld_coef4(O.a, O.ku + k, nv);
        c.d2 = ld_coef4(O.d, O.ku + k, nv);
    } else {
        c.a2 = c.d2 = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__device__ __forceinline__ float4 xform(int kind, const float4 u, const float4 w, const Coef &c)
{
    float4 v;
    switch (kind) {
    case RSB_OPND_RAW:
    case RSB_OPND_GATHER:      // the centre subtraction of columns 0..2 happens in the transform loop (it needs the row)
        v = u;
        break;
    case RSB_OPND_BN_RELU:
        v.x = relu_nan(fmaf(u.x, c.a.x, c.d.x)); v.y = relu_nan(fmaf(u.y, c.a.y, c.d.y));
        v.z = relu_nan(fmaf(u.z, c.a.z, c.d.z)); v.w = relu_nan(fmaf(u.w, c.a.w, c.d.w));
        break;
    case RSB_OPND_DUAL_BN_RELU:
        v.x = relu_nan(fmaf(u.x, c.a.x, c.d.x) + fmaf(w.x, c.a2.x, c.d2.x));
        v.y = relu_nan(fmaf(u.y, c.a.y, c.d.y) + fmaf(w.y, c.a2.y, c.d2.y));
        v.z = relu_nan(fmaf(u.z, c.a.z, c.d.z) + fmaf(w.z, c.a2.z, c.d2.z));
        v.w = relu_nan(fmaf(u.w, c.a.w, c.d.w) + fmaf(w.w, c.a2.w, c.d2.w));
        break;
    default:   // RSB_OPND_AFFINE2
        v.x = fmaf(c.a.x, u.x, fmaf(c.b.x, w.x, c.d.x)); v.y = fmaf(c.a.y, u.y, fmaf(c.b.y, w.y, c.d.y));
        v.z = fmaf(c.a.z, u.z, fmaf(c.b.z, w.z, c.d.z)); v.w = fmaf(c.a.w, u.w, fmaf(c.b.w, w.w, c.d.w));
        break;
    }
    return v;
}

struct alignas(16) Bars2 {
    uint64_t raw_full[W_STAGES_MAX], full[W_STAGES_MAX], empty[W_STAGES_MAX], acc_full[2], acc_empty[2], w_bar;
    uint32_t tmem_slot, pad;
};

template <int COLS>
__device__ __forceinline__ uint32_t cta_setup(Bars2 *B, int tid, int warp, int full_count, int mma_warp = MMA_WARP)
{
    if (tid == 0) {
        for (int s = 0; s < W_STAGES_MAX; s++) { mbar_init(&B->raw_full[s], 1); mbar_init(&B->full[s], full_count); mbar_init(&B->empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&B->acc_full[a], 1); mbar_init(&B->acc_empty[a], 128); }
        mbar_init(&B->w_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == mma_warp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rsb_smem_addr(&B->tmem_slot)), "n"(COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return B->tmem_slot;
}

template <int COLS>
__device__ __forceinline__ void cta_teardown(uint32_t tmem_base, int warp, int mma_warp = MMA_WARP)
{
    tc_fence_before();
    __syncthreads();
    if (warp == mma_warp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(COLS));
    }
}

// ============================================================================================================
// Y = A @ W^T  (forward and input gradient)
// ============================================================================================================
struct Rows2Params {
    CUtensorMap mapA0, mapA1, mapY;
    Opnd A;
    Epi E;
    const float *Wp;
    long rows;
    int N, NT, n_tiles, k_chunks, stages;
    int nsub;         // an N tile is processed as nsub column slices of NT / nsub (when a full tile's weights do not fit a stage)
    int n_pieces;     // raw tensors behind the operand (1: RAW / BN_RELU / GATHER, 2: DUAL / AFFINE2)
    int gather;       // operand rows are gathered by index (TMA gather4) and centred
    int scatter;      // epilogue adds the rows into Y[scatter[r]] instead of storing Y[r]
    int w_resident;   // the pre-split weights of all (N tile, K chunk) pairs stay in shared memory
    int v_bufs;       // staging tiles per epilogue warp for the outgoing block (1 or 2)
    int n_tab;        // per-column epilogue tables in shared memory (bias | sc, sh, mu [, second half])
    int smem_stats;   // statistics accumulate in fp64 in shared memory (single N tile), else per-tile atomics
    int has_y;
};

template <int XFW, int ES>
__global__ void __launch_bounds__(RowsCfg<XFW, ES>::THREADS, 1) gemm_rows2_kernel(const __grid_constant__ Rows2Params P)
{
    typedef RowsCfg<XFW, ES> C;
    constexpr int R_XF_WARPS = C::XF_W, R_XF_THREADS = C::XF_T, R_LOAD_WARP = C::LOAD_W, R_MMA_WARP = C::MMA_W, R_EPI_WARP0 = C::EPI_W0,
                  R_EPI_WARPS = C::EPI_N, R_THREADS = C::THREADS, R_XF_PIECES = C::PIECES;
    extern __shared__ unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw0 = rsb_smem_addr(smem_raw);
    const uint32_t base = (raw0 + 1023u) & ~1023u;
    unsigned char *gbase = smem_raw + (base - raw0);

    const int NT = P.NT, S = P.stages, KCH = P.k_chunks;
    const int NS = P.nsub, NTs = NT / NS;                            // column slice of an N tile
    const uint32_t w_chunk_bytes = 2u * (uint32_t)NT * KC * 4;        // one (N tile, K chunk) of the pre-split weights: hi | lo
    const uint32_t w_half = (uint32_t)NT * KC * 4, w_slice = (uint32_t)NTs * KC * 4;
    const uint32_t stage_bytes = 2 * A_TILE + (P.w_resident ? 0u : 2u * w_slice);
    const uint32_t w_res = base + (uint32_t)S * stage_bytes;
    const uint32_t w_res_bytes = P.w_resident ? (uint32_t)(P.n_tiles * KCH) * w_chunk_bytes : 0u;
    const bool mask = P.E.kind == RSB_EPI_RELU_MASK;
    const int n_stat = mask ? (P.E.dual ? 3 : 2) : 2;
    const int epi_tiles = P.v_bufs + (mask ? (P.E.dual ? 2 : 1) : 0);
    const uint32_t epi0 = w_res + w_res_bytes;
    const int Npad = P.n_tiles * NT;
    float *tabs = reinterpret_cast<float *>(gbase + (epi0 - base) + (uint32_t)R_EPI_WARPS * epi_tiles * 4096u);
    double *sacc = reinterpret_cast<double *>(tabs + (size_t)P.n_tab * Npad);
    Bars2 *B = reinterpret_cast<Bars2 *>(reinterpret_cast<unsigned char *>(sacc) + (P.smem_stats ? (size_t)R_EPI_WARPS * n_stat * NT * 8 : 0));

    const Opnd &A = P.A;
    const Epi &E = P.E;

    // ---- one-time shared-memory state ----
    // epilogue tables, zero-padded to the tile grid
    for (int i = tid; i < P.n_tab * Npad; i += R_THREADS) {
        const int t = i / Npad, n = i - t * Npad;
        float v = 0.f;
        if (n < P.N) {
            if (!mask) v = E.bias ? __ldg(E.bias + n) : 0.f;
            else {
                const int half = t / 3, which = t - half * 3;          // [sc sh mu | sc2 sh2 mu2]
                const float *src = which == 0 ? E.sc : (which == 1 ? E.sh : E.mu);
                v = __ldg(src + half * P.N + n);
            }
        }
        tabs[i] = v;
    }
    if (P.smem_stats)
        for (int i = tid; i < R_EPI_WARPS * n_stat * NT; i += R_THREADS) sacc[i] = 0.0;
    // single-tensor operands: the "lo" tile is written by the transform only where channels exist; clear it once
    if (P.n_pieces == 1)
        for (int s = 0; s < S; s++)
            for (int i = tid; i < A_TILE / 16; i += R_THREADS) sts128(base + (uint32_t)s * stage_bytes + A_TILE + (uint32_t)i * 16, make_float4(0.f, 0.f, 0.f, 0.f));
    fence_proxy_async();
    if (warp == R_LOAD_WARP && lane == 0) {
        prefetch_tmap(&P.mapA0);
        if (P.n_pieces == 2) prefetch_tmap(&P.mapA1);
        if (P.has_y) prefetch_tmap(&P.mapY);
    }
    const uint32_t tmem_base = cta_setup<512>(B, tid, warp, R_XF_THREADS + (P.w_resident ? 0 : 1), R_MMA_WARP);

    const long n_row_tiles = (P.rows + TM - 1) / TM;
    const int per_tile = P.n_tiles * NS;                               // work items of one row tile: (N tile, slice)
    const long n_work = n_row_tiles * per_tile;

    if (warp < R_XF_WARPS) {
        // =============================== transform warps ===============================
        // thread -> the 16-byte pieces tid + R_XF_THREADS j of the 128 x 32 tile (rows tid/8 + R_XF_THREADS/8 j): all share (row & 7)
        // and the physical chunk, hence the logical channel quad and its coefficients
        const int q = (tid & 7) ^ ((tid >> 3) & 7);
        Coef cf;
        const bool hoist = KCH == 1;
        int nv = max(0, min(4, A.K - q * 4));
        if (hoist) load_coef(A, A.k0 + q * 4, nv, cf);
        const bool two = P.n_pieces == 2;
        const bool zero_lo = !two && KCH > 1;       // stale "lo" data of a full chunk under an invalid quad of the last chunk
        uint32_t it = 0;
        for (long w = blockIdx.x; w < n_work; w += gridDim.x) {
            const long tile_row0 = (w / per_tile) * TM;
            for (int kc = 0; kc < KCH; kc++, it++) {
                const int s = it % S;
                if (!hoist) {
                    nv = max(0, min(4, A.K - (kc * KC + q * 4)));
                    load_coef(A, A.k0 + kc * KC + q * 4, nv, cf);
                }
                mbar_wait(&B->raw_full[s], (it / S) & 1);
                const uint32_t a0 = base + (uint32_t)s * stage_bytes + (uint32_t)tid * 16;
                if (nv > 0) {
                    float4 u[R_XF_PIECES], x[R_XF_PIECES];
#pragma unroll
                    for (int j = 0; j < R_XF_PIECES; j++) {
                        u[j] = lds128(a0 + j * (R_XF_THREADS * 16));
                        x[j] = two ? lds128(a0 + A_TILE + j * (R_XF_THREADS * 16)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (P.gather && kc == 0 && q == 0) {
                        // columns 0..2 of a gathered row are the neighbour's coordinates: make them relative to the group centre
                        // (same subtraction, same rounding as the row builder group_rows_fwd)
#pragma unroll
                        for (int j = 0; j < R_XF_PIECES; j++) {
                            const long r = tile_row0 + (tid >> 3) + j * (R_XF_THREADS / 8);
                            if (r < P.rows) {
                                const float *cen = A.V + (r / A.ns) * 3;
                                u[j].x = __fsub_rn(u[j].x, __ldg(cen));
                                u[j].y = __fsub_rn(u[j].y, __ldg(cen + 1));
                                u[j].z = __fsub_rn(u[j].z, __ldg(cen + 2));
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < R_XF_PIECES; j++) {
                        float4 hi, lo;
                        split4(xform(A.kind, u[j], x[j], cf), hi, lo);
                        sts128(a0 + j * (R_XF_THREADS * 16), hi);
                        sts128(a0 + A_TILE + j * (R_XF_THREADS * 16), lo);
                    }
                } else if (zero_lo) {
#pragma unroll
                    for (int j = 0; j < R_XF_PIECES; j++) sts128(a0 + A_TILE + j * (R_XF_THREADS * 16), make_float4(0.f, 0.f, 0.f, 0.f));
                }
                fence_proxy_async();
                mbar_arrive(&B->full[s]);
            }
        }
    } else if (warp == R_LOAD_WARP) {
        // =============================== TMA issuer ===============================
        if (P.gather) {
            // every lane gathers 4 of the tile's 128 rows per K chunk: 32 gather4 instructions fill the same SWIZZLE_128B tile
            // a plain box load would (each lands on 4 consecutive 128-byte lines)
            if (lane == 0 && P.w_resident) {
                mbar_arrive_expect_tx(&B->w_bar, w_res_bytes);
                bulk_g2s(w_res, P.Wp, w_res_bytes, &B->w_bar);
            }
            uint32_t it = 0;
            int idn[4];       // neighbour indices of the NEXT work item's rows (loaded one item ahead: no load latency in the loop)
#pragma unroll
            for (int e = 0; e < 4; e++) idn[e] = __ldg(A.arg + min(((long)blockIdx.x / per_tile) * TM + 4 * lane + e, P.rows - 1));
            for (long w = blockIdx.x; w < n_work; w += gridDim.x) {
                const long tile = w / per_tile;
                const int rem = (int)(w - tile * per_tile);
                const int nt = rem / NS, sub = rem - nt * NS;
                int id[4];
#pragma unroll
                for (int e = 0; e < 4; e++) id[e] = idn[e];
                if (w + gridDim.x < n_work) {
                    const long r1 = ((w + gridDim.x) / per_tile) * TM + 4 * lane;
#pragma unroll
                    for (int e = 0; e < 4; e++) idn[e] = __ldg(A.arg + min(r1 + e, P.rows - 1));
                }
                for (int kc = 0; kc < KCH; kc++, it++) {
                    const int s = it % S;
                    mbar_wait(&B->empty[s], ((it / S) & 1) ^ 1);
                    const uint32_t st = base + (uint32_t)s * stage_bytes;
                    if (lane == 0) mbar_arrive_expect_tx(&B->raw_full[s], A_TILE);
                    __syncwarp();
                    tma_gather4(st + (uint32_t)lane * 512u, &P.mapA0, kc * KC, id[0], id[1], id[2], id[3], &B->raw_full[s]);
                    if (lane == 0 && !P.w_resident) {
                        const float *wc = P.Wp + ((size_t)nt * KCH + kc) * (2 * (size_t)NT * KC) + (size_t)sub * NTs * KC;
                        mbar_arrive_expect_tx(&B->full[s], 2u * w_slice);
                        bulk_g2s(st + 2 * A_TILE, wc, w_slice, &B->full[s]);
                        bulk_g2s(st + 2 * A_TILE + w_slice, wc + (size_t)NT * KC, w_slice, &B->full[s]);
                    }
                }
            }
        } else if (lane == 0) {
            if (P.w_resident) {
                mbar_arrive_expect_tx(&B->w_bar, w_res_bytes);
                bulk_g2s(w_res, P.Wp, w_res_bytes, &B->w_bar);
            }
            uint32_t it = 0;
            for (long w = blockIdx.x; w < n_work; w += gridDim.x) {
                const long tile = w / per_tile;
                const int rem = (int)(w - tile * per_tile);
                const int nt = rem / NS, sub = rem - nt * NS;
                const int row0 = (int)(tile * TM);
                for (int kc = 0; kc < KCH; kc++, it++) {
                    const int s = it % S;
                    mbar_wait(&B->empty[s], ((it / S) & 1) ^ 1);
                    const uint32_t st = base + (uint32_t)s * stage_bytes;
                    mbar_arrive_expect_tx(&B->raw_full[s], (uint32_t)P.n_pieces * A_TILE);
                    tma_load_2d(st, &P.mapA0, kc * KC, row0, &B->raw_full[s]);
                    if (P.n_pieces == 2) tma_load_2d(st + A_TILE, &P.mapA1, kc * KC, row0, &B->raw_full[s]);
                    if (!P.w_resident) {
                        // slice `sub` of the chunk: rows sub*NTs .. of the hi block, then of the lo block (8-row groups of the
                        // canonical layout are 1024 B apart, so a slice is contiguous inside each block)
                        const float *wc = P.Wp + ((size_t)nt * KCH + kc) * (2 * (size_t)NT * KC) + (size_t)sub * NTs * KC;
                        mbar_arrive_expect_tx(&B->full[s], 2u * w_slice);
                        bulk_g2s(st + 2 * A_TILE, wc, w_slice, &B->full[s]);
                        bulk_g2s(st + 2 * A_TILE + w_slice, wc + (size_t)NT * KC, w_slice, &B->full[s]);
                    }
                }
            }
        }
    } else if (warp == R_MMA_WARP) {
        // =============================== MMA issuer ===============================
        const uint32_t idesc = umma_idesc_tf32(TM, NTs, false);
        if (P.w_resident) mbar_wait(&B->w_bar, 0);
        uint32_t it = 0, acc_it = 0;
        for (long w = blockIdx.x; w < n_work; w += gridDim.x, acc_it++) {
            const int rem = (int)(w % per_tile);
            const int nt = rem / NS, sub = rem - nt * NS;
            const int ab = acc_it & 1;
            mbar_wait(&B->acc_empty[ab], ((acc_it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(ab * 256);
            for (int kc = 0; kc < KCH; kc++, it++) {
                const int s = it % S;
                mbar_wait(&B->full[s], (it / S) & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t st = base + (uint32_t)s * stage_bytes;
                    const uint32_t a_hi = st, a_lo = st + A_TILE;
                    const uint32_t b_hi = P.w_resident ? w_res + (uint32_t)(nt * KCH + kc) * w_chunk_bytes + (uint32_t)sub * w_slice : st + 2 * A_TILE;
                    const uint32_t b_lo = b_hi + (P.w_resident ? w_half : w_slice);
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ks++) {
                        // A: K-major SWIZZLE_128B (8 tf32 = 32 B inside the 128-byte row); W: canonical no-swizzle K-major
                        // tiles written by weight_prep_kernel (core matrices 128 B apart along K, 1024 B along N)
                        const uint64_t dah = umma_desc(a_hi + ks * 32, 16, 1024, 2), dal = umma_desc(a_lo + ks * 32, 16, 1024, 2);
                        const uint64_t dbh = umma_desc(b_hi + ks * 256, 128, 1024, 0), dbl = umma_desc(b_lo + ks * 256, 128, 1024, 0);
                        umma_tf32(tmem_d, dal, dbh, idesc, (kc | ks) ? 1u : 0u);   // small terms first
                        umma_tf32(tmem_d, dah, dbl, idesc, 1u);
                        umma_tf32(tmem_d, dah, dbh, idesc, 1u);
                    }
                    umma_commit(&B->empty[s]);
                    if (kc == KCH - 1) umma_commit(&B->acc_full[ab]);
                }
                __syncwarp();
            }
        }
    } else {
        // =============================== epilogue ===============================
        const int q = warp & 3;                 // TMEM lane quadrant this warp may access
        const int ew = warp - R_EPI_WARP0;      // staging / statistics slot
        const int eset = ew >> 2;               // set 0 drains accumulator buffer 0 (even work items), set 1 buffer 1
        const uint32_t my_epi = epi0 + (uint32_t)ew * epi_tiles * 4096u;
        const uint32_t p_tile = my_epi + (uint32_t)P.v_bufs * 4096u;   // mask epilogue: v*(yl-mu) tile(s)
        double *my_acc = sacc + (size_t)ew * n_stat * NT;
        const float *t_bias = tabs, *t_sc = tabs, *t_sh = tabs + Npad, *t_mu = tabs + 2 * Npad;
        const uint32_t my_row_off = (uint32_t)lane * 128u;             // this lane's row inside a staging tile
        const int sw = lane & 7;
        uint32_t acc_it = 0, vb = 0;
        for (long w = blockIdx.x; w < n_work; w += gridDim.x, acc_it++) {
            const int ab = acc_it & 1;
            if (ES == 2 && ab != eset) continue;
            const long tile = w / per_tile;
            const int rem = (int)(w - tile * per_tile);
            const int nt = rem / NS, sub = rem - nt * NS;
            const int col0 = nt * NT + sub * NTs;              // first output column of this work item
            const long row = tile * TM + q * 32 + lane;
            const bool row_ok = row < P.rows;
            if (mask) {
                // the mask operand of this warp set's NEXT work item starts its way from DRAM to L2 now: the epilogue reads it
                // row by row with plain loads, a latency chain that would otherwise start only after the accumulator is ready
                const long wn = w + (long)gridDim.x * ES;
                if (wn < n_work) {
                    const long tn = wn / per_tile;
                    const int remn = (int)(wn - tn * per_tile);
                    const int ntn = remn / NS, subn = remn - ntn * NS;
                    const long rn = tn * TM + q * 32 + lane;
                    if (rn < P.rows) {
                        const float *yn = E.Yl + (size_t)rn * E.ldl + ntn * NT + subn * NTs;
                        const int ncn = min(NTs, P.N - (ntn * NT + subn * NTs));
                        for (int c0 = 0; c0 < ncn; c0 += 32) {
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(yn + c0));
                            if (E.dual) asm volatile("prefetch.global.L2 [%0];" ::"l"(yn + P.N + c0));
                        }
                    }
                }
            }
            mbar_wait(&B->acc_full[ab], (acc_it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 256);
            const int ncols = min(NTs, P.N - col0);
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                float v[32];
                tmem_ld32(taddr + c0, v);
                const int nl = col0 + c0;                          // first column of the block (table index == column)
                // the staging tile of this block must have been read by its previous TMA store
                if (P.has_y) {
                    if (lane == 0) { if (P.v_bufs == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
                    __syncwarp();
                }
                const uint32_t vt = my_epi + (vb % (uint32_t)P.v_bufs) * 4096u + my_row_off;
                if (P.scatter) {
                    // grouping backward fused into the input-gradient GEMM: row r adds its 32-column block to row scatter[r] of
                    // the per-point gradient table (columns >= N carry zeros: the weight tile is zero-padded)
                    if (row_ok) {
                        float *dst = E.Y + (size_t)__ldg(E.scatter + row) * E.ldy + nl;
#pragma unroll
                        for (int c = 0; c < 8; c++)
                            if (nl + 4 * c < P.N)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * c), "f"(v[4 * c]), "f"(v[4 * c + 1]),
                                             "f"(v[4 * c + 2]), "f"(v[4 * c + 3]) : "memory");
                    }
                } else if (!mask) {
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const float4 b4 = *reinterpret_cast<const float4 *>(t_bias + nl + 4 * c);
                        float4 o;
                        o.x = row_ok ? v[4 * c] + b4.x : 0.f; o.y = row_ok ? v[4 * c + 1] + b4.y : 0.f;
                        o.z = row_ok ? v[4 * c + 2] + b4.z : 0.f; o.w = row_ok ? v[4 * c + 3] + b4.w : 0.f;
                        sts128(vt + (uint32_t)((c ^ sw) << 4), o);
                    }
                } else {
                    // dgrad: ReLU mask of the layer below recomputed from its stored pre-BatchNorm output, and the
                    // products the BatchNorm-backward statistics need
                    const float *yl = E.Yl + (size_t)(row_ok ? row : 0) * E.ldl + nl;
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const bool ok = row_ok && nl + 4 * c < P.N;
                        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        const float4 y = ok ? __ldg(reinterpret_cast<const float4 *>(yl + 4 * c)) : z4;
                        const float4 sc = *reinterpret_cast<const float4 *>(t_sc + nl + 4 * c);
                        const float4 sh = *reinterpret_cast<const float4 *>(t_sh + nl + 4 * c);
                        const float4 mu = *reinterpret_cast<const float4 *>(t_mu + nl + 4 * c);
                        float4 z;
                        z.x = fmaf(y.x, sc.x, sh.x); z.y = fmaf(y.y, sc.y, sh.y); z.z = fmaf(y.z, sc.z, sh.z); z.w = fmaf(y.w, sc.w, sh.w);
                        float4 y2 = z4, mu2 = z4;
                        if (E.dual) {
                            y2 = ok ? __ldg(reinterpret_cast<const float4 *>(yl + P.N + 4 * c)) : z4;
                            const float4 sc2 = *reinterpret_cast<const float4 *>(t_sc + 3 * Npad + nl + 4 * c);
                            const float4 sh2 = *reinterpret_cast<const float4 *>(t_sh + 3 * Npad + nl + 4 * c);
                            mu2 = *reinterpret_cast<const float4 *>(t_mu + 3 * Npad + nl + 4 * c);
                            z.x += fmaf(y2.x, sc2.x, sh2.x); z.y += fmaf(y2.y, sc2.y, sh2.y);
                            z.z += fmaf(y2.z, sc2.z, sh2.z); z.w += fmaf(y2.w, sc2.w, sh2.w);
                        }
                        float4 o;
                        o.x = (ok && z.x > 0.f) ? v[4 * c] : 0.f; o.y = (ok && z.y > 0.f) ? v[4 * c + 1] : 0.f;
                        o.z = (ok && z.z > 0.f) ? v[4 * c + 2] : 0.f; o.w = (ok && z.w > 0.f) ? v[4 * c + 3] : 0.f;
                        sts128(vt + (uint32_t)((c ^ sw) << 4), o);
                        sts128(p_tile + my_row_off + (uint32_t)((c ^ sw) << 4),
                               make_float4(o.x * (y.x - mu.x), o.y * (y.y - mu.y), o.z * (y.z - mu.z), o.w * (y.w - mu.w)));
                        if (E.dual)
                            sts128(p_tile + 4096u + my_row_off + (uint32_t)((c ^ sw) << 4),
                                   make_float4(o.x * (y2.x - mu2.x), o.y * (y2.y - mu2.y), o.z * (y2.z - mu2.z), o.w * (y2.w - mu2.w)));
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (P.has_y && lane == 0) {
                    tma_store_2d(&P.mapY, nl, (int)(tile * TM + q * 32), vt - my_row_off);
                    bulk_commit();
                }
                if (E.stats) {
                    // column `lane` of the block: sums over the warp's 32 rows, read back from the staging tile(s)
                    const uint32_t cbase = (vt - my_row_off) + (uint32_t)((lane & 3) << 2);
                    const int cq = lane >> 2;
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
                    if (!mask) {
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            float t;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(cbase + (uint32_t)i * 128u + (uint32_t)((cq ^ (i & 7)) << 4)));
                            s0 += t;
                            s1 = fmaf(t, t, s1);
                        }
                    } else {
                        const uint32_t pbase = p_tile + (uint32_t)((lane & 3) << 2);
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            const uint32_t off = (uint32_t)i * 128u + (uint32_t)((cq ^ (i & 7)) << 4);
                            float t, p1;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(cbase + off));
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(p1) : "r"(pbase + off));
                            s0 += t;
                            s1 += p1;
                            if (E.dual) {
                                float p2;
                                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(p2) : "r"(pbase + 4096u + off));
                                s2 += p2;
                            }
                        }
                    }
                    const int n = nl + lane;
                    if (P.smem_stats) {
                        if (c0 + lane < NTs) {              // a slice is a multiple of 16 columns, not of 32
                            double *a = my_acc + sub * NTs + c0 + lane;
                            a[0] += (double)s0;
                            a[NT] += (double)s1;
                            if (n_stat == 3) a[2 * NT] += (double)s2;
                        }
                    } else if (n < P.N) {
                        atomicAdd(E.stats + n, (double)s0);
                        if (!mask) atomicAdd(E.stats + P.N + n, (double)s1);
                        else {
                            atomicAdd(E.stats + P.N + n, (double)s1 * (double)__ldg(E.inv + n));
                            if (E.dual) atomicAdd(E.stats + 2 * P.N + n, (double)s2 * (double)__ldg(E.inv + P.N + n));
                        }
                    }
                    __syncwarp();      // the product tile is reused by the next block
                }
                vb++;
            }
            tc_fence_before();
            mbar_arrive(&B->acc_empty[ab]);
        }
        if (P.smem_stats && E.stats) {
            __syncwarp();
            for (int c = lane; c < NT; c += 32) {
                if (c >= P.N) break;
                const double s0 = my_acc[c], s1 = my_acc[NT + c];
                if (s0 != 0.0 || s1 != 0.0) {
                    atomicAdd(E.stats + c, s0);
                    atomicAdd(E.stats + P.N + c, mask ? s1 * (double)__ldg(E.inv + c) : s1);
                }
                if (n_stat == 3) {
                    const double s2 = my_acc[2 * NT + c];
                    if (s2 != 0.0) atomicAdd(E.stats + 2 * P.N + c, s2 * (double)__ldg(E.inv + P.N + c));
                }
            }
        }
        if (lane == 0) bulk_wait_all();
    }
    cta_teardown<512>(tmem_base, warp, R_MMA_WARP);
}

// ============================================================================================================
// dW[m, n] += sum_r G(r, m) * X(r, n)   (weight gradient; reduction over rows)
// ============================================================================================================
struct Wgrad2Params {
    CUtensorMap mapG0, mapG1, mapX0, mapX1;
    Opnd G, X;
    float *dW;
    int ldw;
    long rows;
    int M, N, NT, NTB, m_tiles, n_tiles, splits, stages;
    int g_pieces, x_pieces;
    int GB;           // 32-channel boxes per G part of a stage (ceil(min(M, 128) / 32))
    int x_gather;     // X rows are gathered by index (TMA gather4) and centred (RSB_OPND_GATHER)
};

// first column (in the piece's tensor map) of the 32-channel box that starts at logical channel c of operand O
__device__ __forceinline__ int piece0_col(const Opnd &O, int c) { return O.kind == RSB_OPND_AFFINE2 ? (O.k0 + c) % O.ku : O.k0 + c; }

__global__ void __launch_bounds__(THREADS2, 1) gemm_wgrad2_kernel(const __grid_constant__ Wgrad2Params P)
{
    extern __shared__ unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw0 = rsb_smem_addr(smem_raw);
    const uint32_t base = (raw0 + 1023u) & ~1023u;
    unsigned char *gbase = smem_raw + (base - raw0);

    const int NT = P.NT, NTB = P.NTB, S = P.stages;
    const uint32_t x_bytes = (uint32_t)NTB * BOX;                   // one of raw/hi, raw/lo of the X tile
    // stage = [G hi/raw0 | G lo/raw1 | X hi/raw0 | X lo/raw1], each G part only as many 32-channel boxes as the matrix has (GB <= 4):
    // the sa1 launches have 32 or 64 gradient channels, and with a full 128-channel slot per part three quarters of a stage were
    // padding - the freed shared memory becomes pipeline depth (the kernel is latency-bound on its TMA loads: ncu long-scoreboard
    // stall 4-7 per issue at 20-36 % DRAM, profiles/r02_ncu_full_seg.md).  The MMA still reads 4 boxes per part; the ones beyond GB
    // alias the following parts and only feed accumulator rows nobody reads.
    const uint32_t g_bytes = (uint32_t)P.GB * BOX;
    const uint32_t stage_bytes = 2 * g_bytes + 2 * x_bytes;
    constexpr int TABW = TM + 256;
    float *ctab = reinterpret_cast<float *>(gbase + (size_t)S * stage_bytes);   // [a | b | d | a2 | d2] x (G 128 | X 256)
    Bars2 *B = reinterpret_cast<Bars2 *>(ctab + 5 * TABW);

    const int sp = blockIdx.x % P.splits;
    const int tile_id = blockIdx.x / P.splits;
    const int nt = tile_id % P.n_tiles, mt = tile_id / P.n_tiles;
    const int vg = min(TM, P.M - mt * TM), vx = min(NT, P.N - nt * NT);       // valid channels of the two tiles
    const int gb = (vg + 31) / 32, xb = (vx + 31) / 32;                       // boxes that are loaded and transformed
    const long n_chunks = (P.rows + KC - 1) / KC;
    const long cps = (n_chunks + P.splits - 1) / P.splits;
    const long c_begin = (long)sp * cps, c_end = min(n_chunks, c_begin + cps);
    const long my_chunks = max(0L, c_end - c_begin);

    // coefficient tables of the two channel tiles (zeros past the valid channels)
    for (int i = tid; i < 5 * TABW; i += THREADS2) {
        const int t = i / TABW, c = i - t * TABW;
        const bool isg = c < TM;
        const Opnd &O = isg ? P.G : P.X;
        const int cl = isg ? c : c - TM;
        const int ch = (isg ? mt * TM : nt * NT) + cl;
        float v = 0.f;
        if (cl < (isg ? vg : vx)) {
            const int k = O.k0 + ch;
            if (t == 0) v = O.a ? __ldg(O.a + k) : 0.f;
            else if (t == 1) v = O.b ? __ldg(O.b + k) : 0.f;
            else if (t == 2) v = O.d ? __ldg(O.d + k) : 0.f;
            else if (O.kind == RSB_OPND_DUAL_BN_RELU) v = __ldg((t == 3 ? O.a : O.d) + O.ku + k);
        }
        ctab[i] = v;
    }
    // boxes that are never loaded (channels past the matrices) still feed the MMA: their products land in accumulator
    // rows / columns nobody reads, but keep them finite
    for (int i = tid; i < S * (int)(stage_bytes / 16); i += THREADS2) sts128(base + (uint32_t)i * 16, make_float4(0.f, 0.f, 0.f, 0.f));
    fence_proxy_async();
    if (warp == LOAD_WARP && lane == 0) {
        prefetch_tmap(&P.mapG0); prefetch_tmap(&P.mapX0);
        if (P.g_pieces == 2) prefetch_tmap(&P.mapG1);
        if (P.x_pieces == 2) prefetch_tmap(&P.mapX1);
    }
    const uint32_t tmem_base = cta_setup<256>(B, tid, warp, W_GT);

    if (my_chunks > 0) {
        if (warp < XF_WARPS) {
            // =============================== transform warps ===============================
            // 16-byte piece p of a region: box p >> 8, k-row (p >> 3) & 31, physical chunk p & 7; with 32-byte swizzle atoms
            // the logical channel quad is chunk ^ ((row & 3) << 1).
            // The 16 warps work as W_GROUPS independent groups, group g owning the chunks ci = g (mod W_GROUPS): a thread's
            // chunk is a ~1500-cycle dependent chain (mbarrier wait -> LDS -> transform -> STS -> proxy fence -> arrive), and with
            // every thread walking EVERY chunk only one chunk per SM was in transformation at a time (0.8 us per 32-row chunk
            // whatever the pipeline depth); now W_GROUPS chunks are.
            const int grp = tid / W_GT, gtid = tid - grp * W_GT;
            const int kr = (gtid >> 3) & 31;
            const int q = (gtid & 7) ^ ((kr & 3) << 1);
            const bool g2 = P.g_pieces == 2, x2 = P.x_pieces == 2;
            for (long ci = grp; ci < my_chunks; ci += W_GROUPS) {
                const uint32_t it = (uint32_t)ci;
                const int s = it % S;
                const int lim = (int)min((long)KC, P.rows - (c_begin + ci) * KC);   // valid rows of this chunk
                const bool row_ok = kr < lim;
                mbar_wait(&B->raw_full[s], (it / S) & 1);
                const uint32_t st = base + (uint32_t)s * stage_bytes;
                // ---- G tile: boxes 0 .. gb - 1, one 16-byte piece per thread and box (W_GT == 256 pieces == one box) ----
                for (int bx = 0; bx < gb; bx++) {
                    const uint32_t addr = st + (uint32_t)(bx * 256 + gtid) * 16;
                    const int ch = bx * 32 + q * 4;
                    Coef c;
                    c.a = *reinterpret_cast<const float4 *>(ctab + ch);
                    c.b = *reinterpret_cast<const float4 *>(ctab + TABW + ch);
                    c.d = *reinterpret_cast<const float4 *>(ctab + 2 * TABW + ch);
                    c.a2 = *reinterpret_cast<const float4 *>(ctab + 3 * TABW + ch);
                    c.d2 = *reinterpret_cast<const float4 *>(ctab + 4 * TABW + ch);
                    const float4 u = lds128(addr);
                    const float4 x = g2 ? lds128(addr + g_bytes) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 v = xform(P.G.kind, u, x, c);
                    if (!row_ok) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 hi, lo;
                    split4(v, hi, lo);
                    sts128(addr, hi);
                    sts128(addr + g_bytes, lo);
                }
                // ---- X tile ----
                for (int bx = 0; bx < xb; bx++) {
                    const uint32_t addr = st + 2 * g_bytes + (uint32_t)(bx * 256 + gtid) * 16;
                    const int ch = TM + bx * 32 + q * 4;
                    Coef c;
                    c.a = *reinterpret_cast<const float4 *>(ctab + ch);
                    c.b = *reinterpret_cast<const float4 *>(ctab + TABW + ch);
                    c.d = *reinterpret_cast<const float4 *>(ctab + 2 * TABW + ch);
                    c.a2 = *reinterpret_cast<const float4 *>(ctab + 3 * TABW + ch);
                    c.d2 = *reinterpret_cast<const float4 *>(ctab + 4 * TABW + ch);
                    float4 u = lds128(addr);
                    const float4 x = x2 ? lds128(addr + x_bytes) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (P.x_gather && nt == 0 && bx == 0 && q == 0 && row_ok) {
                        // gathered rows: neighbour coordinates -> relative to the group centre (columns 0..2)
                        const float *cen = P.X.V + (((c_begin + ci) * KC + kr) / P.X.ns) * 3;
                        u.x = __fsub_rn(u.x, __ldg(cen)); u.y = __fsub_rn(u.y, __ldg(cen + 1)); u.z = __fsub_rn(u.z, __ldg(cen + 2));
                    }
                    float4 v = xform(P.X.kind, u, x, c);
                    if (!row_ok) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 hi, lo;
                    split4(v, hi, lo);
                    sts128(addr, hi);
                    sts128(addr + x_bytes, lo);
                }
                fence_proxy_async();
                mbar_arrive(&B->full[s]);
            }
        } else if (warp == LOAD_WARP) {
            // =============================== TMA issuer ===============================
            if (P.x_gather) {
                // G boxes by lane 0; the X boxes (32 rows x 32 channels each) by gather4: lane l fetches rows 4 (l % 8) .. + 3 of the
                // chunk for box l / 8, then lanes + 32 ... until all xb boxes are issued
                const uint32_t tx = (uint32_t)(P.g_pieces * gb + xb) * BOX;
                uint32_t it = 0;
                const int g4 = lane & 7;
                int idn[4];   // neighbour indices of the next chunk's rows, loaded one chunk ahead
#pragma unroll
                for (int e = 0; e < 4; e++) idn[e] = __ldg(P.X.arg + min(c_begin * KC + 4 * g4 + e, P.rows - 1));
                for (long ci = 0; ci < my_chunks; ci++, it++) {
                    const int s = it % S;
                    int id[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) id[e] = idn[e];
                    if (ci + 1 < my_chunks) {
#pragma unroll
                        for (int e = 0; e < 4; e++) idn[e] = __ldg(P.X.arg + min((c_begin + ci + 1) * KC + 4 * g4 + e, P.rows - 1));
                    }
                    mbar_wait(&B->empty[s], ((it / S) & 1) ^ 1);
                    const uint32_t st = base + (uint32_t)s * stage_bytes;
                    const int row0 = (int)((c_begin + ci) * KC);
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&B->raw_full[s], tx);
                        for (int g = 0; g < gb; g++) {
                            const int c = mt * TM + g * 32;
                            tma_load_2d(st + (uint32_t)g * BOX, &P.mapG0, piece0_col(P.G, c), row0, &B->raw_full[s]);
                            if (P.g_pieces == 2) tma_load_2d(st + g_bytes + (uint32_t)g * BOX, &P.mapG1, P.G.k0 + c, row0, &B->raw_full[s]);
                        }
                    }
                    __syncwarp();
                    for (int x = lane >> 3; x < xb; x += 4)
                        tma_gather4(st + 2 * g_bytes + (uint32_t)x * BOX + (uint32_t)g4 * 512u, &P.mapX0, nt * NT + x * 32, id[0], id[1], id[2], id[3],
                                    &B->raw_full[s]);
                }
            } else if (lane == 0) {
                const uint32_t tx = (uint32_t)(P.g_pieces * gb + P.x_pieces * xb) * BOX;
                uint32_t it = 0;
                for (long ci = 0; ci < my_chunks; ci++, it++) {
                    const int s = it % S;
                    mbar_wait(&B->empty[s], ((it / S) & 1) ^ 1);
                    const uint32_t st = base + (uint32_t)s * stage_bytes;
                    const int row0 = (int)((c_begin + ci) * KC);
                    mbar_arrive_expect_tx(&B->raw_full[s], tx);
                    for (int g = 0; g < gb; g++) {
                        const int c = mt * TM + g * 32;
                        tma_load_2d(st + (uint32_t)g * BOX, &P.mapG0, piece0_col(P.G, c), row0, &B->raw_full[s]);
                        if (P.g_pieces == 2) tma_load_2d(st + g_bytes + (uint32_t)g * BOX, &P.mapG1, P.G.k0 + c, row0, &B->raw_full[s]);
                    }
                    for (int x = 0; x < xb; x++) {
                        const int c = nt * NT + x * 32;
                        tma_load_2d(st + 2 * g_bytes + (uint32_t)x * BOX, &P.mapX0, piece0_col(P.X, c), row0, &B->raw_full[s]);
                        if (P.x_pieces == 2) tma_load_2d(st + 2 * g_bytes + x_bytes + (uint32_t)x * BOX, &P.mapX1, P.X.k0 + c, row0, &B->raw_full[s]);
                    }
                }
            }
        } else if (warp == MMA_WARP) {
            // =============================== MMA issuer ===============================
            const uint32_t idesc = umma_idesc_tf32(TM, NT, true);
            uint32_t it = 0;
            for (long ci = 0; ci < my_chunks; ci++, it++) {
                const int s = it % S;
                mbar_wait(&B->full[s], (it / S) & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t st = base + (uint32_t)s * stage_bytes;
                    const uint32_t a_hi = st, a_lo = st + g_bytes, b_hi = st + 2 * g_bytes, b_lo = b_hi + x_bytes;
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ks++) {
                        // MN-major, 32-byte swizzle atoms: 32-channel groups BOX bytes apart (LBO), 4-row groups 512 B apart
                        // (SBO), 8 reduction rows = 1024 B per k-step
                        const uint32_t adv = ks * 1024;
                        const uint64_t dah = umma_desc(a_hi + adv, BOX, 512, 1), dal = umma_desc(a_lo + adv, BOX, 512, 1);
                        const uint64_t dbh = umma_desc(b_hi + adv, BOX, 512, 1), dbl = umma_desc(b_lo + adv, BOX, 512, 1);
                        umma_tf32(tmem_base, dal, dbh, idesc, (ci | ks) ? 1u : 0u);
                        umma_tf32(tmem_base, dah, dbl, idesc, 1u);
                        umma_tf32(tmem_base, dah, dbh, idesc, 1u);
                    }
                    umma_commit(&B->empty[s]);
                    if (ci == my_chunks - 1) umma_commit(&B->acc_full[0]);
                }
                __syncwarp();
            }
        } else {
            // =============================== epilogue: flush the dW tile ===============================
            const int q = warp & 3;
            mbar_wait(&B->acc_full[0], 0);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
            const int ml = q * 32 + lane;                          // channel of G inside the tile
            const bool vec = (P.ldw & 3) == 0 && (((uintptr_t)P.dW) & 15) == 0 && ((nt * NT) & 3) == 0;
            for (int c0 = 0; c0 < vx; c0 += 32) {
                float v[32];
                tmem_ld32(taddr + c0, v);
                if (ml < vg) {
                    float *dst = P.dW + (size_t)(mt * TM + ml) * P.ldw + nt * NT + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (vec && c0 + j + 4 <= vx) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; e++)
                                if (c0 + j + e < vx) atomicAdd(dst + j + e, v[j + e]);
                        }
                    }
                }
            }
            tc_fence_before();
        }
    }
    cta_teardown<256>(tmem_base, warp);
}

// ============================================================================================================
// dW for NARROW layers (M <= 32 gradient channels, N <= 32 input channels) on the fp32 pipe
// ============================================================================================================
// The narrowest sa1 weight gradients are 32 x 32 matrices reduced over 2.6 M rows: 1 kFLOP per 250-500 bytes of operands.
// On the tensor core the gradient tile is padded to UMMA's M = 128: three quarters of every operand read from shared memory
// feed accumulator rows nobody uses, and with the in-place transform's own traffic the kernel sits at 55-75 % of the SM's
// shared-memory bandwidth (ncu l1tex throughput; profiles/r02_ncu_full_seg.md) at 0.27-0.46 of the HBM roofline, whatever the
// pipeline depth or the number of transform groups.  Here a warp streams 4 rows per 512-byte load (lane = row x channel quad),
// applies the operand transform in registers, and every lane owns a (4 MQ) x (8 NO) block of dW in registers: per row
// 4 MQ + 8 NO shuffles and 32 MQ NO FMAs.  Plain fp32 (no 3xTF32 split needed), HBM-bound.
struct NarrowParams {
    Opnd G, X;
    float *dW;
    int ldw;
    long rows;
    int M, N;
};

// raw 16-byte piece(s) of operand O at row r (already clamped), channel c (multiple of 4, < O.K)
__device__ __forceinline__ void narrow_load(const Opnd &O, long r, int c, float4 &u, float4 &w)
{
    w = make_float4(0.f, 0.f, 0.f, 0.f);
    switch (O.kind) {
    case RSB_OPND_DUAL_BN_RELU:
        u = __ldg(reinterpret_cast<const float4 *>(O.U + (size_t)r * O.ldu + O.k0 + c));
        w = __ldg(reinterpret_cast<const float4 *>(O.U + (size_t)r * O.ldu + O.ku + O.k0 + c));
        break;
    case RSB_OPND_AFFINE2:
        u = __ldg(reinterpret_cast<const float4 *>(O.U + (size_t)r * O.ldu + (O.k0 + c) % O.ku));
        w = __ldg(reinterpret_cast<const float4 *>(O.V + (size_t)r * O.ldv + O.k0 + c));
        break;
    case RSB_OPND_GATHER: {
        u = __ldg(reinterpret_cast<const float4 *>(O.U + (size_t)__ldg(O.arg + r) * O.ldu + c));
        if (c == 0) {      // neighbour coordinates -> relative to the group centre
            const float *cen = O.V + (r / O.ns) * 3;
            u.x = __fsub_rn(u.x, __ldg(cen)); u.y = __fsub_rn(u.y, __ldg(cen + 1)); u.z = __fsub_rn(u.z, __ldg(cen + 2));
        }
        break;
    }
    default:
        u = __ldg(reinterpret_cast<const float4 *>(O.U + (size_t)r * O.ldu + O.k0 + c));
        break;
    }
}

template <int MQ, int NO>
__global__ void __launch_bounds__(256, (MQ * NO == 1) ? 3 : (MQ * NO == 2 ? 2 : 1)) wgrad_narrow_kernel(const NarrowParams P)
{
    __shared__ __align__(16) float4 s_coef[(MQ + NO) * 8][5];       // [operand quad][a, d, b, a2, d2]
    __shared__ float s_dw[MQ * 32 * NO * 32];
    const int tid = threadIdx.x, lane = tid & 31;
    const int li = lane >> 3, lj = lane & 7;                          // load mapping: row li of the 4-row group, channel quad lj
    for (int e = tid; e < (MQ + NO) * 8; e += 256) {
        const bool isg = e < MQ * 8;
        const Opnd &O = isg ? P.G : P.X;
        const int c = 4 * (isg ? e : e - MQ * 8);
        Coef cf;
        load_coef(O, O.k0 + c, c < O.K ? 4 : 0, cf);
        s_coef[e][0] = cf.a; s_coef[e][1] = cf.d; s_coef[e][2] = cf.b; s_coef[e][3] = cf.a2; s_coef[e][4] = cf.d2;
    }
    for (int e = tid; e < MQ * 32 * NO * 32; e += 256) s_dw[e] = 0.f;
    __syncthreads();

    float acc[MQ][4][NO][8];
#pragma unroll
    for (int a = 0; a < MQ; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
            for (int c = 0; c < NO; c++)
#pragma unroll
                for (int d = 0; d < 8; d++) acc[a][b][c][d] = 0.f;

    const long n_groups = (P.rows + 3) / 4;
    const long wstride = (long)gridDim.x * 8;
    float4 ru[MQ + NO], rw[MQ + NO];                                   // raw pieces of the NEXT group (software prefetch)
    auto fetch = [&](long grp) {
        const long r = min(grp * 4 + li, P.rows - 1);
#pragma unroll
        for (int t = 0; t < MQ + NO; t++) {
            const bool isg = t < MQ;
            const Opnd &O = isg ? P.G : P.X;
            const int c = 4 * (lj + 8 * (isg ? t : t - MQ));
            if (c < O.K) narrow_load(O, r, c, ru[t], rw[t]);
            else ru[t] = rw[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    long grp = (long)blockIdx.x * 8 + (tid >> 5);
    if (grp < n_groups) fetch(grp);
    for (; grp < n_groups; grp += wstride) {
        // ---- transform this group's pieces (this lane: row li, quads lj + 8 t) ----
        float4 tv[MQ + NO];
        const bool row_ok = grp * 4 + li < P.rows;
#pragma unroll
        for (int t = 0; t < MQ + NO; t++) {
            const bool isg = t < MQ;
            const Opnd &O = isg ? P.G : P.X;
            const int e = (isg ? t : MQ + (t - MQ)) * 8 + lj;
            const int c = 4 * (lj + 8 * (isg ? t : t - MQ));
            Coef cf;
            cf.a = s_coef[e][0]; cf.d = s_coef[e][1]; cf.b = s_coef[e][2]; cf.a2 = s_coef[e][3]; cf.d2 = s_coef[e][4];
            float4 v = xform(O.kind, ru[t], rw[t], cf);
            if (!row_ok || c >= O.K) v = make_float4(0.f, 0.f, 0.f, 0.f);
            tv[t] = v;
        }
        if (grp + wstride < n_groups) fetch(grp + wstride);           // next group's loads fly during the outer products
        // ---- 4 rows: dW block of this lane += g (4 MQ channels) x x (8 NO channels) ----
#pragma unroll
        for (int rr = 0; rr < 4; rr++) {
            float g[MQ][4], x[NO][8];
#pragma unroll
            for (int a = 0; a < MQ; a++) {
                const int src = rr * 8 + lj;                           // lane holding quad lj (+ 8 a) of row rr
                g[a][0] = __shfl_sync(0xffffffffu, tv[a].x, src); g[a][1] = __shfl_sync(0xffffffffu, tv[a].y, src);
                g[a][2] = __shfl_sync(0xffffffffu, tv[a].z, src); g[a][3] = __shfl_sync(0xffffffffu, tv[a].w, src);
            }
#pragma unroll
            for (int c = 0; c < NO; c++) {
                const int s0 = rr * 8 + 2 * li;                        // lanes holding quads 2 li, 2 li + 1 (+ 8 c) of row rr
                x[c][0] = __shfl_sync(0xffffffffu, tv[MQ + c].x, s0); x[c][1] = __shfl_sync(0xffffffffu, tv[MQ + c].y, s0);
                x[c][2] = __shfl_sync(0xffffffffu, tv[MQ + c].z, s0); x[c][3] = __shfl_sync(0xffffffffu, tv[MQ + c].w, s0);
                x[c][4] = __shfl_sync(0xffffffffu, tv[MQ + c].x, s0 + 1); x[c][5] = __shfl_sync(0xffffffffu, tv[MQ + c].y, s0 + 1);
                x[c][6] = __shfl_sync(0xffffffffu, tv[MQ + c].z, s0 + 1); x[c][7] = __shfl_sync(0xffffffffu, tv[MQ + c].w, s0 + 1);
            }
#pragma unroll
            for (int a = 0; a < MQ; a++)
#pragma unroll
                for (int b = 0; b < 4; b++)
#pragma unroll
                    for (int c = 0; c < NO; c++)
#pragma unroll
                        for (int d = 0; d < 8; d++) acc[a][b][c][d] = fmaf(g[a][b], x[c][d], acc[a][b][c][d]);
        }
    }
    // ---- block reduction in shared memory, one global reduction per element and block ----
    constexpr int NW = NO * 32;
#pragma unroll
    for (int a = 0; a < MQ; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
            for (int c = 0; c < NO; c++)
#pragma unroll
                for (int d = 0; d < 8; d++) {
                    const int m = 4 * (lj + 8 * a) + b, n = 8 * (li + 4 * c) + d;
                    atomicAdd(&s_dw[m * NW + n], acc[a][b][c][d]);
                }
    __syncthreads();
    for (int e = tid; e < MQ * 32 * NW; e += 256) {
        const int m = e / NW, n = e - m * NW;
        if (m < P.M && n < P.N) atomicAdd(P.dW + (size_t)m * P.ldw + n, s_dw[e]);
    }
}

// Returns 0 when the launch was made, -1 when the problem is not one for this kernel.
int wgrad_narrow_launch(long rows, const rsb_opnd_t *G, const rsb_opnd_t *X, float *dW, int ldw, cudaStream_t stream)
{
    const int M = G->K, N = X->K;
    // measured (scripts/ab_gemm.py, 2.6 M rows): 32 x 32 0.47 -> 0.61 of the HBM roofline; from 64 x 20 up the M N / 32 FMAs per row
    // and lane make the fp32 pipe the limit (64 x 32: 0.38 -> 0.30, 64 x 64: 0.60 -> 0.39), so only the 32 x 32 class comes here
    if (M > 32 || N > 32 || (M & 3) || (N & 3) || rows < 8192) return -1;
    if (G->kind == RSB_OPND_GATHER) return -1;
    auto wrap_ok = [](const Opnd &O) { return O.kind != RSB_OPND_AFFINE2 || ((O.k0 % O.ku) + O.K <= O.ku) || ((O.ku & 3) == 0); };
    if (!wrap_ok(*G) || !wrap_ok(*X)) return -1;
    NarrowParams P;
    P.G = *G; P.X = *X; P.dW = dW; P.ldw = ldw; P.rows = rows; P.M = M; P.N = N;
    const int mq = M > 32 ? 2 : 1, no = N > 32 ? 2 : 1;
    const int blocks = rsb_sm_count() * ((mq * no == 1) ? 3 : (mq * no == 2 ? 2 : 1));
    wgrad_narrow_kernel<1, 1><<<blocks, 256, 0, stream>>>(P);
    RSB_CHECK_LAUNCH("wgrad_narrow_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

// RSB_WGRAD_TC=1: always the tensor-core weight-gradient kernel (A/B against the narrow-layer kernel)
int g_wgrad_tc_only = [] { const char *v = getenv("RSB_WGRAD_TC"); return (v && v[0] && v[0] != '0') ? 1 : 0; }();
int g_force_v1 = [] { const char *v = getenv("RSB_TC_V1"); return (v && v[0] && v[0] != '0') ? 1 : 0; }();
// SM budget of the persistent row GEMMs (0 = all SMs), see rsb_tc_set_sm_budget
int g_sm_budget = 0;

bool opnd_tma_ok(const Opnd &O)
{
    if (O.kind == RSB_OPND_POOLED) return false;
    if (O.kind == RSB_OPND_GATHER) return al16(O.U) && (O.ldu & 3) == 0 && O.k0 == 0 && O.arg && O.V && O.ns >= 1 && O.ku >= 1;
    if (!al16(O.U) || (O.ldu & 3) || (O.k0 & 3)) return false;
    if (O.kind == RSB_OPND_DUAL_BN_RELU && (O.ku & 3)) return false;
    if (O.kind == RSB_OPND_AFFINE2 && (!al16(O.V) || (O.ldv & 3) || (O.ku & 3))) return false;
    return true;
}

}  // namespace

RSB_EXPORT void rsb_tc_set_generation(int gen) { g_force_v1 = gen == 1 ? 1 : 0; }

// The row GEMMs partition their tiles statically over one persistent CTA per SM.  When the caller KNOWS that a concurrent launch
// on another stream holds some SMs for longer than the GEMM lasts (the cluster FPS of the next level, one CTA per SM for
// milliseconds), the CTAs that find no SM run as a second wave after the first has finished: the launch takes twice as long.
// With the grid limited to the SMs that are free it takes 148 / free instead.  Applies to the launches that follow, until reset
// with 0; host-side state of the calling thread's launch sequence (not a device setting).
RSB_EXPORT void rsb_tc_set_sm_budget(int sms) { g_sm_budget = sms > 0 ? sms : 0; }

// Returns 0 when the launch was made, -1 when the problem is not eligible (the caller then uses the first-generation
// kernel), > 0 on errors.
int rsb_gemm_rows2_launch(long rows, int N, const rsb_opnd_t *A, const float *Wp, const rsb_epi_t *E, cudaStream_t stream)
{
    if (g_force_v1 || !encoder()) return -1;
    if (!opnd_tma_ok(*A)) return -1;
    if (rows + TM >= (1L << 31)) return -1;
    if (A->kind == RSB_OPND_AFFINE2 && (A->k0 % A->ku) + A->K > A->ku) return -1;     // U would wrap inside the operand
    const bool mask = E->kind == RSB_EPI_RELU_MASK;
    if (E->Y && (!al16(E->Y) || (E->ldy & 3))) return -1;
    if (E->scatter && (mask || E->bias || E->stats || !E->Y)) return -1;
    if (mask && (!al16(E->Yl) || (E->ldl & 3) || (N & 3))) return -1;
    // same tiling rule as pick_nt() in mlp_tc.cu, which laid out the pre-split weight buffer
    const int nt0 = (N + 255) / 256;
    const int NT = round_up((N + nt0 - 1) / nt0, nt0 > 1 ? 32 : 16);
    const int n_tiles = (N + NT - 1) / NT;

    Rows2Params P;
    memset(&P, 0, sizeof(P));
    P.A = *A; P.E = *E; P.Wp = Wp; P.rows = rows; P.N = N; P.NT = NT; P.n_tiles = n_tiles;
    P.k_chunks = (A->K + KC - 1) / KC;
    P.n_pieces = (A->kind == RSB_OPND_DUAL_BN_RELU || A->kind == RSB_OPND_AFFINE2) ? 2 : 1;
    P.gather = A->kind == RSB_OPND_GATHER ? 1 : 0;
    P.scatter = E->scatter ? 1 : 0;
    P.has_y = E->Y != nullptr && !P.scatter;
    P.n_tab = mask ? (E->dual ? 6 : 3) : 1;
    P.smem_stats = (E->stats && n_tiles == 1) ? 1 : 0;
    const int n_stat = mask ? (E->dual ? 3 : 2) : 2;
    const size_t w_chunk = 2 * (size_t)NT * KC * 4;
    const size_t w_total = (size_t)n_tiles * P.k_chunks * w_chunk;
    const int Npad = n_tiles * NT;
    // warp layout: two epilogue sets when a tile has little transform work (K chunks x raw tensors) per epilogue pass
    const int cfg2 = (P.k_chunks * P.n_pieces <= 4) ? 1 : 0;
    const int R_EPI_WARPS = cfg2 ? 8 : 4;
    bool fit = false;
    const int want_ss = P.smem_stats;
    // preference order: whole N tiles before column slices, shared-memory statistics before per-tile atomics, two staging
    // tiles per epilogue warp before one, resident weights before streamed ones
    for (int ns = 1; ns <= 2 && !fit; ns++) {
        // a full tile's weight chunk (up to 64 KB) may not leave room for two pipeline stages next to the epilogue's
        // staging tiles: then the tile is processed as two column slices (the operand tile is staged once per slice)
        if (ns == 2 && (NT % 32)) break;
        const size_t w_stage = w_chunk / ns;
        for (int ss = want_ss; ss >= 0 && !fit; ss--)
            for (int vb = 2; vb >= 1 && !fit; vb--) {
                const int epi_tiles = vb + (mask ? (E->dual ? 2 : 1) : 0);
                const size_t fixed = R_EPI_WARPS * (size_t)epi_tiles * 4096 + (size_t)P.n_tab * Npad * 4 +
                                     (ss ? (size_t)R_EPI_WARPS * n_stat * NT * 8 : 0) + sizeof(Bars2) + 1024 + 64;
                if (fixed + 2 * (2 * (size_t)A_TILE) > (size_t)SMEM_MAX) continue;
                const size_t budget = SMEM_MAX - fixed;
                for (int res = 1; res >= 0 && !fit; res--) {
                    if (res && (w_total > 64 * 1024 || ns > 1)) continue;
                    const size_t stage_b = 2 * (size_t)A_TILE + (res ? 0 : w_stage);
                    if (budget < (res ? w_total : 0) + 2 * stage_b) continue;
                    int st = (int)((budget - (res ? w_total : 0)) / stage_b);
                    if (st > STAGES_MAX) st = STAGES_MAX;
                    if (res && st < 3) continue;           // residency must not starve the pipeline
                    P.stages = st; P.w_resident = res; P.v_bufs = vb; P.nsub = ns; P.smem_stats = ss;
                    fit = true;
                }
            }
    }
    if (!fit) return -1;
    const int epi_tiles = P.v_bufs + (mask ? (E->dual ? 2 : 1) : 0);
    const size_t stage_b = 2 * (size_t)A_TILE + (P.w_resident ? 0 : w_chunk / P.nsub);
    const size_t smem = (size_t)P.stages * stage_b + (P.w_resident ? w_total : 0) + R_EPI_WARPS * (size_t)epi_tiles * 4096 +
                        (size_t)P.n_tab * Npad * 4 + (P.smem_stats ? (size_t)R_EPI_WARPS * n_stat * NT * 8 : 0) + sizeof(Bars2) + 1024 + 64;

    // operand pieces: columns [0, K) of the map are the operand's channels k0 .. k0 + K
    const float *p0 = A->kind == RSB_OPND_AFFINE2 ? A->U + (A->k0 % A->ku) : A->U + A->k0;
    if (P.gather) {     // gather4 wants a one-row box; the table has ku rows
        if (make_map(&P.mapA0, A->U, A->K, A->ku, A->ldu, KC, 1, CU_TENSOR_MAP_SWIZZLE_128B)) return (int)cudaErrorInvalidValue;
    } else if (make_map(&P.mapA0, p0, A->K, rows, A->ldu, KC, TM, CU_TENSOR_MAP_SWIZZLE_128B)) return (int)cudaErrorInvalidValue;
    if (P.n_pieces == 2) {
        const float *p1 = A->kind == RSB_OPND_DUAL_BN_RELU ? A->U + A->ku + A->k0 : A->V + A->k0;
        const long ld1 = A->kind == RSB_OPND_DUAL_BN_RELU ? A->ldu : A->ldv;
        if (make_map(&P.mapA1, p1, A->K, rows, ld1, KC, TM, CU_TENSOR_MAP_SWIZZLE_128B)) return (int)cudaErrorInvalidValue;
    }
    if (P.has_y && make_map(&P.mapY, E->Y, N, rows, E->ldy, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return (int)cudaErrorInvalidValue;

    static bool attr_set = false;
    if (!attr_set) {
        RSB_CUDA(cudaFuncSetAttribute(gemm_rows2_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
        RSB_CUDA(cudaFuncSetAttribute(gemm_rows2_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
        attr_set = true;
    }
    const long n_work = ((rows + TM - 1) / TM) * n_tiles * P.nsub;
    int sms = rsb_sm_count();
    if (g_sm_budget > 0 && g_sm_budget < sms) sms = g_sm_budget;
    const int grid = (int)(n_work < sms ? n_work : sms);
    if (cfg2) gemm_rows2_kernel<8, 2><<<grid, RowsCfg<8, 2>::THREADS, smem, stream>>>(P);
    else gemm_rows2_kernel<16, 1><<<grid, RowsCfg<16, 1>::THREADS, smem, stream>>>(P);
    RSB_CHECK_LAUNCH("gemm_rows2_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

int rsb_gemm_wgrad2_launch(long rows, const rsb_opnd_t *G, const rsb_opnd_t *X, float *dW, int ldw, cudaStream_t stream)
{
    if (g_force_v1 || !encoder()) return -1;
    if (!opnd_tma_ok(*G) || !opnd_tma_ok(*X)) return -1;
    if (G->kind == RSB_OPND_GATHER) return -1;           // only the X operand (the layer's input rows) is ever gathered
    if (rows + KC >= (1L << 31)) return -1;
    if (!g_wgrad_tc_only) {                              // narrow layers over many rows: the fp32-pipe kernel
        const int rn = wgrad_narrow_launch(rows, G, X, dW, ldw, stream);
        if (rn >= 0) return rn;
    }
    // AFFINE2 operands whose U tensor wraps (k % ku) inside the operand: every 32-channel box must stay inside one period
    auto wrap_ok = [](const Opnd &O) {
        if (O.kind != RSB_OPND_AFFINE2) return true;
        if ((O.k0 % O.ku) + O.K <= O.ku) return true;
        return (O.ku % 32) == 0 && (O.k0 % 32) == 0;
    };
    if (!wrap_ok(*G) || !wrap_ok(*X)) return -1;

    Wgrad2Params P;
    memset(&P, 0, sizeof(P));
    P.G = *G; P.X = *X; P.dW = dW; P.ldw = ldw; P.rows = rows;
    P.M = G->K; P.N = X->K;
    P.m_tiles = (P.M + TM - 1) / TM;
    P.n_tiles = (P.N + 255) / 256;
    P.NT = round_up((P.N + P.n_tiles - 1) / P.n_tiles, 16);
    // N tiles start at multiples of NT: with more than one tile they must also be multiples of the 32-channel boxes
    if (P.n_tiles > 1 && (P.NT % 32)) P.NT = round_up(P.NT, 32);
    if (P.NT > 256) return -1;
    P.n_tiles = (P.N + P.NT - 1) / P.NT;
    P.NTB = (P.NT + 31) / 32;
    P.g_pieces = (G->kind == RSB_OPND_DUAL_BN_RELU || G->kind == RSB_OPND_AFFINE2) ? 2 : 1;
    P.x_pieces = (X->kind == RSB_OPND_DUAL_BN_RELU || X->kind == RSB_OPND_AFFINE2) ? 2 : 1;
    P.GB = ((P.M < TM ? P.M : TM) + 31) / 32;
    const size_t stage_b = 2 * (size_t)P.GB * BOX + 2 * (size_t)P.NTB * BOX;
    // + 16 KB of slack behind the last stage: the MMA's 4-box read of the last G part may run past it (see the kernel)
    const size_t fixed = 5 * (size_t)(TM + 256) * 4 + sizeof(Bars2) + 1024 + 64 + 4 * (size_t)BOX;
    int st = (int)((SMEM_MAX - fixed) / stage_b);
    if (st > W_STAGES_MAX) st = W_STAGES_MAX;
    if (st < 2) return -1;
    P.stages = st;
    const size_t smem = (size_t)st * stage_b + fixed;

    const long n_chunks = (rows + KC - 1) / KC;
    const int tiles = P.m_tiles * P.n_tiles;
    long splits = rsb_sm_count() / tiles;
    if (splits < 1) splits = 1;
    if (splits > (n_chunks + 3) / 4) splits = (n_chunks + 3) / 4;      // at least ~4 chunks of reduction per flush
    if (splits < 1) splits = 1;
    P.splits = (int)splits;

    // piece maps: the column coordinate is the stored tensor's own column index (see piece0_col)
    P.x_gather = X->kind == RSB_OPND_GATHER ? 1 : 0;
    auto maps = [&](const Opnd &O, CUtensorMap *m0, CUtensorMap *m1) {
        if (O.kind == RSB_OPND_GATHER)      // one-row box for gather4 over the per-point table (ku rows)
            return make_map(m0, O.U, O.K, O.ku, O.ldu, 32, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        const long ext0 = O.kind == RSB_OPND_AFFINE2 ? O.ku : (long)O.k0 + O.K;
        if (make_map(m0, O.U, ext0, rows, O.ldu, 32, KC, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
        if (O.kind == RSB_OPND_DUAL_BN_RELU)
            return make_map(m1, O.U + O.ku, (long)O.k0 + O.K, rows, O.ldu, 32, KC, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (O.kind == RSB_OPND_AFFINE2)
            return make_map(m1, O.V, (long)O.k0 + O.K, rows, O.ldv, 32, KC, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        return 0;
    };
    if (maps(*G, &P.mapG0, &P.mapG1) || maps(*X, &P.mapX0, &P.mapX1)) return (int)cudaErrorInvalidValue;

    static bool attr_set = false;
    if (!attr_set) {
        RSB_CUDA(cudaFuncSetAttribute(gemm_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
        attr_set = true;
    }
    gemm_wgrad2_kernel<<<tiles * P.splits, THREADS2, smem, stream>>>(P);
    RSB_CHECK_LAUNCH("gemm_wgrad2_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}
