// mlp_tc.cu — the shared-MLP GEMM of RepSurf-U on the 5th-gen tensor cores (tcgen05 + TMEM), fp32-faithful.
//
//   Y[r, n] = sum_k act(X)[r, k] * W[n, k] + bias[n]            rows x K  @  (N x K)^T
//
// replaces the reference's library calls nn.Conv2d/Conv1d(1x1)/nn.Linear on the RepSurf path
//   classification/modules/repsurface_utils.py:236-243, segmentation/modules/repsurface_utils.py:220-227, 267-282
// and FUSES what the reference runs as separate cuDNN / elementwise passes around them:
//   * prologue  act(X): the previous layer's train-mode BatchNorm + ReLU is applied while the A tile is
//     staged (per-channel scale/shift), including the channel-de-differentiated first layer
//     relu(bn_l(y_l) + bn_f(y_f)) ("dual" mode) — the normalised activations never exist in HBM;
//   * epilogue: + bias, store Y once, and the per-channel sum / sum-of-squares needed by THIS layer's
//     BatchNorm statistics, accumulated in fp64.
//
// Numerics: tcgen05 has no fp32 x fp32 MMA.  Every operand is split a = hi + lo with hi = tf32(a),
// lo = tf32(a - hi) and three kind::tf32 MMAs (hi*hi + hi*lo + lo*hi) accumulate in fp32 in TMEM
// ("3xTF32"): ~2^-21 relative per product, i.e. fp32-level, which the 1e-5 parity bar needs (plain TF32
// gives 1e-3).  The dropped lo*lo term is < 2^-22.
//
// Structure (one CTA per SM, persistent over 128-row tiles; 9 warps):
//   warps 0-3  producers: one thread per row; load K-chunk (32) of X, apply act(), split hi/lo, write both
//              tiles to shared memory in the UMMA canonical K-major no-swizzle layout (8x16B core matrices);
//              one elected thread also fetches the matching pre-split weight chunk with cp.async.bulk (TMA 1D).
//   warp  4    MMA issuer: one elected lane issues 4 k-steps x 3 tcgen05.mma per chunk into a double-buffered
//              TMEM accumulator (128 lanes x N columns), tcgen05.commit -> mbarriers.
//   warps 5-8  epilogue: tcgen05.ld 32x32b -> registers -> +bias -> global store + fp64 column statistics.
#include "common.cuh"

namespace {

constexpr int TM = 128;          // rows per tile (UMMA M)
constexpr int KC = 32;           // K elements per pipeline chunk (4 UMMA k-steps of 8)
constexpr int STAGES = 2;
constexpr int NT_MAX = 256;      // columns per N tile (UMMA N <= 256)
constexpr int A_TILE_BYTES = TM * KC * 4;           // 16 KB (one of hi / lo)
constexpr int THREADS = 9 * 32;

struct TcParams {
    const float *X;      // [rows, ldx]
    const float *Wp;     // pre-split weights, canonical layout: [n_tiles][k_chunks][2 (hi,lo)][NT x KC]
    const float *bias;   // [N] or nullptr
    const float *sc;     // prologue scale  [K] (mode 1) or [2K] (mode 2)
    const float *sh;     // prologue shift
    float *Y;            // [rows, N]
    double *stats;       // [2N]: sum, sum of squares (accumulated with atomics) or nullptr
    long rows;
    int K, ldx, N, NT, n_tiles, k_chunks;
    int mode;            // 0: A = X;  1: A = relu(X*sc+sh);  2: A = relu(X[:, :K]*sc+sh + X[:, K:2K]*sc'+sh')
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
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     rsb_smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(rsb_smem_addr(bar))
                 : "memory");
}

// UMMA shared-memory descriptor, K-major, no swizzle (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   [0,14) start>>4, [16,30) leading-dim byte offset>>4 (between core matrices along K),
//   [32,46) stride-dim byte offset>>4 (between 8-row groups), [46,48) version = 1, [61,64) layout = 0.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major (InstrDescriptor bit layout)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N)
{
    return (1u << 4) /* D = f32 */ | (2u << 7) /* A = tf32 */ | (2u << 10) /* B = tf32 */ |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
__device__ __forceinline__ float to_tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// shared-memory carve-up
struct Smem {
    // stage s: A_hi, A_lo (16 KB each), B_hi, B_lo (NT*128 B each)
    static __device__ __forceinline__ size_t stage_bytes(int NT) { return 2 * A_TILE_BYTES + 2 * (size_t)NT * KC * 4; }
};

__global__ void __launch_bounds__(THREADS, 1) linear_tc_kernel(TcParams P)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NT = P.NT;
    const size_t stage_bytes = Smem::stage_bytes(NT);
    const size_t b_bytes = (size_t)NT * KC * 4;  // one of hi / lo
    unsigned char *stage_base = smem;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * stage_bytes);
    uint64_t *full_bar = bars;                 // [STAGES]  producers (128 arrivals + expect_tx thread) -> MMA
    uint64_t *empty_bar = bars + STAGES;       // [STAGES]  MMA commit -> producers
    uint64_t *acc_full = bars + 2 * STAGES;    // [2]       MMA commit -> epilogue
    uint64_t *acc_empty = bars + 2 * STAGES + 2;  // [2]    epilogue (128 arrivals) -> MMA
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);
    float *stat_tile = reinterpret_cast<float *>(bars + 2 * STAGES + 6);  // [4 warps][32][33]

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 128 + 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(rsb_smem_addr(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long n_row_tiles = (P.rows + TM - 1) / TM;
    const int kc_count = P.k_chunks;

    if (warp < 4) {
        // =============================== producers ===============================
        const int r = tid;  // row inside the tile
        uint32_t it = 0;    // global chunk counter -> stage / phase
        for (long tile = blockIdx.x; tile < n_row_tiles; tile += gridDim.x) {
            const long row = tile * TM + r;
            const bool row_ok = row < P.rows;
            const float *xrow = P.X + (size_t)(row_ok ? row : 0) * P.ldx;
            for (int nt = 0; nt < P.n_tiles; nt++) {
                for (int kc = 0; kc < kc_count; kc++, it++) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    unsigned char *st = stage_base + (size_t)s * stage_bytes;
                    if (tid == 0) {
                        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(2 * b_bytes));
                        const float *src = P.Wp + ((size_t)nt * kc_count + kc) * (2 * (size_t)NT * KC);
                        bulk_g2s(st + 2 * A_TILE_BYTES, src, (uint32_t)(2 * b_bytes), &full_bar[s]);
                    }
                    float *a_hi = reinterpret_cast<float *>(st);
                    float *a_lo = reinterpret_cast<float *>(st + A_TILE_BYTES);
                    const int row_off = (r >> 3) * (KC / 4) * 32 + (r & 7) * 4;  // in floats: (r/8)*SBO + (r%8)*16B
#pragma unroll
                    for (int c4 = 0; c4 < KC / 4; c4++) {
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int k = kc * KC + c4 * 4 + e;
                            float x = 0.f;
                            if (row_ok && k < P.K) {
                                x = __ldg(xrow + k);
                                if (P.mode >= 1) {
                                    x = fmaf(x, __ldg(P.sc + k), __ldg(P.sh + k));
                                    if (P.mode == 2) x += fmaf(__ldg(xrow + P.K + k), __ldg(P.sc + P.K + k), __ldg(P.sh + P.K + k));
                                    x = fmaxf(x, 0.f);
                                }
                            }
                            v[e] = x;
                        }
                        float4 hi, lo;
                        hi.x = to_tf32(v[0]); hi.y = to_tf32(v[1]); hi.z = to_tf32(v[2]); hi.w = to_tf32(v[3]);
                        lo.x = to_tf32(v[0] - hi.x); lo.y = to_tf32(v[1] - hi.y); lo.z = to_tf32(v[2] - hi.z); lo.w = to_tf32(v[3] - hi.w);
                        *reinterpret_cast<float4 *>(a_hi + row_off + c4 * 32) = hi;
                        *reinterpret_cast<float4 *>(a_lo + row_off + c4 * 32) = lo;
                    }
                    fence_proxy_async();
                    mbar_arrive(&full_bar[s]);
                }
            }
        }
    } else if (warp == 4) {
        // =============================== MMA issuer ===============================
        const uint32_t idesc = umma_idesc_tf32(TM, NT);
        uint32_t it = 0, acc_it = 0;
        for (long tile = blockIdx.x; tile < n_row_tiles; tile += gridDim.x) {
            for (int nt = 0; nt < P.n_tiles; nt++, acc_it++) {
                const int ab = acc_it & 1;
                const uint32_t aph = (acc_it >> 1) & 1;
                mbar_wait(&acc_empty[ab], aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(ab * NT_MAX);
                for (int kc = 0; kc < kc_count; kc++, it++) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t st = rsb_smem_addr(stage_base + (size_t)s * stage_bytes);
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
                        umma_commit(&empty_bar[s]);
                        if (kc == kc_count - 1) umma_commit(&acc_full[ab]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // =============================== epilogue ===============================
        const int q = warp & 3;                 // TMEM lane quadrant this warp may access
        const int r = q * 32 + lane;            // row inside the tile
        float *my_tile = stat_tile + (warp - 5) * 32 * 33;
        uint32_t acc_it = 0;
        for (long tile = blockIdx.x; tile < n_row_tiles; tile += gridDim.x) {
            const long row = tile * TM + r;
            const bool row_ok = row < P.rows;
            for (int nt = 0; nt < P.n_tiles; nt++, acc_it++) {
                const int ab = acc_it & 1;
                const uint32_t aph = (acc_it >> 1) & 1;
                mbar_wait(&acc_full[ab], aph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * NT_MAX);
                const int ncols = min(NT, P.N - nt * NT);
                for (int c0 = 0; c0 < ncols; c0 += 32) {
                    float v[32];
                    tmem_ld32(taddr + c0, v);
                    const int n0 = nt * NT + c0;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int n = n0 + j;
                        float y = v[j] + ((P.bias && n < P.N) ? __ldg(P.bias + n) : 0.f);
                        v[j] = (row_ok && n < P.N) ? y : 0.f;
                    }
                    if (row_ok) {
                        float *yrow = P.Y + (size_t)row * P.N + n0;
                        if (n0 + 32 <= P.N && (P.N & 3) == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4 *>(yrow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
                            for (int j = 0; j < 32 && n0 + j < P.N; j++) yrow[j] = v[j];
                        }
                    }
                    if (P.stats) {
                        // column sums over this warp's 32 rows through a padded shared tile, then fp64 atomics
#pragma unroll
                        for (int j = 0; j < 32; j++) my_tile[lane * 33 + j] = v[j];
                        __syncwarp();
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            const float t = my_tile[i * 33 + lane];
                            s1 += t;
                            s2 = fmaf(t, t, s2);
                        }
                        __syncwarp();
                        const int n = n0 + lane;
                        if (n < P.N) {
                            atomicAdd(P.stats + n, (double)s1);
                            atomicAdd(P.stats + P.N + n, (double)s2);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[ab]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

// W [N, K] (row-major, optionally transposed source) -> pre-split canonical chunks
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
        float hi, lo;
        {
            uint32_t r;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(w));
            hi = __uint_as_float(r);
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(w - hi));
            lo = __uint_as_float(r);
        }
        // canonical K-major no-swizzle: core matrix (nn/8, kk/4) at (nn/8)*SBO + (kk/4)*128 B, row (nn%8)*16 B
        const size_t off = (size_t)(nn >> 3) * ((KC / 4) * 32) + (size_t)(kk >> 2) * 32 + (nn & 7) * 4 + (kk & 3);
        float *blk = Wp + ((size_t)nt * k_chunks + kc) * (2 * (size_t)NT * KC);
        blk[off] = hi;
        blk[(size_t)NT * KC + off] = lo;
    }
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

}  // namespace

// Size (in floats) of the pre-split weight buffer for an [N, K] weight.
RSB_EXPORT long rsb_linear_tc_weight_floats(int N, int K)
{
    const int NT = N <= NT_MAX ? round_up(N, 16) : NT_MAX;
    const int n_tiles = (N + NT - 1) / NT;
    const int k_chunks = (K + KC - 1) / KC;
    return (long)n_tiles * k_chunks * 2 * NT * KC;
}

// W: [N, K] row-major with leading dimension ldw (transposed != 0: W is stored [K, N] and used as its transpose).
RSB_EXPORT int rsb_linear_tc_prep_weight(int N, int K, const float *W, int ldw, int transposed, float *Wp,
                                         cudaStream_t stream)
{
    RSB_REQUIRE(N >= 1 && K >= 1, "bad sizes");
    const int NT = N <= NT_MAX ? round_up(N, 16) : NT_MAX;
    const int n_tiles = (N + NT - 1) / NT;
    const int k_chunks = (K + KC - 1) / KC;
    const long total = (long)n_tiles * k_chunks * NT * KC;
    const int grid = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    weight_prep_kernel<<<grid, 256, 0, stream>>>(W, N, K, ldw, transposed, NT, n_tiles, k_chunks, Wp);
    RSB_CHECK_LAUNCH("weight_prep_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

// Y[rows,N] = act(X)[rows,K] @ W[N,K]^T + bias; stats (fp64 [2N], caller-zeroed) += column sum / sum of squares.
// mode 0: act = identity; 1: relu(x*sc+sh), sc/sh [K]; 2: relu(x[:, :K]*sc[:K]+sh[:K] + x[:, K:2K]*sc[K:]+sh[K:]).
RSB_EXPORT int rsb_linear_tc_forward(long rows, int K, int N, const float *X, int ldx, const float *Wp,
                                     const float *bias, int mode, const float *sc, const float *sh, float *Y,
                                     double *stats, cudaStream_t stream)
{
    RSB_REQUIRE(rows >= 0 && K >= 1 && N >= 1, "bad sizes");
    RSB_REQUIRE(mode >= 0 && mode <= 2, "bad mode");
    RSB_REQUIRE(mode == 0 || (sc && sh), "prologue needs scale/shift");
    if (rows == 0) return 0;
    TcParams P = {};
    P.X = X; P.Wp = Wp; P.bias = bias; P.sc = sc; P.sh = sh; P.Y = Y; P.stats = stats;
    P.rows = rows; P.K = K; P.ldx = ldx; P.N = N; P.mode = mode;
    P.NT = N <= NT_MAX ? round_up(N, 16) : NT_MAX;
    P.n_tiles = (N + P.NT - 1) / P.NT;
    P.k_chunks = (K + KC - 1) / KC;
    const size_t smem = STAGES * (2 * (size_t)A_TILE_BYTES + 2 * (size_t)P.NT * KC * 4) + 256 + 4 * 32 * 33 * 4;
    static bool attr_set = false;
    if (!attr_set) {
        RSB_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const long n_row_tiles = (rows + TM - 1) / TM;
    const int grid = (int)(n_row_tiles < rsb_sm_count() ? n_row_tiles : rsb_sm_count());
    linear_tc_kernel<<<grid, THREADS, smem, stream>>>(P);
    RSB_CHECK_LAUNCH("linear_tc_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}
