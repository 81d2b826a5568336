// umma_probe.cu — one-shot hardware probe for the shared-memory layouts the round-2 GEMM kernels rely on:
//   (1) TMA 2-D tile load with SWIZZLE_128B  ->  tcgen05.mma kind::tf32, K-major SWIZZLE_128B descriptors
//   (2) TMA 2-D tile load with SWIZZLE_128B_ATOM_32B -> tcgen05.mma kind::tf32, MN-major descriptors (layout type 1,
//       the only MN-major layout tf32 has): LBO = stride of 32-channel groups, SBO = 512 B (4 k-rows), k-step = 1024 B
//   (3) generic-proxy in-place rewrite of a TMA-loaded tile before the MMA (fence.proxy.async)
//   (4) TMA 2-D tile store from a SWIZZLE_128B staging tile, with out-of-bounds clipping
// Each case is checked against a host computation on tf32-exact inputs and prints PASS / FAIL with the max error.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/umma_probe scripts/umma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled g_encode;

static CUtensorMap make_map(const float *base, uint64_t cols, uint64_t rows, uint64_t ld_floats, uint32_t box_cols, uint32_t box_rows,
                            CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_floats * 4};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(2); }
    return m;
}

__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(saddr(b)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(saddr(dst)), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(saddr(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, const void *src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"((uint64_t)map), "r"(c0), "r"(c1), "r"(saddr(src)) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;     // version
    d |= (uint64_t)layout_type << 61;     // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
    return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool mn_major)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(saddr(bar)) : "memory");
}

struct Params {
    CUtensorMap mapA, mapB, mapD;
    int mode;        // 0: K-major, 1: MN-major
    int layout_type; // descriptor layout type (2 = SW128, 1 = SW128 with 32-byte atoms)
    int lbo, sbo;    // descriptor byte offsets
    int rewrite;     // 1: every thread rewrites the A tile in place (x -> 2x) through the generic proxy before the MMA
    int store_tma;   // 1: result leaves through a swizzled staging tile + TMA store (clipped to mapD), else plain stores
    float *D;        // [128, 64]
    int a_c0, a_c1;  // tile coordinates of A (probe for non-zero origins / OOB fill)
};

constexpr int M = 128, N = 64, KC = 32;

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Params P)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float *sA = (float *)smem;                     // 16 KB
    float *sB = (float *)(smem + 16384);           // 8 KB
    float *sD = (float *)(smem + 16384 + 8192);    // 2 x 16 KB staging (128 rows x 32 floats each)
    __shared__ uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&bar_full, 1);
        mbar_init(&bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(saddr(&tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    if (tid == 0) {
        mbar_expect(&bar_full, (M + N) * KC * 4);
        if (P.mode == 0) {
            tma_load_2d(sA, &P.mapA, P.a_c0, P.a_c1, &bar_full);     // box 32 floats x 128 rows
            tma_load_2d(sB, &P.mapB, 0, 0, &bar_full);               // box 32 floats x 64 rows
        } else {
            for (int g = 0; g < M / 32; g++) tma_load_2d(sA + g * 1024, &P.mapA, g * 32, 0, &bar_full);   // box 32 ch x 32 k-rows
            for (int g = 0; g < N / 32; g++) tma_load_2d(sB + g * 1024, &P.mapB, g * 32, 0, &bar_full);
        }
    }
    mbar_wait(&bar_full, 0);
    if (P.rewrite) {
        float4 *a4 = (float4 *)sA;
        for (int i = tid; i < M * KC / 4; i += 128) { float4 v = a4[i]; v.x *= 2.f; v.y *= 2.f; v.z *= 2.f; v.w *= 2.f; a4[i] = v; }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = idesc_tf32(M, N, P.mode == 1);
        for (int ks = 0; ks < KC / 8; ks++) {
            const uint32_t adv = P.mode == 0 ? ks * 32 : ks * 1024;   // K-major: 8 tf32 inside the 128-byte row; MN-major: 8 k-rows
            umma(tmem, desc_sw128(saddr(sA) + adv, P.lbo, P.sbo, P.layout_type), desc_sw128(saddr(sB) + adv, P.lbo, P.sbo, P.layout_type), idesc, ks ? 1u : 0u);
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!P.store_tma) {
            for (int j = 0; j < 32; j++) P.D[row * N + c0 + j] = __uint_as_float(r[j]);
        } else {
            // staging tile of this 32-column block: row-major 128 B rows, 16-byte chunk c of row r at chunk (c ^ (r & 7))
            float *st = sD + (c0 / 32) * (128 * 32);
            for (int c = 0; c < 8; c++) {
                float4 v = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]), __uint_as_float(r[4 * c + 2]), __uint_as_float(r[4 * c + 3]));
                *(float4 *)(st + row * 32 + ((c ^ (row & 7)) * 4)) = v;
            }
        }
    }
    if (P.store_tma) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            for (int c0 = 0; c0 < N; c0 += 32) tma_store_2d(&P.mapD, c0, warp * 32, sD + (c0 / 32) * (128 * 32) + warp * 32 * 32);   // box 32 cols x 32 rows
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
    }
}

__global__ void dump_kernel(const __grid_constant__ CUtensorMap map, float *out)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect(&bar, 4096);
        tma_load_2d(smem, &map, 0, 0, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = ((float *)smem)[i];
}

static float tfx(int i) { return (float)((i * 37 + 11) % 17 - 8) * 0.25f; }   // tf32-exact values

int main()
{
    cudaDriverEntryPointQueryResult q;
    void *fn = nullptr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
    g_encode = (EncodeTiled)fn;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));

    int fails = 0;
    {   // physical placement of a 32-channel x 32-row box under the 32-byte-atom swizzle: element (r, c) = 100 r + c
        std::vector<float> h(32 * 32), o(1024);
        for (int r = 0; r < 32; r++) for (int c = 0; c < 32; c++) h[r * 32 + c] = 100.f * r + c;
        float *d, *dout;
        CK(cudaMalloc(&d, 4096)); CK(cudaMalloc(&dout, 4096));
        CK(cudaMemcpy(d, h.data(), 4096, cudaMemcpyHostToDevice));
        CUtensorMap m = make_map(d, 32, 32, 32, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        CK(cudaFuncSetAttribute(dump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192));
        dump_kernel<<<1, 128, 8192>>>(m, dout);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(o.data(), dout, 4096, cudaMemcpyDeviceToHost));
        printf("ATOM_32B placement: smem row r, 16-byte chunk p holds channels starting at:\n");
        int bad = 0;
        for (int r = 0; r < 32; r++) {
            if (r < 9) printf("  row %2d:", r);
            for (int p = 0; p < 8; p++) {
                const float v = o[r * 32 + p * 4];
                const int rr = (int)(v / 100.f), c = (int)(v - 100.f * rr);
                if (r < 9) printf(" %2d%s", c, rr == r ? "" : "!");
                if (rr != r || c != ((p ^ ((r & 3) << 1)) * 4)) bad++;
            }
            if (r < 9) printf("\n");
        }
        printf("%s placement == chunk ^ ((row & 3) << 1) : %d mismatches\n", bad ? "FAIL" : "PASS", bad);
        cudaFree(d); cudaFree(dout);
    }
    // ---------------- K-major cases: A [rowsA, ldA] logical K = Kv columns, B [64, 32]
    struct KCase { const char *name; int lbo, sbo, rewrite, store, Kv, rowsA, a_c1, rowsD, colsD; };
    KCase kc[] = {
        {"K-major lbo=0    sbo=1024", 0, 1024, 0, 0, 32, 128, 0, 128, 64},
        {"K-major lbo=16   sbo=1024", 16, 1024, 0, 0, 32, 128, 0, 128, 64},
        {"K-major lbo=1024 sbo=1024", 1024, 1024, 0, 0, 32, 128, 0, 128, 64},
        {"K-major K=20 (OOB cols zero-filled)", 0, 1024, 0, 0, 20, 128, 0, 128, 64},
        {"K-major 100 valid rows, tile origin row 128 (OOB rows zero-filled)", 0, 1024, 0, 0, 32, 228, 128, 128, 64},
        {"K-major in-place rewrite x2", 0, 1024, 1, 0, 32, 128, 0, 128, 64},
        {"K-major TMA store", 0, 1024, 0, 1, 32, 128, 0, 128, 64},
        {"K-major TMA store clipped to 100 rows x 40 cols", 0, 1024, 0, 1, 32, 128, 0, 100, 40},
    };
    for (auto &c : kc) {
        const int ldA = 48;
        std::vector<float> hA((size_t)c.rowsA * ldA), hB(N * KC), hD(M * N, -777.f);
        for (size_t i = 0; i < hA.size(); i++) hA[i] = tfx((int)i);
        for (int i = 0; i < N * KC; i++) hB[i] = tfx(i * 3 + 1);
        float *dA, *dB, *dD;
        CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, M * N * 4));
        CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dD, hD.data(), M * N * 4, cudaMemcpyHostToDevice));
        Params P;
        P.mapA = make_map(dA, c.Kv, c.rowsA, ldA, 32, 128);
        P.mapB = make_map(dB, KC, N, KC, 32, 64);
        P.mapD = make_map(dD, c.colsD, c.rowsD, N, 32, 32);
        P.mode = 0; P.layout_type = 2; P.lbo = c.lbo; P.sbo = c.sbo; P.rewrite = c.rewrite; P.store_tma = c.store; P.D = dD; P.a_c0 = 0; P.a_c1 = c.a_c1;
        probe_kernel<<<1, 128, 80 * 1024>>>(P);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("FAIL %-70s : %s\n", c.name, cudaGetErrorString(e)); return 3; }
        CK(cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int m = 0; m < M; m++)
            for (int n = 0; n < N; n++) {
                double ref = 0;
                const int gr = c.a_c1 + m;
                for (int k = 0; k < KC; k++) {
                    const float a = (gr < c.rowsA && k < c.Kv) ? hA[(size_t)gr * ldA + k] * (c.rewrite ? 2.f : 1.f) : 0.f;
                    ref += (double)a * hB[n * KC + k];
                }
                if (c.store && (m >= c.rowsD || n >= c.colsD)) ref = -777.0;   // clipped: untouched
                const double err = fabs(ref - hD[m * N + n]);
                if (err > maxerr) maxerr = err;
            }
        printf("%s %-70s : max err %.3g\n", maxerr < 1e-3 ? "PASS" : "FAIL", c.name, maxerr);
        if (maxerr >= 1e-3) fails++;
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    // ---------------- MN-major cases: G [32 rows, ldG >= 128 ch], X [32 rows, 64 ch];  D[m, n] = sum_k G[k, m] X[k, n]
    struct MCase { const char *name; int lbo, sbo, Mv, lt; CUtensorMapSwizzle sw; };
    MCase mc[] = {
        {"MN-major atom32B lt=1 lbo=4096 sbo=512", 4096, 512, 128, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B},
        {"MN-major atom32B lt=1 lbo=512 sbo=4096", 512, 4096, 128, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B},
        {"MN-major atom32B lt=1 lbo=4096 sbo=1024", 4096, 1024, 128, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B},
        {"MN-major atom32B lt=1 lbo=4096 sbo=512, 72 valid channels (OOB zero)", 4096, 512, 72, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B},
        {"MN-major atom32B-flip8B lt=1 lbo=4096 sbo=512", 4096, 512, 128, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B_FLIP_8B},
        {"MN-major sw128 lt=1 lbo=4096 sbo=512", 4096, 512, 128, 1, CU_TENSOR_MAP_SWIZZLE_128B},
        {"MN-major atom32B lt=2 lbo=4096 sbo=1024", 4096, 1024, 128, 2, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B},
    };
    for (auto &c : mc) {
        const int ldG = 136, ldX = 64;
        std::vector<float> hG(KC * ldG), hX(KC * ldX), hD(M * N, -777.f);
        for (size_t i = 0; i < hG.size(); i++) hG[i] = tfx((int)i * 5 + 2);
        for (size_t i = 0; i < hX.size(); i++) hX[i] = tfx((int)i * 7 + 3);
        float *dG, *dX, *dD;
        CK(cudaMalloc(&dG, hG.size() * 4)); CK(cudaMalloc(&dX, hX.size() * 4)); CK(cudaMalloc(&dD, M * N * 4));
        CK(cudaMemcpy(dG, hG.data(), hG.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dX, hX.data(), hX.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dD, hD.data(), M * N * 4, cudaMemcpyHostToDevice));
        Params P;
        P.mapA = make_map(dG, c.Mv, KC, ldG, 32, 32, c.sw);
        P.mapB = make_map(dX, N, KC, ldX, 32, 32, c.sw);
        P.mapD = make_map(dD, N, M, N, 32, 32);
        P.mode = 1; P.layout_type = c.lt; P.lbo = c.lbo; P.sbo = c.sbo; P.rewrite = 0; P.store_tma = 0; P.D = dD; P.a_c0 = 0; P.a_c1 = 0;
        probe_kernel<<<1, 128, 80 * 1024>>>(P);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("FAIL %-70s : %s\n", c.name, cudaGetErrorString(e)); return 3; }
        CK(cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int m = 0; m < M; m++)
            for (int n = 0; n < N; n++) {
                double ref = 0;
                for (int k = 0; k < KC; k++) ref += (double)(m < c.Mv ? hG[k * ldG + m] : 0.f) * hX[k * ldX + n];
                const double err = fabs(ref - hD[m * N + n]);
                if (err > maxerr) maxerr = err;
            }
        printf("%s %-70s : max err %.3g\n", maxerr < 1e-3 ? "PASS" : "FAIL", c.name, maxerr);
        if (maxerr >= 1e-3) fails++;
        cudaFree(dG); cudaFree(dX); cudaFree(dD);
    }
    printf("probe done, %d failing variants (some variants are EXPECTED to fail: they bracket the encoding)\n", fails);
    return 0;
}
