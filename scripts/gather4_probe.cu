// gather4_probe.cu — one-shot hardware probe for TMA tile::gather4 (cp.async.bulk.tensor.2d ... tile::gather4): which tensor-map
// box shape it wants, how many bytes it completes on the mbarrier, where the four gathered rows land in shared memory and how
// the 128-byte swizzle is applied to them (needed by the gather-fused first-layer GEMM operand of csrc/mlp_tc2.cu).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/gather4_probe scripts/gather4_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled g_encode;

static bool make_map(CUtensorMap *m, const float *base, uint64_t cols, uint64_t rows, uint64_t ld_floats, uint32_t box_cols, uint32_t box_rows,
                     CUtensorMapSwizzle sw)
{
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_floats * 4};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("  cuTensorMapEncodeTiled(box %u x %u) failed: %d\n", box_cols, box_rows, (int)r); return false; }
    return true;
}

__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// out[0] = status (1 ok, 0 timeout), out[1..] = 4 KB of shared memory after the gather
__global__ void probe(const __grid_constant__ CUtensorMap map, int col0, int r0, int r1, int r2, int r3, uint32_t expect_bytes, float *out)
{
    __shared__ __align__(1024) float tile[1024];
    __shared__ __align__(8) uint64_t bar;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tile[i] = -1.f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(&bar)), "r"(expect_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(saddr(tile)), "l"((uint64_t)&map), "r"(col0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(saddr(&bar)) : "memory");
        uint32_t done = 0;
        for (int spin = 0; spin < 2000000 && !done; spin++)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(saddr(&bar)) : "memory");
        out[0] = done ? 1.f : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[1 + i] = tile[i];
}

static void run(const char *name, const float *d_T, int box_cols, int box_rows, CUtensorMapSwizzle sw, uint32_t expect, int ld, int cols)
{
    printf("== %s: box %d x %d, expect_tx %u\n", name, box_cols, box_rows, expect);
    CUtensorMap m;
    if (!make_map(&m, d_T, cols, 64, ld, box_cols, box_rows, sw)) return;
    float *d_out;
    CK(cudaMalloc(&d_out, 1025 * 4));
    probe<<<1, 128>>>(m, 0, 5, 17, 2, 40, expect, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  kernel error: %s\n", cudaGetErrorString(e)); exit(3); }
    std::vector<float> h(1025);
    CK(cudaMemcpy(h.data(), d_out, 1025 * 4, cudaMemcpyDeviceToHost));
    printf("  completed: %s\n", h[0] > 0.5f ? "yes" : "TIMEOUT");
    // where did each source row's elements land?  value = row * 100 + col
    for (int slot = 0; slot < 8; slot++) {          // 8 x 128-byte lines
        printf("  line %d:", slot);
        for (int c = 0; c < 32; c += 4) printf(" %7.0f", h[1 + slot * 32 + c]);
        printf("\n");
    }
    CK(cudaFree(d_out));
}

int main()
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    g_encode = (EncodeTiled)fn;
    const int ld = 32, cols = 20;                         // 20 valid columns of a 32-float pitch (like the sa1 point table)
    std::vector<float> T(64 * ld);
    for (int r = 0; r < 64; r++) for (int c = 0; c < ld; c++) T[r * ld + c] = r * 100.f + c;
    float *d_T;
    CK(cudaMalloc(&d_T, T.size() * 4));
    CK(cudaMemcpy(d_T, T.data(), T.size() * 4, cudaMemcpyHostToDevice));
    run("A: box rows 1, SWIZZLE_128B", d_T, 32, 1, CU_TENSOR_MAP_SWIZZLE_128B, 512, ld, cols);
    // (box rows 4 raises "illegal instruction": the gather4 tensor map must have a one-row box)
    run("C: box rows 1, no swizzle", d_T, 32, 1, CU_TENSOR_MAP_SWIZZLE_NONE, 512, ld, cols);
    run("D: box rows 1, SWIZZLE_128B_ATOM_32B", d_T, 32, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 512, ld, cols);
    printf("done\n");
    return 0;
}
