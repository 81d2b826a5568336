// common.cuh — shared device/host helpers for the repsurf_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define RSB_DIVUP(a, b) (((a) + (b) - 1) / (b))
#define RSB_EXPORT extern "C" __attribute__((visibility("default")))

// ---- error plumbing -----------------------------------------------------------------------
// Every C-ABI entry returns 0 or a cudaError_t value; the message is kept per thread and read
// back with rsb_last_error().  (The reference's `_fast` launchers print and exit(-1) instead:
// classification/modules/pointops/src/ballquery/ballquery_cuda_kernel.cu:96-100.)
RSB_EXPORT const char *rsb_last_error(void);
void rsb_set_error(const char *fmt, ...);

#define RSB_CHECK_LAUNCH(name)                                               \
    do {                                                                     \
        cudaError_t e__ = cudaGetLastError();                                \
        if (e__ != cudaSuccess) {                                            \
            rsb_set_error("%s: %s", name, cudaGetErrorString(e__));          \
            return (int)e__;                                                 \
        }                                                                    \
    } while (0)

#define RSB_CUDA(call)                                                       \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) {                                            \
            rsb_set_error("%s: %s", #call, cudaGetErrorString(e__));         \
            return (int)e__;                                                 \
        }                                                                    \
    } while (0)

#define RSB_REQUIRE(cond, msg)                                               \
    do {                                                                     \
        if (!(cond)) {                                                       \
            rsb_set_error("%s: %s", __func__, msg);                          \
            return (int)cudaErrorInvalidValue;                               \
        }                                                                    \
    } while (0)

// launch counter (bench.py reports gpu_launches from it)
extern unsigned long long g_rsb_launches;
#define RSB_COUNT_LAUNCH(n) (g_rsb_launches += (n))

int rsb_sm_count();

// ---- exact squared distance (rule R1) -------------------------------------------------------
// nvcc contracts the reference's `dx*dx + dy*dy + dz*dz` into FMUL(y) / FFMA(x) / FFMA(z)
// (SURVEY.md §8(c)); spelled with intrinsics so no optimisation level can change the rounding.
// i / d for 0 <= i, d > 0: the flat element index of the gather / elementwise kernels is a 64-bit `long`, whose
// division is a ~60-instruction software routine; whenever the dividend fits 32 bits (every launch at the workloads'
// sizes except the largest row matrices) the 32-bit hardware path gives the same quotient.
__device__ __forceinline__ long rsb_div(long i, long d)
{
    if (((unsigned long)i | (unsigned long)d) >> 32) return i / d;
    return (long)((unsigned)i / (unsigned)d);
}

__device__ __forceinline__ float rsb_sqdist(float ax, float ay, float az, float bx, float by, float bz)
{
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    float t = __fmul_rn(dy, dy);
    t = __fmaf_rn(dx, dx, t);
    return __fmaf_rn(dz, dz, t);
}

// ReLU that propagates NaN like torch.relu (fmaxf returns the non-NaN operand: a diverged activation would turn into 0)
__device__ __forceinline__ float rsb_relu(float x)
{
    float r;
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(x));
    return r;
}

// ---- cluster / DSMEM primitives -------------------------------------------------------------
__device__ __forceinline__ uint32_t rsb_cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t rsb_cluster_nctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t rsb_cluster_id_x()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void rsb_cluster_arrive_release()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void rsb_cluster_wait_acquire()
{
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t rsb_smem_addr(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
// map a local shared-memory address to the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t rsb_mapa(uint32_t local_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void rsb_st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}
__device__ __forceinline__ void rsb_st_cluster_v2(uint32_t addr, uint32_t a, uint32_t b)
{
    asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
