// sector.cu — the azimuth sector split of the sectorized FPS, on the device.
//
// Replaces the host-side Python of segmentation/modules/pointops/functions/pointops.py:61-93 (per cloud: angle = atan2(x, y),
// `num_sectors` equal-width sectors over [min, max + 1e-4], per sector `torch.where` + `.item()` round trips) and the ~35 torch
// launches the first generation of this package used for it (repeat_interleave, atan2, linspace arithmetic, compare/sum,
// stable radix sort, bincount, cumsum, gather).  Four small kernels, no host synchronisation:
//   1. sector_angle_kernel    angle[i] = atan2f(x, y), per-cloud min / max (ordered-int atomics)
//   2. sector_classify_kernel sector of every point (count of inner edges <= angle, edges as ATen's CPU linspace computes
//                             them in fp32), segment id = first segment of the cloud + sector, per-block histogram
//   3. sector_scan_kernel     exclusive prefix over (segment-major, block-minor): where each block's points of each segment
//                             go; cumulative segment ends (the packed FPS's `offset`), largest segment
//   4. sector_scatter_kernel  stable counting-sort scatter: order[dest] = i, sector_xyz[dest] = xyz[i]
// Output order = what torch.sort(seg_id, stable=True) gives: segment-major, ascending point index inside a segment, which
// is the order the reference builds with torch.where per sector (:88-93).
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int SEC_TPB = 256;
constexpr int SEC_PPT = 8;                        // consecutive rounds of 256 points per block
constexpr int SEC_CHUNK = SEC_TPB * SEC_PPT;      // points per block of the classify / scatter kernels
constexpr int SEC_MAX_SEG = 512;

__device__ __forceinline__ void sec_atomic_min(float *a, float v)
{
    if (v >= 0.f) atomicMin(reinterpret_cast<int *>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}
__device__ __forceinline__ void sec_atomic_max(float *a, float v)
{
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(a), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}

__global__ void sector_init_kernel(int b, float *amin, float *amax)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b) { amin[i] = CUDART_INF_F; amax[i] = -CUDART_INF_F; }
}

// grid (slices, clouds)
__global__ void __launch_bounds__(SEC_TPB) sector_angle_kernel(const float *__restrict__ xyz, const int *__restrict__ offset,
                                                               float *__restrict__ angle, float *__restrict__ amin,
                                                               float *__restrict__ amax)
{
    const int c = blockIdx.y;
    const long beg = c ? __ldg(offset + c - 1) : 0, end = __ldg(offset + c);
    float lo = CUDART_INF_F, hi = -CUDART_INF_F;
    for (long i = beg + blockIdx.x * (long)SEC_TPB + threadIdx.x; i < end; i += (long)gridDim.x * SEC_TPB) {
        const float a = atan2f(__ldg(xyz + i * 3), __ldg(xyz + i * 3 + 1));      // torch.atan2(x, y): pointops.py:69
        angle[i] = a;
        lo = fminf(lo, a);
        hi = fmaxf(hi, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        sec_atomic_min(amin + c, lo);
        sec_atomic_max(amax + c, hi);
    }
}

// cloud of point i: offsets are cumulative ends, b is small (binary search in shared memory)
__device__ __forceinline__ int cloud_of(const int *s_off, int b, int i)
{
    int lo = 0, hi = b - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (i < s_off[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// edge k (0 <= k <= S) of torch.linspace(start, end, S + 1) in fp32 as ATen's CPU kernel evaluates it
// (aten/src/ATen/native/RangeFactories.cpp: step = (end - start) / (steps - 1); i < steps / 2 ? start + step * i
// : end - step * (steps - 1 - i)), every operation rounded separately
__device__ __forceinline__ float linspace_edge(float start, float end, int S, int k)
{
    const float step = __fdiv_rn(__fsub_rn(end, start), (float)S);
    if (k < (S + 1) / 2) return __fadd_rn(start, __fmul_rn(step, (float)k));
    return __fsub_rn(end, __fmul_rn(step, (float)(S - k)));
}

__global__ void __launch_bounds__(SEC_TPB) sector_classify_kernel(int n, int b, int S, int nseg, const float *__restrict__ angle,
                                                                  const int *__restrict__ offset, const int *__restrict__ nsec,
                                                                  const int *__restrict__ seg_first, const float *__restrict__ amin,
                                                                  const float *__restrict__ amax, int *__restrict__ seg_id,
                                                                  int *__restrict__ hist)
{
    extern __shared__ int s_mem[];
    int *s_off = s_mem;                 // [b]
    int *s_hist = s_mem + b;            // [nseg]
    for (int i = threadIdx.x; i < b; i += SEC_TPB) s_off[i] = __ldg(offset + i);
    for (int i = threadIdx.x; i < nseg; i += SEC_TPB) s_hist[i] = 0;
    __syncthreads();
    const int base = blockIdx.x * SEC_CHUNK;
#pragma unroll
    for (int j = 0; j < SEC_PPT; j++) {
        const int i = base + j * SEC_TPB + threadIdx.x;
        if (i < n) {
            const int c = cloud_of(s_off, b, i);
            int sec = 0;
            if (__ldg(nsec + c) > 1) {
                const float a = __ldg(angle + i);
                const float start = __ldg(amin + c), end = __fadd_rn(__ldg(amax + c), 1e-4f);     // pointops.py:73
                for (int k = 1; k < S; k++) sec += (a >= linspace_edge(start, end, S, k)) ? 1 : 0;
            }
            const int sg = __ldg(seg_first + c) + sec;
            seg_id[i] = sg;
            atomicAdd(&s_hist[sg], 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nseg; i += SEC_TPB) hist[(size_t)blockIdx.x * nseg + i] = s_hist[i];
}

// one block; hist[blk][seg] (counts) -> destination of the first point of segment `seg` in block `blk`
__global__ void __launch_bounds__(1024) sector_scan_kernel(int nblocks, int nseg, int *__restrict__ hist, int *__restrict__ sector_offset,
                                                           int *__restrict__ count_max)
{
    __shared__ int s_total[SEC_MAX_SEG], s_start[SEC_MAX_SEG];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int sg = warp; sg < nseg; sg += nwarps) {
        int carry = 0;
        for (int b0 = 0; b0 < nblocks; b0 += 32) {
            const int blk = b0 + lane;
            const int v = blk < nblocks ? hist[(size_t)blk * nseg + sg] : 0;
            int x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (blk < nblocks) hist[(size_t)blk * nseg + sg] = carry + x - v;     // exclusive prefix inside the segment
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) s_total[sg] = carry;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0, mx = 0;
        for (int sg = 0; sg < nseg; sg++) {
            s_start[sg] = run;
            run += s_total[sg];
            mx = max(mx, s_total[sg]);
            sector_offset[sg] = run;
        }
        if (count_max) *count_max = mx;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nblocks * nseg; i += blockDim.x) hist[i] += s_start[i % nseg];
}

__global__ void __launch_bounds__(SEC_TPB) sector_scatter_kernel(int n, int nseg, const float *__restrict__ xyz, const int *__restrict__ seg_id,
                                                                 const int *__restrict__ hist, int *__restrict__ order,
                                                                 float *__restrict__ sector_xyz)
{
    extern __shared__ int s_mem[];
    int *s_run = s_mem;                              // [nseg] next destination of each segment for this block
    int *s_cnt = s_mem + nseg;                       // [8 warps][nseg] points of the current round
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = SEC_TPB / 32;
    for (int i = threadIdx.x; i < nseg; i += SEC_TPB) s_run[i] = __ldg(hist + (size_t)blockIdx.x * nseg + i);
    const int base = blockIdx.x * SEC_CHUNK;
    for (int j = 0; j < SEC_PPT; j++) {
        for (int i = threadIdx.x; i < NW * nseg; i += SEC_TPB) s_cnt[i] = 0;
        __syncthreads();
        const int i = base + j * SEC_TPB + threadIdx.x;
        const bool ok = i < n;
        const int sg = ok ? __ldg(seg_id + i) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, sg);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (ok && rank == 0) s_cnt[warp * nseg + sg] = __popc(peers);
        __syncthreads();
        if (ok) {
            int before = 0;
            for (int w = 0; w < warp; w++) before += s_cnt[w * nseg + sg];
            const int dest = s_run[sg] + before + rank;
            order[dest] = i;
            sector_xyz[(size_t)dest * 3] = __ldg(xyz + (size_t)i * 3);
            sector_xyz[(size_t)dest * 3 + 1] = __ldg(xyz + (size_t)i * 3 + 1);
            sector_xyz[(size_t)dest * 3 + 2] = __ldg(xyz + (size_t)i * 3 + 2);
        }
        __syncthreads();
        for (int s = threadIdx.x; s < nseg; s += SEC_TPB) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) tot += s_cnt[w * nseg + s];
            s_run[s] += tot;
        }
        __syncthreads();
    }
}

__global__ void sector_map_back_kernel(int m, const int *__restrict__ order, const int *__restrict__ idx, long long *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = (long long)__ldg(order + __ldg(idx + i));
}

inline size_t al256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

RSB_EXPORT long rsb_sector_split_workspace_bytes(int n, int b, int nseg)
{
    const size_t nblocks = ((size_t)n + SEC_CHUNK - 1) / SEC_CHUNK;
    return (long)(al256(sizeof(float) * (size_t)n) + al256(sizeof(float) * 2 * (size_t)b) + al256(sizeof(int) * (size_t)n) +
                  al256(sizeof(int) * nblocks * (size_t)nseg));
}

// xyz [n,3] packed, offset [b] cumulative ends; nsec [b] = sectors of each cloud (1 or num_sectors), seg_first [b] = index of
// each cloud's first segment; nseg = sum nsec.  Outputs: order [n] (stable segment-major permutation of 0..n-1), sector_xyz
// [n,3] = xyz[order], sector_offset [nseg] cumulative segment ends, count_max [1] (largest segment, optional).
RSB_EXPORT int rsb_sector_split(int b, int n, int num_sectors, int nseg, const float *xyz, const int *offset, const int *nsec,
                                const int *seg_first, void *workspace, long workspace_bytes, int *order, float *sector_xyz,
                                int *sector_offset, int *count_max, cudaStream_t stream)
{
    RSB_REQUIRE(b >= 1 && n >= 1 && num_sectors >= 1 && nseg >= b && nseg <= SEC_MAX_SEG, "bad sizes (at most 512 segments)");
    RSB_REQUIRE(workspace && workspace_bytes >= rsb_sector_split_workspace_bytes(n, b, nseg), "workspace too small");
    const int nblocks = (n + SEC_CHUNK - 1) / SEC_CHUNK;
    unsigned char *w = static_cast<unsigned char *>(workspace);
    float *angle = reinterpret_cast<float *>(w); w += al256(sizeof(float) * (size_t)n);
    float *amin = reinterpret_cast<float *>(w), *amax = amin + b; w += al256(sizeof(float) * 2 * (size_t)b);
    int *seg_id = reinterpret_cast<int *>(w); w += al256(sizeof(int) * (size_t)n);
    int *hist = reinterpret_cast<int *>(w);
    sector_init_kernel<<<RSB_DIVUP(b, 128), 128, 0, stream>>>(b, amin, amax);
    RSB_CHECK_LAUNCH("sector_init_kernel");
    int slices = RSB_DIVUP(n / b + 1, SEC_TPB * 8);
    if (slices < 1) slices = 1;
    if (slices > 64) slices = 64;
    sector_angle_kernel<<<dim3((unsigned)slices, (unsigned)b), SEC_TPB, 0, stream>>>(xyz, offset, angle, amin, amax);
    RSB_CHECK_LAUNCH("sector_angle_kernel");
    sector_classify_kernel<<<nblocks, SEC_TPB, sizeof(int) * (size_t)(b + nseg), stream>>>(n, b, num_sectors, nseg, angle, offset, nsec,
                                                                                      seg_first, amin, amax, seg_id, hist);
    RSB_CHECK_LAUNCH("sector_classify_kernel");
    sector_scan_kernel<<<1, 1024, 0, stream>>>(nblocks, nseg, hist, sector_offset, count_max);
    RSB_CHECK_LAUNCH("sector_scan_kernel");
    sector_scatter_kernel<<<nblocks, SEC_TPB, sizeof(int) * (size_t)(nseg * (1 + SEC_TPB / 32)), stream>>>(n, nseg, xyz, seg_id, hist, order,
                                                                                                      sector_xyz);
    RSB_CHECK_LAUNCH("sector_scatter_kernel");
    RSB_COUNT_LAUNCH(5);
    return 0;
}

// out[i] = order[idx[i]] as int64: FPS picks inside the sector-major copy -> row ids of the original cloud (pointops.py:105)
RSB_EXPORT int rsb_sector_map_back(int m, const int *order, const int *idx, long long *out, cudaStream_t stream)
{
    if (m <= 0) return 0;
    sector_map_back_kernel<<<RSB_DIVUP(m, 256), 256, 0, stream>>>(m, order, idx, out);
    RSB_CHECK_LAUNCH("sector_map_back_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}
