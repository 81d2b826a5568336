// datapath.cu — device side of the segmentation input pipeline: grid subsampling and nearest-crop, the two
// per-cloud steps that run in front of the packed layout.
//   segmentation/modules/voxelize_utils.py:4-17    fnv_hash_vec   ("FNV64-1A": h = basis; per column h *= prime, h ^= value)
//   segmentation/modules/voxelize_utils.py:38-58   voxelize       (floor(coord / voxel_size) -> key -> sort -> one point per voxel)
//   segmentation/util/data_util.py:45-51           nearest crop   (voxel_max points closest to a seed point)
// Kernels: per-column minimum, voxel keys, run detection over the sorted keys (flags -> block scan -> compaction), the
// per-voxel pick, squared distance to a seed.  The 64-bit key sort and the distance sort between them are cub radix sorts
// (through torch.sort, stable) - the one library primitive of this path.  All kernels are single-pass and HBM-bound.
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int DP_TPB = 256;
constexpr int DP_SCAN = 1024;

inline int dp_grid(long work)
{
    long b = (work + DP_TPB - 1) / DP_TPB;
    const long cap = (long)rsb_sm_count() * 16;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

__device__ __forceinline__ void dp_atomic_min(float *a, float v)
{
    if (v >= 0.f) atomicMin(reinterpret_cast<int *>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}

// out3 (pre-set to +inf) = column minima of coord [n,3]
__global__ void __launch_bounds__(DP_TPB) coord_min_kernel(long n, const float *__restrict__ coord, float *__restrict__ out3)
{
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F};
    for (long i = blockIdx.x * (long)DP_TPB + threadIdx.x; i < n; i += (long)gridDim.x * DP_TPB)
#pragma unroll
        for (int a = 0; a < 3; a++) lo[a] = fminf(lo[a], __ldg(coord + i * 3 + a));
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
        if ((threadIdx.x & 31) == 0 && lo[a] < CUDART_INF_F) dp_atomic_min(out3 + a, lo[a]);
    }
}

// key[i] = FNV64-1A over floor((coord[i] - cmin) / voxel_size) as uint64, stored with the sign bit flipped so that a signed
// 64-bit sort orders the keys as unsigned numbers (numpy sorts uint64)
__global__ void __launch_bounds__(DP_TPB) voxel_key_kernel(long n, const float *__restrict__ coord, const float *__restrict__ cmin,
                                                           float voxel_size, long long *__restrict__ key)
{
    const float m0 = __ldg(cmin), m1 = __ldg(cmin + 1), m2 = __ldg(cmin + 2);
    for (long i = blockIdx.x * (long)DP_TPB + threadIdx.x; i < n; i += (long)gridDim.x * DP_TPB) {
        const float d[3] = {floorf(__fdiv_rn(__fsub_rn(__ldg(coord + i * 3), m0), voxel_size)),
                            floorf(__fdiv_rn(__fsub_rn(__ldg(coord + i * 3 + 1), m1), voxel_size)),
                            floorf(__fdiv_rn(__fsub_rn(__ldg(coord + i * 3 + 2), m2), voxel_size))};
        unsigned long long h = 14695981039346656037ull;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            h *= 1099511628211ull;
            h ^= (unsigned long long)d[a];
        }
        key[i] = (long long)(h ^ 0x8000000000000000ull);
    }
}

// ---- runs of equal keys in the sorted sequence -------------------------------------------------------------------------------
__global__ void __launch_bounds__(DP_SCAN) run_count_kernel(long n, const long long *__restrict__ key, int *__restrict__ block_sum)
{
    __shared__ int s[32];
    const long i = blockIdx.x * (long)DP_SCAN + threadIdx.x;
    const int f = (i < n && (i == 0 || key[i] != key[i - 1])) ? 1 : 0;
    int x = f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        x = s[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (threadIdx.x == 0) block_sum[blockIdx.x] = x;
    }
}

// one block: exclusive scan of the block sums in place; total -> *n_runs
__global__ void __launch_bounds__(DP_SCAN) run_scan_kernel(int nblocks, int *__restrict__ block_sum, int *__restrict__ n_runs)
{
    __shared__ int s[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += DP_SCAN) {
        const int i = b0 + threadIdx.x;
        const int v = i < nblocks ? block_sum[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            s[threadIdx.x] = w;
        }
        __syncthreads();
        const int before = carry_s + ((threadIdx.x >> 5) ? s[(threadIdx.x >> 5) - 1] : 0) + x - v;
        if (i < nblocks) block_sum[i] = before;
        __syncthreads();
        if (threadIdx.x == DP_SCAN - 1) carry_s = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_runs = carry_s;
}

// start[r] = first position of run r in the sorted sequence
__global__ void __launch_bounds__(DP_SCAN) run_compact_kernel(long n, const long long *__restrict__ key, const int *__restrict__ block_off,
                                                              int *__restrict__ start)
{
    __shared__ int s[32];
    const long i = blockIdx.x * (long)DP_SCAN + threadIdx.x;
    const int f = (i < n && (i == 0 || key[i] != key[i - 1])) ? 1 : 0;
    int x = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        int w = s[threadIdx.x];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if (threadIdx.x >= o) w += y;
        }
        s[threadIdx.x] = w;
    }
    __syncthreads();
    if (f) start[block_off[blockIdx.x] + ((threadIdx.x >> 5) ? s[(threadIdx.x >> 5) - 1] : 0) + x - 1] = (int)i;
}

// count[r] = start[r+1] - start[r] (last: n - start); *count_max = max count
__global__ void __launch_bounds__(DP_TPB) run_len_kernel(int n_runs, long n, const int *__restrict__ start, int *__restrict__ count,
                                                         int *__restrict__ count_max)
{
    int mx = 0;
    for (int r = blockIdx.x * DP_TPB + threadIdx.x; r < n_runs; r += gridDim.x * DP_TPB) {
        const int c = (r + 1 < n_runs ? start[r + 1] : (int)n) - start[r];
        count[r] = c;
        mx = max(mx, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(count_max, mx);
}

// out[r] = order[start[r] + draw[r] % count[r]]   (voxelize_utils.py:53-55, train mode)
__global__ void __launch_bounds__(DP_TPB) voxel_pick_kernel(int n_runs, const int *__restrict__ start, const int *__restrict__ count,
                                                            const long long *__restrict__ draw, const long long *__restrict__ order,
                                                            long long *__restrict__ out)
{
    for (int r = blockIdx.x * DP_TPB + threadIdx.x; r < n_runs; r += gridDim.x * DP_TPB)
        out[r] = order[start[r] + (int)(draw[r] % (long long)count[r])];
}

// d[i] = sum((coord[i] - coord[seed])^2): numpy's np.sum(np.square(.), 1) in fp32, every operation rounded separately
__global__ void __launch_bounds__(DP_TPB) seed_dist_kernel(long n, const float *__restrict__ coord, long seed, float *__restrict__ d)
{
    const float sx = __ldg(coord + seed * 3), sy = __ldg(coord + seed * 3 + 1), sz = __ldg(coord + seed * 3 + 2);
    for (long i = blockIdx.x * (long)DP_TPB + threadIdx.x; i < n; i += (long)gridDim.x * DP_TPB) {
        const float dx = __fsub_rn(__ldg(coord + i * 3), sx), dy = __fsub_rn(__ldg(coord + i * 3 + 1), sy),
                    dz = __fsub_rn(__ldg(coord + i * 3 + 2), sz);
        d[i] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    }
}

}  // namespace

RSB_EXPORT int rsb_coord_min(long n, const float *coord, float *out3, cudaStream_t stream)
{
    if (n <= 0) return 0;
    coord_min_kernel<<<dp_grid(n), DP_TPB, 0, stream>>>(n, coord, out3);
    RSB_CHECK_LAUNCH("coord_min_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_voxel_keys(long n, const float *coord, const float *cmin, float voxel_size, long long *key, cudaStream_t stream)
{
    RSB_REQUIRE(voxel_size > 0.f, "voxel_size must be positive");
    if (n <= 0) return 0;
    voxel_key_kernel<<<dp_grid(n), DP_TPB, 0, stream>>>(n, coord, cmin, voxel_size, key);
    RSB_CHECK_LAUNCH("voxel_key_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

// sorted keys [n] -> start [>= n_runs entries, caller allocates n], scalars[0] = number of runs; scratch: ceil(n / 1024) ints
RSB_EXPORT int rsb_voxel_runs(long n, const long long *sorted_key, int *scratch, int *start, int *scalars, cudaStream_t stream)
{
    RSB_REQUIRE(n >= 1 && n < (1L << 31), "bad size");
    const int nblocks = (int)((n + DP_SCAN - 1) / DP_SCAN);
    run_count_kernel<<<nblocks, DP_SCAN, 0, stream>>>(n, sorted_key, scratch);
    RSB_CHECK_LAUNCH("run_count_kernel");
    run_scan_kernel<<<1, DP_SCAN, 0, stream>>>(nblocks, scratch, scalars);
    RSB_CHECK_LAUNCH("run_scan_kernel");
    run_compact_kernel<<<nblocks, DP_SCAN, 0, stream>>>(n, sorted_key, scratch, start);
    RSB_CHECK_LAUNCH("run_compact_kernel");
    RSB_COUNT_LAUNCH(3);
    return 0;
}

// count [n_runs], count_max [1] (pre-zeroed)
RSB_EXPORT int rsb_voxel_counts(int n_runs, long n, const int *start, int *count, int *count_max, cudaStream_t stream)
{
    if (n_runs <= 0) return 0;
    run_len_kernel<<<dp_grid(n_runs), DP_TPB, 0, stream>>>(n_runs, n, start, count, count_max);
    RSB_CHECK_LAUNCH("run_len_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_voxel_pick(int n_runs, const int *start, const int *count, const long long *draw, const long long *order,
                              long long *out, cudaStream_t stream)
{
    if (n_runs <= 0) return 0;
    voxel_pick_kernel<<<dp_grid(n_runs), DP_TPB, 0, stream>>>(n_runs, start, count, draw, order, out);
    RSB_CHECK_LAUNCH("voxel_pick_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_seed_distance(long n, const float *coord, long seed, float *dist, cudaStream_t stream)
{
    RSB_REQUIRE(n >= 1 && seed >= 0 && seed < n, "bad seed");
    seed_dist_kernel<<<dp_grid(n), DP_TPB, 0, stream>>>(n, coord, seed, dist);
    RSB_CHECK_LAUNCH("seed_dist_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}
