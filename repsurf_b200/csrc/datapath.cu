// datapath.cu — device side of the segmentation input pipeline: grid subsampling and nearest-crop, the two
// per-cloud steps that run in front of the packed layout.
//   segmentation/modules/voxelize_utils.py:4-17    fnv_hash_vec   ("FNV64-1A": h = basis; per column h *= prime, h ^= value)
//   segmentation/modules/voxelize_utils.py:38-58   voxelize       (floor(coord / voxel_size) -> key -> sort -> one point per voxel)
//   segmentation/util/data_util.py:45-51           nearest crop   (voxel_max points closest to a seed point)
//   segmentation/modules/voxelize_utils.py:20-35   ravel_hash_vec (row-major rank of the voxel inside the occupied bounding box)
//   segmentation/tool/test_s3dis.py:131-159        data_process   (overlapping nearest crops until every point is covered)
// Kernels: per-column minimum / maximum, voxel keys (FNV and ravel), run detection over the sorted keys (flags -> block scan ->
// compaction), the per-voxel pick, squared distance to a seed (host- or device-resident seed index), first-occurrence argmin of
// a float64 array, the coverage update of the crop planner.  The 64-bit key sort and the distance sort between them are cub radix sorts
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

// float atomics through the integer orderings of the IEEE bit patterns.  `v + 0.f` turns -0.0 into +0.0 first: the pattern of
// -0.0 is INT_MIN, which would win a signed atomicMin against every negative number (and lose every atomicMax)
__device__ __forceinline__ void dp_atomic_min(float *a, float v)
{
    v = __fadd_rn(v, 0.f);
    if (v >= 0.f) atomicMin(reinterpret_cast<int *>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}

__device__ __forceinline__ void dp_atomic_max(float *a, float v)
{
    v = __fadd_rn(v, 0.f);
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(a), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
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

// out3 (pre-set to -inf) = column maxima of coord [n,3]
__global__ void __launch_bounds__(DP_TPB) coord_max_kernel(long n, const float *__restrict__ coord, float *__restrict__ out3)
{
    float hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (long i = blockIdx.x * (long)DP_TPB + threadIdx.x; i < n; i += (long)gridDim.x * DP_TPB)
#pragma unroll
        for (int a = 0; a < 3; a++) hi[a] = fmaxf(hi[a], __ldg(coord + i * 3 + a));
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        if ((threadIdx.x & 31) == 0 && hi[a] > -CUDART_INF_F) dp_atomic_max(out3 + a, hi[a]);
    }
}

// ravel_hash_vec (voxelize_utils.py:20-35): d = floor(coord / voxel_size) - min over the cloud of the same (floor and the fp32
// division are monotone, so that minimum is floor(cmin / voxel_size), likewise the maximum); extent e_a = max d_a + 1;
// key = (d_0 * e_1 + d_1) * e_2 + d_2 in uint64.  Keys of a real cloud stay far below 2^63, so they sort as signed numbers.
__global__ void __launch_bounds__(DP_TPB) voxel_key_ravel_kernel(long n, const float *__restrict__ coord, const float *__restrict__ cmin,
                                                                 const float *__restrict__ cmax, float voxel_size,
                                                                 long long *__restrict__ key)
{
    float lo[3];
    unsigned long long ext[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        lo[a] = floorf(__fdiv_rn(__ldg(cmin + a), voxel_size));
        ext[a] = (unsigned long long)__fsub_rn(floorf(__fdiv_rn(__ldg(cmax + a), voxel_size)), lo[a]) + 1ull;
    }
    for (long i = blockIdx.x * (long)DP_TPB + threadIdx.x; i < n; i += (long)gridDim.x * DP_TPB) {
        unsigned long long d[3];
#pragma unroll
        for (int a = 0; a < 3; a++)
            d[a] = (unsigned long long)__fsub_rn(floorf(__fdiv_rn(__ldg(coord + i * 3 + a), voxel_size)), lo[a]);
        key[i] = (long long)((d[0] * ext[1] + d[1]) * ext[2] + d[2]);
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

// same with the seed index read from device memory (the crop planner's argmin result never visits the host)
__global__ void __launch_bounds__(DP_TPB) seed_dist_dev_kernel(long n, const float *__restrict__ coord,
                                                               const unsigned long long *__restrict__ seed_p, float *__restrict__ d)
{
    const long seed = (long)__ldg(seed_p);
    if (seed < 0 || seed >= n) return;
    const float sx = __ldg(coord + seed * 3), sy = __ldg(coord + seed * 3 + 1), sz = __ldg(coord + seed * 3 + 2);
    for (long i = blockIdx.x * (long)DP_TPB + threadIdx.x; i < n; i += (long)gridDim.x * DP_TPB) {
        const float dx = __fsub_rn(__ldg(coord + i * 3), sx), dy = __fsub_rn(__ldg(coord + i * 3 + 1), sy),
                    dz = __fsub_rn(__ldg(coord + i * 3 + 2), sz);
        d[i] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    }
}

// ---- first-occurrence argmin of a float64 array (np.argmin), two passes over work[2] = {smallest key, its first index} -------
// order-preserving map of a double onto uint64: negative -> all bits flipped, non-negative -> sign bit set
__device__ __forceinline__ unsigned long long dp_f64_key(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(DP_TPB) argmin_key_kernel(long n, const double *__restrict__ v, unsigned long long *__restrict__ work)
{
    unsigned long long best = ~0ull;
    for (long i = blockIdx.x * (long)DP_TPB + threadIdx.x; i < n; i += (long)gridDim.x * DP_TPB) {
        const unsigned long long k = dp_f64_key(v[i]);
        best = k < best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) atomicMin(work, best);
}

__global__ void __launch_bounds__(DP_TPB) argmin_index_kernel(long n, const double *__restrict__ v, unsigned long long *__restrict__ work)
{
    const unsigned long long want = work[0];
    for (long i = blockIdx.x * (long)DP_TPB + threadIdx.x; i < n; i += (long)gridDim.x * DP_TPB)
        if (dp_f64_key(v[i]) == want) atomicMin(work + 1, (unsigned long long)i);
}

// ---- coverage update of the crop planner (test_s3dis.py:146-153) -------------------------------------------------------------
// crop [m] = the rows of this crop in ascending distance, dist [n] = squared distances of ALL rows to the seed:
// priority[crop] += (1 - dist[crop] / max(dist[crop]))^2 (fp32 arithmetic, added to the float64 priorities like numpy does),
// covered[crop] = 1, *n_covered += rows covered for the first time.  Rows of one crop are distinct.
__global__ void __launch_bounds__(DP_TPB) crop_update_kernel(int m, const long long *__restrict__ crop, const float *__restrict__ dist,
                                                             double *__restrict__ priority, int *__restrict__ covered,
                                                             int *__restrict__ n_covered)
{
    const float dmax = __ldg(dist + crop[m - 1]);
    int fresh = 0;
    for (int j = blockIdx.x * DP_TPB + threadIdx.x; j < m; j += gridDim.x * DP_TPB) {
        const long long r = crop[j];
        const float t = __fsub_rn(1.f, __fdiv_rn(__ldg(dist + r), dmax));
        priority[r] += (double)__fmul_rn(t, t);
        if (covered[r] == 0) {
            covered[r] = 1;
            fresh++;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fresh += __shfl_xor_sync(0xffffffffu, fresh, o);
    if ((threadIdx.x & 31) == 0 && fresh) atomicAdd(n_covered, fresh);
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

RSB_EXPORT int rsb_coord_max(long n, const float *coord, float *out3, cudaStream_t stream)
{
    if (n <= 0) return 0;
    coord_max_kernel<<<dp_grid(n), DP_TPB, 0, stream>>>(n, coord, out3);
    RSB_CHECK_LAUNCH("coord_max_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_voxel_keys_ravel(long n, const float *coord, const float *cmin, const float *cmax, float voxel_size, long long *key,
                                    cudaStream_t stream)
{
    RSB_REQUIRE(voxel_size > 0.f, "voxel_size must be positive");
    if (n <= 0) return 0;
    voxel_key_ravel_kernel<<<dp_grid(n), DP_TPB, 0, stream>>>(n, coord, cmin, cmax, voxel_size, key);
    RSB_CHECK_LAUNCH("voxel_key_ravel_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_seed_distance_dev(long n, const float *coord, const unsigned long long *seed, float *dist, cudaStream_t stream)
{
    RSB_REQUIRE(n >= 1 && seed != nullptr, "bad seed");
    seed_dist_dev_kernel<<<dp_grid(n), DP_TPB, 0, stream>>>(n, coord, seed, dist);
    RSB_CHECK_LAUNCH("seed_dist_dev_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

// work: two uint64 on the device (any content); on return work[1] = index of the first smallest element of v [n] (NaN-free input)
RSB_EXPORT int rsb_argmin_f64(long n, const double *v, unsigned long long *work, cudaStream_t stream)
{
    RSB_REQUIRE(n >= 1 && work != nullptr, "bad size");
    RSB_CUDA(cudaMemsetAsync(work, 0xFF, 2 * sizeof(unsigned long long), stream));
    argmin_key_kernel<<<dp_grid(n), DP_TPB, 0, stream>>>(n, v, work);
    RSB_CHECK_LAUNCH("argmin_key_kernel");
    argmin_index_kernel<<<dp_grid(n), DP_TPB, 0, stream>>>(n, v, work);
    RSB_CHECK_LAUNCH("argmin_index_kernel");
    RSB_COUNT_LAUNCH(2);
    return 0;
}

RSB_EXPORT int rsb_crop_update(int m, const long long *crop, const float *dist, double *priority, int *covered, int *n_covered,
                               cudaStream_t stream)
{
    if (m <= 0) return 0;
    crop_update_kernel<<<dp_grid(m), DP_TPB, 0, stream>>>(m, crop, dist, priority, covered, n_covered);
    RSB_CHECK_LAUNCH("crop_update_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}
