// harness.cu — device side of the classification evaluation harness.
//
// fps_native_kernel: the torch-native farthest point sampling that `sample()` uses to resample every evaluation cloud
// (classification/modules/pointnet2_utils.py:62-75, :114-124, called with cuda=False on GPU tensors by
// classification/tool/train_cls_scanobjectnn.py:79), with ITS semantics rather than the pointops kernel's:
//   * the first pick is a given start index (the reference draws it with torch.randint on the host),
//   * squared distance as torch evaluates `torch.sum((xyz - centroid) ** 2, -1)`: every operation rounded separately,
//     ((dx*dx + dy*dy) + dz*dz) - no fused multiply-add (the pointops kernels contract to fma, rule R1),
//   * running minimum updated on strict `<`, next pick = first maximum (lowest index) like torch.max.
// Reads the channel-first evaluation batch [B, C, N] directly (coalesced along N), emits the indices AND the resampled
// batch [B, C, m] (the reference's index_points + two permutes).  One CTA per cloud, points resident in registers;
// ~2 block barriers per pick.  The reference launches ~8 torch kernels per pick (1024 picks per batch).
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int HN_TPB = 1024;
constexpr int HN_MAX_PPT = 16;     // clouds of up to 16384 points

template <int PPT>
__global__ void __launch_bounds__(HN_TPB) fps_native_kernel(int n, int c, int m, const float *__restrict__ feat,
                                                            const long long *__restrict__ start, long long *__restrict__ idx_out,
                                                            float *__restrict__ out)
{
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_pick;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *f = feat + (size_t)b * c * n;
    float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
    for (int i = 0; i < PPT; i++) {
        const int k = tid + i * HN_TPB;
        const bool ok = k < n;
        px[i] = ok ? __ldg(f + k) : 0.f;
        py[i] = ok ? __ldg(f + n + k) : 0.f;
        pz[i] = ok ? __ldg(f + 2 * n + k) : 0.f;
        md[i] = ok ? 1e10f : -CUDART_INF_F;             // torch.ones(B, N) * 1e10; padding never wins
    }
    int pick = (int)start[b];
    for (int it = 0; it < m; it++) {
        if (tid == 0) idx_out[(size_t)b * m + it] = pick;
        // gather every channel of the picked point (index_points of the whole feature row)
        for (int ch = tid; ch < c; ch += HN_TPB) out[((size_t)b * c + ch) * m + it] = __ldg(f + (size_t)ch * n + pick);
        const float cx = __ldg(f + pick), cy = __ldg(f + n + pick), cz = __ldg(f + 2 * n + pick);
        float best = -CUDART_INF_F;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < PPT; i++) {
            const float dx = __fsub_rn(px[i], cx), dy = __fsub_rn(py[i], cy), dz = __fsub_rn(pz[i], cz);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d < md[i]) md[i] = d;
            if (md[i] > best) { best = md[i]; bi = tid + i * HN_TPB; }          // ascending index inside the thread
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            best = s_val[lane];
            bi = s_idx[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) s_pick = bi;
        }
        __syncthreads();
        pick = s_pick;
    }
}

}  // namespace

// feat [b, c, n] channel-first with xyz in channels 0..2; start [b] int64 first picks; idx [b, m] int64; out [b, c, m]
RSB_EXPORT int rsb_fps_native_sample(int b, int c, int n, int m, const float *feat, const long long *start, long long *idx, float *out,
                                     cudaStream_t stream)
{
    RSB_REQUIRE(b >= 0 && c >= 3 && n >= 1 && m >= 1 && n <= HN_TPB * HN_MAX_PPT, "bad sizes (clouds of at most 16384 points)");
    if (b == 0) return 0;
    const int ppt = (n + HN_TPB - 1) / HN_TPB;
    switch (ppt) {
#define RSB_HN_CASE(P) case P: fps_native_kernel<P><<<b, HN_TPB, 0, stream>>>(n, c, m, feat, start, idx, out); break;
        RSB_HN_CASE(1) RSB_HN_CASE(2) RSB_HN_CASE(3) RSB_HN_CASE(4)
        case 5: case 6: case 7: case 8: fps_native_kernel<8><<<b, HN_TPB, 0, stream>>>(n, c, m, feat, start, idx, out); break;
        default: fps_native_kernel<16><<<b, HN_TPB, 0, stream>>>(n, c, m, feat, start, idx, out); break;
#undef RSB_HN_CASE
    }
    RSB_CHECK_LAUNCH("fps_native_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}
