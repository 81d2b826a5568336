// scene.cu — device side of the whole-scene inference loop (segmentation/tool/test_s3dis.py:186-238):
//   * vote accumulation: pred[idx[r], :] += softmax(logits[r, :]), count[idx[r]] += 1          (test_s3dis.py:208-213)
//   * decision: argmax_c pred[p, c] / count[p]                                                  (test_s3dis.py:217)
//   * kNN median filter of the decided labels: median over the labels of the k nearest points   (util/utils.py:235-245;
//     the k nearest come from rsb_knnquery_grid on the whole scene as ONE segment of ~10^6 points)
// One pass each over [rows, classes] / [points, k]: HBM-bound, coalesced along the class / neighbour dimension.
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int SC_TPB = 256;
constexpr int SC_MAX_K = 128;

// warp per row: softmax over `nc` classes (fp32, max-subtracted like torch.softmax), atomics into the scene-wide accumulators
// (crops of one batch overlap, so two rows may vote for the same point)
__global__ void __launch_bounds__(SC_TPB) vote_kernel(long rows, int nc, const float *__restrict__ logits, int ld,
                                                      const long long *__restrict__ idx, float *__restrict__ pred,
                                                      float *__restrict__ count)
{
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)SC_TPB + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * SC_TPB) >> 5;
    for (long r = warp; r < rows; r += nwarps) {
        const float *x = logits + r * ld;
        float mx = -CUDART_INF_F;
        for (int c = lane; c < nc; c += 32) mx = fmaxf(mx, __ldg(x + c));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int c = lane; c < nc; c += 32) sum += expf(__ldg(x + c) - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const long p = idx[r];
        for (int c = lane; c < nc; c += 32) atomicAdd(pred + p * nc + c, expf(__ldg(x + c) - mx) / sum);
        if (lane == 0) atomicAdd(count + p, 1.f);
    }
}

// label[p] = argmax_c pred[p, c] / count[p]  (first maximum, like numpy.argmax; count 0 -> NaN rows -> class 0 as numpy does)
__global__ void __launch_bounds__(SC_TPB) decide_kernel(long n, int nc, const float *__restrict__ pred, const float *__restrict__ count,
                                                        int *__restrict__ label)
{
    for (long p = blockIdx.x * (long)SC_TPB + threadIdx.x; p < n; p += (long)gridDim.x * SC_TPB) {
        const float cnt = __ldg(count + p);
        float best = -CUDART_INF_F;
        int bi = 0;
        for (int c = 0; c < nc; c++) {
            const float v = __fdiv_rn(__ldg(pred + p * nc + c), cnt);
            if (v > best || (v != v && c == 0)) { best = v; bi = c; }      // NaN in front wins in numpy.argmax
            if (best != best) break;
        }
        label[p] = bi;
    }
}

// out[p] = lower median of label[nbr[p, 0..k)]  (torch.median of an even-sized row returns the smaller middle element)
__global__ void __launch_bounds__(SC_TPB) label_median_kernel(long n, int k, const int *__restrict__ nbr, const int *__restrict__ label,
                                                              int *__restrict__ out)
{
    const int r = (k - 1) / 2;
    for (long p = blockIdx.x * (long)SC_TPB + threadIdx.x; p < n; p += (long)gridDim.x * SC_TPB) {
        int v[SC_MAX_K];
        for (int i = 0; i < k; i++) v[i] = __ldg(label + __ldg(nbr + p * k + i));
        int med = v[0];
        for (int i = 0; i < k; i++) {
            int less = 0, leq = 0;
            for (int j = 0; j < k; j++) { less += v[j] < v[i]; leq += v[j] <= v[i]; }
            if (less <= r && r < leq) { med = v[i]; break; }
        }
        out[p] = med;
    }
}

inline int sc_grid(long work)
{
    long b = (work + SC_TPB - 1) / SC_TPB;
    const long cap = (long)rsb_sm_count() * 16;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

RSB_EXPORT int rsb_scene_vote(long rows, int num_class, const float *logits, int ld, const long long *idx, float *pred, float *count,
                              cudaStream_t stream)
{
    RSB_REQUIRE(rows >= 0 && num_class >= 1 && ld >= num_class, "bad sizes");
    if (rows == 0) return 0;
    vote_kernel<<<sc_grid(rows * 32), SC_TPB, 0, stream>>>(rows, num_class, logits, ld, idx, pred, count);
    RSB_CHECK_LAUNCH("vote_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_scene_decide(long n, int num_class, const float *pred, const float *count, int *label, cudaStream_t stream)
{
    RSB_REQUIRE(n >= 0 && num_class >= 1, "bad sizes");
    if (n == 0) return 0;
    decide_kernel<<<sc_grid(n), SC_TPB, 0, stream>>>(n, num_class, pred, count, label);
    RSB_CHECK_LAUNCH("decide_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_label_median(long n, int k, const int *nbr, const int *label, int *out, cudaStream_t stream)
{
    RSB_REQUIRE(n >= 0 && k >= 1 && k <= SC_MAX_K, "1 <= k <= 128");
    if (n == 0) return 0;
    label_median_kernel<<<sc_grid(n), SC_TPB, 0, stream>>>(n, k, nbr, label, out);
    RSB_CHECK_LAUNCH("label_median_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

// ---- cross-entropy over rows (the criterion of the segmentation step: segmentation/tool/train.py uses nn.CrossEntropyLoss with
// ignore_index) in ONE pass: per row log-sum-exp over the classes, loss_sum += -(x[t] - lse), and the gradient
// softmax(x) - onehot(t) written at once (scaled by 1 / #valid rows in the backward).  torch runs log_softmax + a single-block
// nll reduction (0.28 ms for 327 680 rows) + two backward kernels.
namespace {
__global__ void __launch_bounds__(256) cross_entropy_kernel(long rows, int nc, const float *__restrict__ logits, int ld,
                                                            const long long *__restrict__ target, long long ignore_index,
                                                            float *__restrict__ grad, int ldg, double *__restrict__ acc)
{
    double loss = 0.0, cnt = 0.0;
    for (long r = blockIdx.x * 256L + threadIdx.x; r < rows; r += (long)gridDim.x * 256) {
        const float *x = logits + r * ld;
        const long long t = target[r];
        float mx = -CUDART_INF_F;
        for (int c = 0; c < nc; c++) mx = fmaxf(mx, __ldg(x + c));
        float s = 0.f;
        for (int c = 0; c < nc; c++) s += expf(__ldg(x + c) - mx);
        const float lse = mx + logf(s);
        const bool valid = t != ignore_index;
        if (valid) { loss += (double)(lse - __ldg(x + t)); cnt += 1.0; }
        float *g = grad + r * ldg;
        for (int c = 0; c < nc; c++) g[c] = valid ? (expf(__ldg(x + c) - lse) - (c == t ? 1.f : 0.f)) : 0.f;
    }
    // block reduction (fp64), one atomic pair per block
    __shared__ double sl[8], sc[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { loss += __shfl_xor_sync(0xffffffffu, loss, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5] = loss; sc[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; w++) { a += sl[w]; b += sc[w]; }
        atomicAdd(acc, a);
        atomicAdd(acc + 1, b);
    }
}

// loss = acc[0] / acc[1];  grad *= upstream / acc[1]
__global__ void cross_entropy_finish_kernel(const double *__restrict__ acc, float *__restrict__ loss)
{
    *loss = (float)(acc[0] / acc[1]);
}
__global__ void __launch_bounds__(256) cross_entropy_scale_kernel(long n, float *__restrict__ grad, const double *__restrict__ acc,
                                                                  const float *__restrict__ upstream)
{
    const float k = (float)((double)__ldg(upstream) / acc[1]);
    for (long i = blockIdx.x * 256L + threadIdx.x; i < n; i += (long)gridDim.x * 256) grad[i] *= k;
}
}  // namespace

// logits [rows, ld >= nc]; target int64 [rows]; grad [rows, ldg] receives softmax - onehot (unscaled); acc fp64[2] pre-zeroed
// receives (sum of row losses, number of non-ignored rows); loss[0] = mean over the non-ignored rows (torch's 'mean' reduction)
RSB_EXPORT int rsb_cross_entropy_forward(long rows, int num_class, const float *logits, int ld, const long long *target,
                                         long long ignore_index, float *grad, int ldg, double *acc, float *loss, cudaStream_t stream)
{
    RSB_REQUIRE(rows >= 1 && num_class >= 1 && ld >= num_class && ldg >= num_class, "bad sizes");
    cross_entropy_kernel<<<sc_grid(rows), 256, 0, stream>>>(rows, num_class, logits, ld, target, ignore_index, grad, ldg, acc);
    RSB_CHECK_LAUNCH("cross_entropy_kernel");
    cross_entropy_finish_kernel<<<1, 1, 0, stream>>>(acc, loss);
    RSB_CHECK_LAUNCH("cross_entropy_finish_kernel");
    RSB_COUNT_LAUNCH(2);
    return 0;
}

// grad (as left by the forward) *= upstream[0] / acc[1]
RSB_EXPORT int rsb_cross_entropy_backward(long n_elements, float *grad, const double *acc, const float *upstream, cudaStream_t stream)
{
    if (n_elements <= 0) return 0;
    cross_entropy_scale_kernel<<<sc_grid(n_elements), 256, 0, stream>>>(n_elements, grad, acc, upstream);
    RSB_CHECK_LAUNCH("cross_entropy_scale_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}
