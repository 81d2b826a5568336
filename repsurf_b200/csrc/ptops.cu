// ptops.cu — the remaining packed-layout operators of the reference's pointops package (used by the PointTransformer
// baseline that shares it): subtraction and aggregation, forward + backward.
//   replaces  segmentation/modules/pointops/src/subtraction/subtraction_cuda_kernel.cu:5-31
//             segmentation/modules/pointops/src/aggregation/aggregation_cuda_kernel.cu:5-42
// Forward results are bit-identical to the reference's (same operation order: the aggregation accumulates
// fma(input + position, weight, acc) over the samples in ascending order, as nvcc contracts the reference's
// `output += (a + b) * w`).  Backward: the reference scatters every term with fp32 atomics; here the sums that have ONE
// destination row per thread are reduced in registers in a fixed order (grad_input1 of subtraction, grad_weight of
// aggregation: deterministic, no atomics) and only the true scatters (indexed rows) use atomics — vector reductions
// (red.global.add.v4.f32) when the channel count allows.  All of it is HBM-bound gather / scatter work: 128-bit accesses
// along the channel dimension, the index of a (point, sample) pair read once per quad, full-GPU grids.
#include "common.cuh"

namespace {

constexpr int PT_TPB = 256;

inline int pt_grid(long work)
{
    long b = (work + PT_TPB - 1) / PT_TPB;
    const long cap = (long)rsb_sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// out[n, s, c] = in1[n, c] - in2[idx[n, s], c]
template <int V>
__global__ void __launch_bounds__(PT_TPB) subtraction_fwd_kernel(long n, int ns, int c, const float *__restrict__ in1,
                                                                 const float *__restrict__ in2, const int *__restrict__ idx,
                                                                 float *__restrict__ out)
{
    const int cq = c / V;
    const long total = n * ns * cq;
    for (long i = blockIdx.x * (long)PT_TPB + threadIdx.x; i < total; i += (long)gridDim.x * PT_TPB) {
        const long row = rsb_div(i, cq);                 // (point, sample) pair
        const int q = (int)(i - row * cq);
        const long p = rsb_div(row, ns);
        const long src = __ldg(idx + row);
        if (V == 4) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(in1 + p * c) + q);
            const float4 b = __ldg(reinterpret_cast<const float4 *>(in2 + src * c) + q);
            reinterpret_cast<float4 *>(out + row * c)[q] = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
        } else {
            out[row * c + q] = __ldg(in1 + p * c + q) - __ldg(in2 + src * c + q);
        }
    }
}

// grad_in1[n, c] = sum_s go[n, s, c]  (one owner per element: register reduction, fixed order);
// grad_in2[idx[n, s], c] -= go[n, s, c]  (scatter: atomics)
template <int V>
__global__ void __launch_bounds__(PT_TPB) subtraction_bwd_kernel(long n, int ns, int c, const int *__restrict__ idx,
                                                                 const float *__restrict__ go, float *__restrict__ g1,
                                                                 float *__restrict__ g2)
{
    const int cq = c / V;
    const long total = n * cq;
    for (long i = blockIdx.x * (long)PT_TPB + threadIdx.x; i < total; i += (long)gridDim.x * PT_TPB) {
        const long p = rsb_div(i, cq);
        const int q = (int)(i - p * cq);
        if (V == 4) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s = 0; s < ns; s++) {
                const float4 g = __ldg(reinterpret_cast<const float4 *>(go + (p * ns + s) * c) + q);
                acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
                float *dst = g2 + (long)__ldg(idx + p * ns + s) * c + q * 4;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(-g.x), "f"(-g.y), "f"(-g.z), "f"(-g.w) : "memory");
            }
            reinterpret_cast<float4 *>(g1 + p * c)[q] = acc;
        } else {
            float acc = 0.f;
            for (int s = 0; s < ns; s++) {
                const float g = __ldg(go + (p * ns + s) * c + q);
                acc += g;
                atomicAdd(g2 + (long)__ldg(idx + p * ns + s) * c + q, -g);
            }
            g1[p * c + q] = acc;
        }
    }
}

// out[n, c] = sum_s (in[idx[n, s], c] + pos[n, s, c]) * w[n, s, c % w_c]   (ascending s, one fma per sample)
__global__ void __launch_bounds__(PT_TPB) aggregation_fwd_kernel(long n, int ns, int c, int w_c, const float *__restrict__ in,
                                                                 const float *__restrict__ pos, const float *__restrict__ w,
                                                                 const int *__restrict__ idx, float *__restrict__ out)
{
    const long total = n * c;
    for (long i = blockIdx.x * (long)PT_TPB + threadIdx.x; i < total; i += (long)gridDim.x * PT_TPB) {
        const long p = rsb_div(i, c);
        const int ch = (int)(i - p * c);
        const int wc = ch % w_c;
        float acc = 0.f;
        for (int s = 0; s < ns; s++) {
            const long src = __ldg(idx + p * ns + s);
            const float t = __fadd_rn(__ldg(in + src * c + ch), __ldg(pos + (p * ns + s) * c + ch));
            acc = __fmaf_rn(t, __ldg(w + (p * ns + s) * w_c + wc), acc);
        }
        out[i] = acc;
    }
}

// grad_pos[n, s, c] = go[n, c] * w[n, s, c % w_c];  grad_in[idx[n, s], c] += the same (scatter: atomics)
__global__ void __launch_bounds__(PT_TPB) aggregation_bwd_in_kernel(long n, int ns, int c, int w_c, const float *__restrict__ w,
                                                                    const int *__restrict__ idx, const float *__restrict__ go,
                                                                    float *__restrict__ g_in, float *__restrict__ g_pos)
{
    const long total = n * ns * c;
    for (long i = blockIdx.x * (long)PT_TPB + threadIdx.x; i < total; i += (long)gridDim.x * PT_TPB) {
        const long row = rsb_div(i, c);
        const int ch = (int)(i - row * c);
        const long p = rsb_div(row, ns);
        const float v = __fmul_rn(__ldg(go + p * c + ch), __ldg(w + row * w_c + ch % w_c));
        g_pos[i] = v;
        atomicAdd(g_in + (long)__ldg(idx + row) * c + ch, v);
    }
}

// grad_w[n, s, wc] = sum over the channels ch with ch % w_c == wc of go[n, ch] * (in[idx[n, s], ch] + pos[n, s, ch])
// (one owner per element: register reduction in ascending channel order)
__global__ void __launch_bounds__(PT_TPB) aggregation_bwd_w_kernel(long n, int ns, int c, int w_c, const float *__restrict__ in,
                                                                   const float *__restrict__ pos, const int *__restrict__ idx,
                                                                   const float *__restrict__ go, float *__restrict__ g_w)
{
    const long total = n * ns * w_c;
    for (long i = blockIdx.x * (long)PT_TPB + threadIdx.x; i < total; i += (long)gridDim.x * PT_TPB) {
        const long row = rsb_div(i, w_c);
        const int wc = (int)(i - row * w_c);
        const long p = rsb_div(row, ns);
        const long src = __ldg(idx + row);
        float acc = 0.f;
        for (int ch = wc; ch < c; ch += w_c)
            acc = __fmaf_rn(__ldg(go + p * c + ch), __fadd_rn(__ldg(in + src * c + ch), __ldg(pos + row * c + ch)), acc);
        g_w[i] = acc;
    }
}

inline bool al16p(const void *p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

RSB_EXPORT int rsb_subtraction_forward(int n, int nsample, int c, const float *input1, const float *input2, const int *idx,
                                       float *output, cudaStream_t stream)
{
    RSB_REQUIRE(n >= 0 && nsample >= 1 && c >= 1, "bad sizes");
    if (n == 0) return 0;
    if (c % 4 == 0 && al16p(input1) && al16p(input2) && al16p(output))
        subtraction_fwd_kernel<4><<<pt_grid((long)n * nsample * (c / 4)), PT_TPB, 0, stream>>>(n, nsample, c, input1, input2, idx, output);
    else
        subtraction_fwd_kernel<1><<<pt_grid((long)n * nsample * c), PT_TPB, 0, stream>>>(n, nsample, c, input1, input2, idx, output);
    RSB_CHECK_LAUNCH("subtraction_fwd_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

// grad_input1 is written (no pre-zeroing needed); grad_input2 is accumulated into and must be zero on entry
RSB_EXPORT int rsb_subtraction_backward(int n, int nsample, int c, const int *idx, const float *grad_output, float *grad_input1,
                                        float *grad_input2, cudaStream_t stream)
{
    RSB_REQUIRE(n >= 0 && nsample >= 1 && c >= 1, "bad sizes");
    if (n == 0) return 0;
    if (c % 4 == 0 && al16p(grad_output) && al16p(grad_input1) && al16p(grad_input2))
        subtraction_bwd_kernel<4><<<pt_grid((long)n * (c / 4)), PT_TPB, 0, stream>>>(n, nsample, c, idx, grad_output, grad_input1, grad_input2);
    else
        subtraction_bwd_kernel<1><<<pt_grid((long)n * c), PT_TPB, 0, stream>>>(n, nsample, c, idx, grad_output, grad_input1, grad_input2);
    RSB_CHECK_LAUNCH("subtraction_bwd_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_aggregation_forward(int n, int nsample, int c, int w_c, const float *input, const float *position,
                                       const float *weight, const int *idx, float *output, cudaStream_t stream)
{
    RSB_REQUIRE(n >= 0 && nsample >= 1 && c >= 1 && w_c >= 1, "bad sizes");
    if (n == 0) return 0;
    aggregation_fwd_kernel<<<pt_grid((long)n * c), PT_TPB, 0, stream>>>(n, nsample, c, w_c, input, position, weight, idx, output);
    RSB_CHECK_LAUNCH("aggregation_fwd_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

// grad_position and grad_weight are written; grad_input is accumulated into and must be zero on entry
RSB_EXPORT int rsb_aggregation_backward(int n, int nsample, int c, int w_c, const float *input, const float *position,
                                        const float *weight, const int *idx, const float *grad_output, float *grad_input,
                                        float *grad_position, float *grad_weight, cudaStream_t stream)
{
    RSB_REQUIRE(n >= 0 && nsample >= 1 && c >= 1 && w_c >= 1, "bad sizes");
    if (n == 0) return 0;
    aggregation_bwd_in_kernel<<<pt_grid((long)n * nsample * c), PT_TPB, 0, stream>>>(n, nsample, c, w_c, weight, idx, grad_output, grad_input,
                                                                                     grad_position);
    RSB_CHECK_LAUNCH("aggregation_bwd_in_kernel");
    aggregation_bwd_w_kernel<<<pt_grid((long)n * nsample * w_c), PT_TPB, 0, stream>>>(n, nsample, c, w_c, input, position, idx, grad_output,
                                                                                      grad_weight);
    RSB_CHECK_LAUNCH("aggregation_bwd_w_kernel");
    RSB_COUNT_LAUNCH(2);
    return 0;
}
