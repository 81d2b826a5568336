// group.cu — index gathers / scatters of the pointops API (both layouts) for sm_100a.
//
// Replaces:
//   classification/modules/pointops/src/sampling/sampling_cuda_kernel.cu:6-46        gathering fwd/bwd
//   classification/modules/pointops/src/grouping/grouping_cuda_kernel.cu:28-92       grouping fwd/bwd
//   classification/modules/pointops/src/grouping_int/grouping_int_cuda_kernel.cu:33-65
//   classification/modules/pointops/src/interpolation/interpolation_cuda_kernel.cu:90-114,181-210
//   segmentation/modules/pointops/src/grouping/grouping_cuda_kernel.cu:5-40
//   segmentation/modules/pointops/src/interpolation/interpolation_cuda_kernel.cu:5-48
//
// These are HBM-bound byte movers.  Each thread reads its index ONCE and loops over channels
// (the reference re-reads idx for every channel and launches a (m*ns, c, b) grid of 1-element
// threads); stores are coalesced along the output's innermost dimension; grids cover the whole
// tensor instead of one block per cloud (the reference's backward kernels use `b` blocks).
// Backward scatters use fp32 atomics exactly like the reference (rule R6: sums are order-dependent).
#include "common.cuh"

namespace {

constexpr int TPB = 256;

inline int grid_for(long work, int per_block)
{
    long g = (work + per_block - 1) / per_block;
    const long cap = (long)rsb_sm_count() * 16;  // grid-stride beyond 16 CTAs/SM
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// out[b, c, j] = f[b, c, idx[b, j]]   (j over `per` = m or m*ns)
template <typename T>
__global__ void __launch_bounds__(TPB) dense_gather_fwd(int b, int c, int n, long per, const T *__restrict__ f,
                                                        const int *__restrict__ idx, T *__restrict__ out)
{
    const long total = (long)b * per;
    for (long i = blockIdx.x * (long)TPB + threadIdx.x; i < total; i += (long)gridDim.x * TPB) {
        const long bi = rsb_div(i, per), j = i - bi * per;
        const int a = __ldg(idx + i);
        const T *src = f + (size_t)bi * c * n + a;
        T *dst = out + (size_t)bi * c * per + j;
        for (int l = 0; l < c; l++) dst[(size_t)l * per] = __ldg(src + (size_t)l * n);
    }
}

// grad_f[b, c, idx[b, j]] += grad_out[b, c, j]
__global__ void __launch_bounds__(TPB) dense_gather_bwd(int b, int c, int n, long per, const float *__restrict__ go,
                                                        const int *__restrict__ idx, float *__restrict__ gf)
{
    const long total = (long)b * per;
    for (long i = blockIdx.x * (long)TPB + threadIdx.x; i < total; i += (long)gridDim.x * TPB) {
        const long bi = rsb_div(i, per), j = i - bi * per;
        const int a = __ldg(idx + i);
        float *dst = gf + (size_t)bi * c * n + a;
        const float *src = go + (size_t)bi * c * per + j;
        for (int l = 0; l < c; l++) atomicAdd(dst + (size_t)l * n, __ldg(src + (size_t)l * per));
    }
}

// out[b, c, p] = w0*f[b,c,i0] + w1*f[b,c,i1] + w2*f[b,c,i2]   (same contraction as the reference's SASS)
__global__ void __launch_bounds__(TPB) dense_interp_fwd(int b, int c, int m, int n, const float *__restrict__ f,
                                                        const int *__restrict__ idx, const float *__restrict__ w,
                                                        float *__restrict__ out)
{
    const long total = (long)b * n;
    for (long i = blockIdx.x * (long)TPB + threadIdx.x; i < total; i += (long)gridDim.x * TPB) {
        const long bi = rsb_div(i, n), p = i - bi * n;
        const int i0 = idx[i * 3], i1 = idx[i * 3 + 1], i2 = idx[i * 3 + 2];
        const float w0 = w[i * 3], w1 = w[i * 3 + 1], w2 = w[i * 3 + 2];
        const float *src = f + (size_t)bi * c * m;
        float *dst = out + (size_t)bi * c * n + p;
        for (int l = 0; l < c; l++) {
            const float *row = src + (size_t)l * m;
            float t = __fmul_rn(w0, __ldg(row + i0));
            t = __fmaf_rn(w1, __ldg(row + i1), t);
            dst[(size_t)l * n] = __fmaf_rn(w2, __ldg(row + i2), t);
        }
    }
}

__global__ void __launch_bounds__(TPB) dense_interp_bwd(int b, int c, int n, int m, const float *__restrict__ go,
                                                        const int *__restrict__ idx, const float *__restrict__ w,
                                                        float *__restrict__ gf)
{
    const long total = (long)b * n;
    for (long i = blockIdx.x * (long)TPB + threadIdx.x; i < total; i += (long)gridDim.x * TPB) {
        const long bi = rsb_div(i, n), p = i - bi * n;
        const int i0 = idx[i * 3], i1 = idx[i * 3 + 1], i2 = idx[i * 3 + 2];
        const float w0 = w[i * 3], w1 = w[i * 3 + 1], w2 = w[i * 3 + 2];
        float *dst = gf + (size_t)bi * c * m;
        const float *src = go + (size_t)bi * c * n + p;
        for (int l = 0; l < c; l++) {
            const float g = __ldg(src + (size_t)l * n);
            float *row = dst + (size_t)l * m;
            atomicAdd(row + i0, g * w0);
            atomicAdd(row + i1, g * w1);
            atomicAdd(row + i2, g * w2);
        }
    }
}

// packed: out[r, :] = in[idx[r], :]  — one thread per (row, 4-channel group) when c % 4 == 0
template <int VEC>
__global__ void __launch_bounds__(TPB) packed_group_fwd(long rows, int c, const float *__restrict__ in,
                                                        const int *__restrict__ idx, float *__restrict__ out)
{
    const int cv = c / VEC;
    const long total = rows * cv;
    for (long i = blockIdx.x * (long)TPB + threadIdx.x; i < total; i += (long)gridDim.x * TPB) {
        const long r = rsb_div(i, cv);
        const int ch = (int)(i - r * cv);
        const int a = __ldg(idx + r);
        if (VEC == 4) {
            reinterpret_cast<float4 *>(out)[i] = __ldg(reinterpret_cast<const float4 *>(in + (size_t)a * c) + ch);
        } else {
            out[i] = __ldg(in + (size_t)a * c + ch);
        }
    }
}

__global__ void __launch_bounds__(TPB) packed_group_bwd(long rows, int c, const float *__restrict__ go,
                                                        const int *__restrict__ idx, float *__restrict__ gi)
{
    const long total = rows * c;
    for (long i = blockIdx.x * (long)TPB + threadIdx.x; i < total; i += (long)gridDim.x * TPB) {
        const long r = rsb_div(i, c);
        const int ch = (int)(i - r * c);
        atomicAdd(gi + (size_t)__ldg(idx + r) * c + ch, __ldg(go + i));
    }
}

// packed: out[p, ch] (+)= sum_i in[idx[p,i], ch] * w[p,i]  (accumulates onto `out` like the reference)
__global__ void __launch_bounds__(TPB) packed_interp_fwd(long n, int c, int k, const float *__restrict__ in,
                                                         const int *__restrict__ idx, const float *__restrict__ w,
                                                         float *__restrict__ out)
{
    const long total = n * c;
    for (long i = blockIdx.x * (long)TPB + threadIdx.x; i < total; i += (long)gridDim.x * TPB) {
        const long p = rsb_div(i, c);
        const int ch = (int)(i - p * c);
        float acc = out[i];
        for (int j = 0; j < k; j++)
            acc = __fmaf_rn(__ldg(in + (size_t)__ldg(idx + p * k + j) * c + ch), __ldg(w + p * k + j), acc);
        out[i] = acc;
    }
}

__global__ void __launch_bounds__(TPB) packed_interp_bwd(long n, int c, int k, const float *__restrict__ go,
                                                         const int *__restrict__ idx, const float *__restrict__ w,
                                                         float *__restrict__ gi)
{
    const long total = n * c;
    for (long i = blockIdx.x * (long)TPB + threadIdx.x; i < total; i += (long)gridDim.x * TPB) {
        const long p = rsb_div(i, c);
        const int ch = (int)(i - p * c);
        const float g = __ldg(go + i);
        for (int j = 0; j < k; j++)
            atomicAdd(gi + (size_t)__ldg(idx + p * k + j) * c + ch, g * __ldg(w + p * k + j));
    }
}

// Row-matrix builder of one grouped level (the input of the fused shared MLP), one pass:
//   rows[r, :] = [ xyz[idx[r]] - new_xyz[r / ns]  (+ its polar form when `polar`) | 0-pad to P4 |
//                  normal[idx[r], :Cn] | feature[idx[r], :Cf] | 0-pad to ld ]
// replaces the reference's three index gathers + subtraction + xyz2sphere + cat + transpose().contiguous()
// (segmentation/modules/repsurface_utils.py:36-49, classification/modules/repsurface_utils.py:37-57).
__global__ void __launch_bounds__(TPB) group_rows_fwd(long rows, int ns, int polar, int P4, int Cn, int Cf, int ld,
                                                      const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                                                      const int *__restrict__ idx, const float *__restrict__ normal,
                                                      const float *__restrict__ feature, float *__restrict__ out)
{
    const long total = rows * ld;
    for (long i = blockIdx.x * (long)TPB + threadIdx.x; i < total; i += (long)gridDim.x * TPB) {
        const long r = rsb_div(i, ld);
        const int c = (int)(i - r * ld);
        const int src = __ldg(idx + r);
        float v = 0.f;
        if (c < P4) {
            const int P = polar ? 6 : 3;
            if (c < P) {
                const float *q = new_xyz + rsb_div(r, ns) * 3;
                const float dx = __ldg(xyz + (size_t)src * 3) - __ldg(q), dy = __ldg(xyz + (size_t)src * 3 + 1) - __ldg(q + 1),
                            dz = __ldg(xyz + (size_t)src * 3 + 2) - __ldg(q + 2);
                if (c == 0) v = dx;
                else if (c == 1) v = dy;
                else if (c == 2) v = dz;
                else {
                    const float rho = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
                    if (c == 3) v = rho;
                    // torch divides by a host scalar as a multiplication by its fp32 reciprocal, in separate roundings
                    else if (c == 4) v = rho == 0.f ? 0.f : __fmul_rn(acosf(__fdiv_rn(dz, rho)), 1.0f / 3.14159265358979323846f);
                    else v = __fadd_rn(__fmul_rn(atan2f(dy, dx), 1.0f / 6.28318530717958647692f), 0.5f);
                }
            }
        } else if (c < P4 + Cn) {
            v = __ldg(normal + (size_t)src * Cn + (c - P4));
        } else if (c < P4 + Cn + Cf) {
            v = __ldg(feature + (size_t)src * Cf + (c - P4 - Cn));
        }
        out[i] = v;
    }
}

// per-point table of a level, the source of the gathered first-layer operand (RSB_OPND_GATHER):
//   out[i, :] = [ xyz[i] (3) | 0 | normal[i, :Cn] | feature[i, :Cf] | 0-pad to ld ]      (ld % 4 == 0: 16-byte rows for TMA)
__global__ void __launch_bounds__(TPB) point_table_kernel(long n, int Cn, int Cf, int ld, const float *__restrict__ xyz,
                                                          const float *__restrict__ normal, const float *__restrict__ feature,
                                                          float *__restrict__ out)
{
    const long total = n * ld;
    for (long i = blockIdx.x * (long)TPB + threadIdx.x; i < total; i += (long)gridDim.x * TPB) {
        const long r = rsb_div(i, ld);
        const int c = (int)(i - r * ld);
        float v = 0.f;
        if (c < 3) v = __ldg(xyz + r * 3 + c);
        else if (c >= 4 && c < 4 + Cn) v = __ldg(normal + r * Cn + (c - 4));
        else if (c >= 4 + Cn && c < 4 + Cn + Cf) v = __ldg(feature + r * Cf + (c - 4 - Cn));
        out[i] = v;
    }
}

// scatter-add of the feature columns of d(rows) back to the per-point tensors (fp32 atomics, like the reference's
// grouping backward: order-dependent sums, rule R6)
__global__ void __launch_bounds__(TPB) group_rows_bwd(long rows, int P4, int Cn, int Cf, int ld, const float *__restrict__ drows,
                                                      const int *__restrict__ idx, float *__restrict__ dnormal,
                                                      float *__restrict__ dfeature)
{
    const int C = Cn + Cf;
    const long total = rows * C;
    for (long i = blockIdx.x * (long)TPB + threadIdx.x; i < total; i += (long)gridDim.x * TPB) {
        const long r = rsb_div(i, C);
        const int c = (int)(i - r * C);
        const float g = __ldg(drows + (size_t)r * ld + P4 + c);
        const int src = __ldg(idx + r);
        if (c < Cn) { if (dnormal) atomicAdd(dnormal + (size_t)src * Cn + c, g); }
        else if (dfeature) atomicAdd(dfeature + (size_t)src * Cf + (c - Cn), g);
    }
}

}  // namespace

#define RSB_LAUNCH_1D(kern, work, ...)                                         \
    do {                                                                       \
        if ((work) > 0) {                                                      \
            kern<<<grid_for((work), TPB), TPB, 0, stream>>>(__VA_ARGS__);      \
            RSB_CHECK_LAUNCH(#kern);                                           \
            RSB_COUNT_LAUNCH(1);                                               \
        }                                                                      \
    } while (0)

RSB_EXPORT int rsb_gathering_forward(int b, int c, int n, int m, const float *points, const int *idx, float *out,
                                     cudaStream_t stream)
{
    RSB_LAUNCH_1D(dense_gather_fwd<float>, (long)b * m, b, c, n, (long)m, points, idx, out);
    return 0;
}

RSB_EXPORT int rsb_gathering_backward(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                      float *grad_points, cudaStream_t stream)
{
    RSB_LAUNCH_1D(dense_gather_bwd, (long)b * m, b, c, n, (long)m, grad_out, idx, grad_points);
    return 0;
}

RSB_EXPORT int rsb_grouping_forward(int b, int c, int n, int m, int nsample, const float *points, const int *idx,
                                    float *out, cudaStream_t stream)
{
    RSB_LAUNCH_1D(dense_gather_fwd<float>, (long)b * m * nsample, b, c, n, (long)m * nsample, points, idx, out);
    return 0;
}

RSB_EXPORT int rsb_grouping_backward(int b, int c, int n, int m, int nsample, const float *grad_out, const int *idx,
                                     float *grad_points, cudaStream_t stream)
{
    RSB_LAUNCH_1D(dense_gather_bwd, (long)b * m * nsample, b, c, n, (long)m * nsample, grad_out, idx, grad_points);
    return 0;
}

RSB_EXPORT int rsb_grouping_int_forward(int b, int c, int n, int m, int nsample, const long long *points,
                                        const int *idx, long long *out, cudaStream_t stream)
{
    RSB_LAUNCH_1D(dense_gather_fwd<long long>, (long)b * m * nsample, b, c, n, (long)m * nsample, points, idx, out);
    return 0;
}

RSB_EXPORT int rsb_interpolation_forward(int b, int c, int m, int n, const float *points, const int *idx,
                                         const float *weight, float *out, cudaStream_t stream)
{
    RSB_LAUNCH_1D(dense_interp_fwd, (long)b * n, b, c, m, n, points, idx, weight, out);
    return 0;
}

RSB_EXPORT int rsb_interpolation_backward(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                          const float *weight, float *grad_points, cudaStream_t stream)
{
    RSB_LAUNCH_1D(dense_interp_bwd, (long)b * n, b, c, n, m, grad_out, idx, weight, grad_points);
    return 0;
}

RSB_EXPORT int rsb_grouping_packed_forward(int m, int nsample, int c, const float *input, const int *idx,
                                           float *output, cudaStream_t stream)
{
    const long rows = (long)m * nsample;
    const bool vec = (c % 4 == 0) && (((uintptr_t)input | (uintptr_t)output) % 16 == 0);
    if (vec) RSB_LAUNCH_1D(packed_group_fwd<4>, rows * (c / 4), rows, c, input, idx, output);
    else RSB_LAUNCH_1D(packed_group_fwd<1>, rows * c, rows, c, input, idx, output);
    return 0;
}

RSB_EXPORT int rsb_grouping_packed_backward(int m, int nsample, int c, const float *grad_output, const int *idx,
                                            float *grad_input, cudaStream_t stream)
{
    const long rows = (long)m * nsample;
    RSB_LAUNCH_1D(packed_group_bwd, rows * c, rows, c, grad_output, idx, grad_input);
    return 0;
}

RSB_EXPORT int rsb_interpolation_packed_forward(int n, int c, int k, const float *input, const int *idx,
                                                const float *weight, float *output, cudaStream_t stream)
{
    RSB_LAUNCH_1D(packed_interp_fwd, (long)n * c, (long)n, c, k, input, idx, weight, output);
    return 0;
}

RSB_EXPORT int rsb_interpolation_packed_backward(int n, int c, int k, const float *grad_output, const int *idx,
                                                 const float *weight, float *grad_input, cudaStream_t stream)
{
    RSB_LAUNCH_1D(packed_interp_bwd, (long)n * c, (long)n, c, k, grad_output, idx, weight, grad_input);
    return 0;
}

RSB_EXPORT int rsb_group_rows_forward(long rows, int ns, int polar, int P4, int Cn, int Cf, int ld, const float *xyz,
                                      const float *new_xyz, const int *idx, const float *normal, const float *feature,
                                      float *out, cudaStream_t stream)
{
    RSB_REQUIRE(P4 >= (polar ? 6 : 3) && ld >= P4 + Cn + Cf, "bad column layout");
    RSB_LAUNCH_1D(group_rows_fwd, rows * ld, rows, ns, polar, P4, Cn, Cf, ld, xyz, new_xyz, idx, normal, feature, out);
    return 0;
}

RSB_EXPORT int rsb_point_table(long n, int Cn, int Cf, int ld, const float *xyz, const float *normal, const float *feature,
                               float *out, cudaStream_t stream)
{
    RSB_REQUIRE(ld % 4 == 0 && ld >= 4 + Cn + Cf && (Cn == 0 || normal) && (Cf == 0 || feature), "bad column layout");
    RSB_LAUNCH_1D(point_table_kernel, n * ld, n, Cn, Cf, ld, xyz, normal, feature, out);
    return 0;
}

RSB_EXPORT int rsb_group_rows_backward(long rows, int P4, int Cn, int Cf, int ld, const float *drows, const int *idx,
                                       float *dnormal, float *dfeature, cudaStream_t stream)
{
    RSB_LAUNCH_1D(group_rows_bwd, rows * (Cn + Cf), rows, P4, Cn, Cf, ld, drows, idx, dnormal, dfeature);
    return 0;
}

