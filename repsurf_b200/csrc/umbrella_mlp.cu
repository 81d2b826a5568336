// Umbrella surface MLP of the segmentation tree, fused and recomputing (csrc of the a10 row of the scope table):
//
//     rows [n*g, 10] --Conv1d(10,10)+b--> Y1 --BatchNorm(train)--> ReLU --Conv1d(10,10)+b--> sum over the g triangles
//
// reference: segmentation/modules/repsurface_utils.py:297-302 (the nn.Sequential) and :322-327 (aggregation 'sum').
// At 10 channels a GEMM tile is almost all padding, and every intermediate of the torch formulation is a
// [n*g, 10] tensor written to and read back from HBM (linear, BN, ReLU, linear, sum: ~10 passes forward and as many
// backward).  Here nothing but the input rows and the [n, 10] output ever touches HBM: each pass recomputes the
// 10x10 layers from the rows in registers (200 FMAs per row) -
//   forward : (1) column sums / sums of squares of Y1 -> rsb_bn_finalize,   (2) out = sum_g W2 relu(bn(Y1)) + b2
//   backward: (1) dW2, db2, sum dZ, sum dZ*xhat,                            (2) dW1, db1 through the BN backward
// A thread owns one row; the g rows of a point are summed through shared memory (a block covers whole points).
// Reductions over all rows: fp32 per thread, warp shuffles, shared memory, then ONE fp64 atomic per value and block.
#include "common.cuh"

namespace {

constexpr int UC = 10;        // channels in and out (in_channel = 10: centroid 3 + polar 3 + normal 3 + position 1)
constexpr int UTPB = 256;

struct UmbW {
    float w1[UC][UC], b1[UC], w2[UC][UC], b2[UC];
    float sc[UC], sh[UC], mu[UC], inv[UC];
};

__device__ __forceinline__ void load_w(UmbW &S, const float *W1, const float *b1, const float *W2, const float *b2,
                                       const float *sc, const float *sh, const float *mu, const float *inv)
{
    for (int i = threadIdx.x; i < UC * UC; i += blockDim.x) {
        S.w1[i / UC][i % UC] = __ldg(W1 + i);
        S.w2[i / UC][i % UC] = W2 ? __ldg(W2 + i) : 0.f;
    }
    for (int i = threadIdx.x; i < UC; i += blockDim.x) {
        S.b1[i] = __ldg(b1 + i);
        S.b2[i] = b2 ? __ldg(b2 + i) : 0.f;
        S.sc[i] = sc ? __ldg(sc + i) : 0.f;
        S.sh[i] = sh ? __ldg(sh + i) : 0.f;
        S.mu[i] = mu ? __ldg(mu + i) : 0.f;
        S.inv[i] = inv ? __ldg(inv + i) : 0.f;
    }
    __syncthreads();
}

// a row is 40 bytes: 8-byte aligned, five 64-bit loads; consecutive lanes read consecutive rows
__device__ __forceinline__ void load_row(const float *__restrict__ X, long r, float (&x)[UC])
{
    const float2 *p = reinterpret_cast<const float2 *>(X + r * UC);
#pragma unroll
    for (int i = 0; i < UC / 2; i++) {
        const float2 v = __ldg(p + i);
        x[2 * i] = v.x;
        x[2 * i + 1] = v.y;
    }
}

__device__ __forceinline__ void layer1(const UmbW &S, const float (&x)[UC], float (&y)[UC])
{
#pragma unroll
    for (int c = 0; c < UC; c++) {
        float a = S.b1[c];
#pragma unroll
        for (int k = 0; k < UC; k++) a = fmaf(S.w1[c][k], x[k], a);
        y[c] = a;
    }
}

// block-wide sum of NV per-thread values into acc[0..NV) (fp64 atomics, one per value and block)
template <int NV>
__device__ __forceinline__ void block_accumulate(float (&v)[NV], double *__restrict__ acc, float *red /* [warps][NV] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        float s = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp * NV + i] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NV; i += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < nw; w++) s += (double)red[w * NV + i];
        atomicAdd(acc + i, s);
    }
}

// forward pass 1: stats[0..C) += sum_r Y1[r, c],  stats[C..2C) += sum_r Y1[r, c]^2
__global__ void __launch_bounds__(UTPB) umb_stats_kernel(long rows, const float *__restrict__ X, const float *__restrict__ W1,
                                                         const float *__restrict__ b1, double *__restrict__ stats)
{
    __shared__ UmbW S;
    __shared__ float red[(UTPB / 32) * 2 * UC];
    load_w(S, W1, b1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    float acc[2 * UC];
#pragma unroll
    for (int i = 0; i < 2 * UC; i++) acc[i] = 0.f;
    for (long r = blockIdx.x * (long)UTPB + threadIdx.x; r < rows; r += (long)gridDim.x * UTPB) {
        float x[UC], y[UC];
        load_row(X, r, x);
        layer1(S, x, y);
#pragma unroll
        for (int c = 0; c < UC; c++) { acc[c] += y[c]; acc[UC + c] = fmaf(y[c], y[c], acc[UC + c]); }
    }
    block_accumulate<2 * UC>(acc, stats, red);
}

// forward pass 2: out[p, :] = sum_{j<g} ( W2 relu(sc*Y1 + sh) + b2 )[p*g + j, :]
// a block covers (UTPB / g) whole points per iteration; the per-row results meet in shared memory
__global__ void __launch_bounds__(UTPB) umb_forward_kernel(long rows, int g, const float *__restrict__ X,
                                                           const float *__restrict__ W1, const float *__restrict__ b1,
                                                           const float *__restrict__ W2, const float *__restrict__ b2,
                                                           const float *__restrict__ sc, const float *__restrict__ sh,
                                                           float *__restrict__ out)
{
    __shared__ UmbW S;
    __shared__ float so[UTPB][UC + 1];
    load_w(S, W1, b1, W2, b2, sc, sh, nullptr, nullptr);
    const int ppb = UTPB / g, rpb = ppb * g;                  // points / rows per block iteration
    const long n_iter = (rows + rpb - 1) / rpb;
    for (long it = blockIdx.x; it < n_iter; it += gridDim.x) {
        const long base = it * rpb;
        const long r = base + threadIdx.x;
        if ((int)threadIdx.x < rpb && r < rows) {
            float x[UC], y[UC];
            load_row(X, r, x);
            layer1(S, x, y);
#pragma unroll
            for (int c = 0; c < UC; c++) y[c] = rsb_relu(fmaf(S.sc[c], y[c], S.sh[c]));
#pragma unroll
            for (int c = 0; c < UC; c++) {
                float a = S.b2[c];
#pragma unroll
                for (int k = 0; k < UC; k++) a = fmaf(S.w2[c][k], y[k], a);
                so[threadIdx.x][c] = a;
            }
        }
        __syncthreads();
        const long p0 = base / g;
        const long np = min((long)ppb, rows / g - p0);
        for (int i = threadIdx.x; i < np * UC; i += UTPB) {
            const int p = i / UC, c = i - p * UC;
            float a = 0.f;
            for (int j = 0; j < g; j++) a += so[p * g + j][c];
            out[p0 * UC + i] = a;
        }
        __syncthreads();
    }
}

// shared by both backward passes: recompute the row, return xhat, relu mask product dZ
__device__ __forceinline__ void backward_row(const UmbW &S, const float (&x)[UC], const float (&dO)[UC], float (&xh)[UC],
                                             float (&h)[UC], float (&dZ)[UC])
{
    float y[UC];
    layer1(S, x, y);
#pragma unroll
    for (int k = 0; k < UC; k++) {
        xh[k] = (y[k] - S.mu[k]) * S.inv[k];
        const float z = fmaf(S.sc[k], y[k], S.sh[k]);
        h[k] = rsb_relu(z);
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < UC; c++) d = fmaf(S.w2[c][k], dO[c], d);
        dZ[k] = z > 0.f ? d : 0.f;
    }
}

// backward pass 1: acc = [ dW2 (C x C, row = output channel) | db2 (C) | sum dZ (C) | sum dZ*xhat (C) ]
constexpr int B1_VALS = UC * UC + 3 * UC;
__global__ void __launch_bounds__(UTPB) umb_backward1_kernel(long rows, int g, const float *__restrict__ X,
                                                             const float *__restrict__ dOut, const float *__restrict__ W1,
                                                             const float *__restrict__ b1, const float *__restrict__ W2,
                                                             const float *__restrict__ sc, const float *__restrict__ sh,
                                                             const float *__restrict__ mu, const float *__restrict__ inv,
                                                             double *__restrict__ acc)
{
    __shared__ UmbW S;
    __shared__ float red[(UTPB / 32) * B1_VALS];
    load_w(S, W1, b1, W2, nullptr, sc, sh, mu, inv);
    float a[B1_VALS];
#pragma unroll
    for (int i = 0; i < B1_VALS; i++) a[i] = 0.f;
    for (long r = blockIdx.x * (long)UTPB + threadIdx.x; r < rows; r += (long)gridDim.x * UTPB) {
        float x[UC], dO[UC], xh[UC], h[UC], dZ[UC];
        load_row(X, r, x);
        load_row(dOut, rsb_div(r, g), dO);
        backward_row(S, x, dO, xh, h, dZ);
#pragma unroll
        for (int c = 0; c < UC; c++) {
#pragma unroll
            for (int k = 0; k < UC; k++) a[c * UC + k] = fmaf(dO[c], h[k], a[c * UC + k]);
            a[UC * UC + c] += dO[c];
            a[UC * UC + UC + c] += dZ[c];
            a[UC * UC + 2 * UC + c] = fmaf(dZ[c], xh[c], a[UC * UC + 2 * UC + c]);
        }
    }
    block_accumulate<B1_VALS>(a, acc, red);
}

// backward pass 2: dY1 = gamma*inv * (dZ - mean(dZ) - xhat*mean(dZ*xhat));  acc2 = [ dW1 (C x Cin) | db1 (C) ]
constexpr int B2_VALS = UC * UC + UC;
__global__ void __launch_bounds__(UTPB) umb_backward2_kernel(long rows, int g, int train, const float *__restrict__ X,
                                                             const float *__restrict__ dOut, const float *__restrict__ W1,
                                                             const float *__restrict__ b1, const float *__restrict__ W2,
                                                             const float *__restrict__ sc, const float *__restrict__ sh,
                                                             const float *__restrict__ mu, const float *__restrict__ inv,
                                                             const double *__restrict__ acc1, double *__restrict__ acc2)
{
    __shared__ UmbW S;
    __shared__ float red[(UTPB / 32) * B2_VALS];
    __shared__ float m_dz[UC], m_dzx[UC];
    if (threadIdx.x < UC) {
        // eval mode: the statistics are constants, dY1 = gamma*inv*dZ
        m_dz[threadIdx.x] = train ? (float)(acc1[UC * UC + UC + threadIdx.x] / (double)rows) : 0.f;
        m_dzx[threadIdx.x] = train ? (float)(acc1[UC * UC + 2 * UC + threadIdx.x] / (double)rows) : 0.f;
    }
    load_w(S, W1, b1, W2, nullptr, sc, sh, mu, inv);
    float a[B2_VALS];
#pragma unroll
    for (int i = 0; i < B2_VALS; i++) a[i] = 0.f;
    for (long r = blockIdx.x * (long)UTPB + threadIdx.x; r < rows; r += (long)gridDim.x * UTPB) {
        float x[UC], dO[UC], xh[UC], h[UC], dZ[UC];
        load_row(X, r, x);
        load_row(dOut, rsb_div(r, g), dO);
        backward_row(S, x, dO, xh, h, dZ);
#pragma unroll
        for (int k = 0; k < UC; k++) {
            // sc = gamma * inv
            const float dy = S.sc[k] * (dZ[k] - m_dz[k] - xh[k] * m_dzx[k]);
#pragma unroll
            for (int j = 0; j < UC; j++) a[k * UC + j] = fmaf(dy, x[j], a[k * UC + j]);
            a[UC * UC + k] += dy;
        }
    }
    block_accumulate<B2_VALS>(a, acc2, red);
}

inline int umb_grid(long rows)
{
    const long want = (rows + UTPB - 1) / UTPB;
    const long cap = (long)rsb_sm_count() * 4;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace

RSB_EXPORT int rsb_umbrella_mlp_stats(long rows, int cin, int c, const float *X, const float *W1, const float *b1,
                                      double *stats, cudaStream_t stream)
{
    RSB_REQUIRE(cin == UC && c == UC, "umbrella MLP kernels are built for 10 -> 10 -> 10 channels");
    if (rows == 0) return 0;
    umb_stats_kernel<<<umb_grid(rows), UTPB, 0, stream>>>(rows, X, W1, b1, stats);
    RSB_CHECK_LAUNCH("umb_stats_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_umbrella_mlp_forward(long rows, int g, int cin, int c, const float *X, const float *W1, const float *b1,
                                        const float *W2, const float *b2, const float *sc, const float *sh, float *out,
                                        cudaStream_t stream)
{
    RSB_REQUIRE(cin == UC && c == UC, "umbrella MLP kernels are built for 10 -> 10 -> 10 channels");
    RSB_REQUIRE(g >= 1 && g <= UTPB && rows % g == 0, "rows must hold whole groups of g <= 256 triangles");
    if (rows == 0) return 0;
    umb_forward_kernel<<<umb_grid(rows + rows / 8), UTPB, 0, stream>>>(rows, g, X, W1, b1, W2, b2, sc, sh, out);
    RSB_CHECK_LAUNCH("umb_forward_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_umbrella_mlp_backward(long rows, int g, int cin, int c, int train, const float *X, const float *dOut, const float *W1,
                                         const float *b1, const float *W2, const float *sc, const float *sh, const float *mu,
                                         const float *inv, double *acc1, double *acc2, cudaStream_t stream)
{
    RSB_REQUIRE(cin == UC && c == UC, "umbrella MLP kernels are built for 10 -> 10 -> 10 channels");
    RSB_REQUIRE(g >= 1 && rows % g == 0, "bad group size");
    if (rows == 0) return 0;
    umb_backward1_kernel<<<umb_grid(rows), UTPB, 0, stream>>>(rows, g, X, dOut, W1, b1, W2, sc, sh, mu, inv, acc1);
    RSB_CHECK_LAUNCH("umb_backward1_kernel");
    umb_backward2_kernel<<<umb_grid(rows), UTPB, 0, stream>>>(rows, g, train, X, dOut, W1, b1, W2, sc, sh, mu, inv, acc1, acc2);
    RSB_CHECK_LAUNCH("umb_backward2_kernel");
    RSB_COUNT_LAUNCH(2);
    return 0;
}
