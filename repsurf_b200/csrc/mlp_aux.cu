// mlp_aux.cu — the small per-channel / pooling kernels around the tensor-core GEMMs of the fused shared MLP:
// BatchNorm statistics -> affine coefficients (+ running-stat update), max-pool over nsample with arg-max,
// max-pool backward statistics, BatchNorm-backward coefficients.  All fp64 where sums over millions of rows
// are finalised.  Replaces torch's batch_norm / max / their autograd on the RepSurf path
// (classification/modules/repsurface_utils.py:236-244, segmentation/modules/repsurface_utils.py:220-228).
#include "common.cuh"

namespace {

// stats = (sum y, sum y^2) over `rows` rows -> sc = gamma/sqrt(var+eps), sh = beta - mean*sc, mu, inv;
// running_mean/var updated like torch (biased var for normalisation, unbiased for the running estimate).
__global__ void bn_finalize_kernel(int C, double rows, const double *__restrict__ stats, const float *__restrict__ gamma,
                                   const float *__restrict__ beta, float eps, float momentum, float *running_mean,
                                   float *running_var, float *__restrict__ sc, float *__restrict__ sh,
                                   float *__restrict__ mu, float *__restrict__ inv)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
        const double mean = stats[c] / rows;
        double var = stats[C + c] / rows - mean * mean;
        if (var < 0) var = 0;
        const double istd = 1.0 / sqrt(var + (double)eps);
        const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
        const float scale = (float)((double)g * istd);
        sc[c] = scale;
        sh[c] = (float)((double)b - mean * (double)g * istd);
        mu[c] = (float)mean;
        inv[c] = (float)istd;
        if (running_mean) {
            const double unbiased = rows > 1 ? var * rows / (rows - 1) : var;
            running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
            running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
        }
    }
}

// running_mean / running_var / num_batches_tracked of one BatchNorm from its batch sums (torch semantics: unbiased
// variance for the running value) - one launch instead of the dozen tiny tensor ops of the Python formulation
__global__ void bn_running_kernel(int C, double rows, const double *__restrict__ sum, const double *__restrict__ sumsq,
                                  float momentum, float *running_mean, float *running_var, long long *nbt)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
        const double mean = sum[c] / rows;
        double var = sumsq[c] / rows - mean * mean;
        if (var < 0) var = 0;
        const double unbiased = rows > 1 ? var * rows / (rows - 1) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
    if (nbt && blockIdx.x == 0 && threadIdx.x == 0) *nbt += 1;
}

// out[g,c] = max_s relu(sc[c]*Y[g*ns+s, c] + sh[c]); arg[g,c] = first s attaining it
__global__ void __launch_bounds__(256) pool_fwd_kernel(long G, int ns, int C, const float *__restrict__ Y, int ldy,
                                                       const float *__restrict__ sc, const float *__restrict__ sh,
                                                       float *__restrict__ out, int *__restrict__ arg)
{
    const long total = G * C;
    for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const long g = rsb_div(i, C);
        const int c = (int)(i - g * C);
        const float a = __ldg(sc + c), b = __ldg(sh + c);
        const float *y = Y + (size_t)g * ns * ldy + c;
        float best = -1.f;
        int bi = 0;
        for (int s = 0; s < ns; s++) {
            const float z = rsb_relu(fmaf(__ldg(y + (size_t)s * ldy), a, b));
            if (z > best || (z != z && best == best)) { best = z; bi = s; }      // the first NaN wins and stays (torch.max propagates it)
        }
        out[i] = best;
        arg[i] = bi;
    }
}

// dm = dOut where the pooled value was > 0 else 0;  stats[c] += sum_g dm, stats[C+c] += sum_g dm * xhat(arg row)
__global__ void __launch_bounds__(256) pool_bwd_stats_kernel(long G, int ns, int C, const float *__restrict__ dOut,
                                                             const int *__restrict__ arg, const float *__restrict__ Y,
                                                             int ldy, const float *__restrict__ sc,
                                                             const float *__restrict__ sh, const float *__restrict__ mu,
                                                             const float *__restrict__ inv, float *__restrict__ dm,
                                                             double *__restrict__ stats, int g_per_block)
{
    const long g0 = (long)blockIdx.x * g_per_block;
    const long g1 = min(G, g0 + g_per_block);
    for (int c = threadIdx.x; c < C; c += 256) {
        const float a = __ldg(sc + c), b = __ldg(sh + c), m = __ldg(mu + c), is = __ldg(inv + c);
        double s1 = 0, s2 = 0;
        for (long g = g0; g < g1; g++) {
            const int s = arg[g * C + c];
            const float y = __ldg(Y + ((size_t)g * ns + s) * ldy + c);
            const float z = fmaf(y, a, b);
            const float d = z > 0.f ? dOut[g * C + c] : 0.f;
            dm[g * C + c] = d;
            s1 += d;
            s2 += (double)d * (double)((y - m) * is);
        }
        atomicAdd(stats + c, s1);
        atomicAdd(stats + C + c, s2);
    }
}

// BatchNorm backward as an affine map of stored tensors: dY = a*dZ + b*Y + d, with
//   m1 = mean(dZ), m2 = mean(dZ*xhat):  a = sc, b = -sc*m2*inv, d = sc*(m2*mu*inv - m1);  dgamma = sum dZ*xhat, dbeta = sum dZ.
// dual: two BatchNorms share dZ (stats = [sum dZ | sum dZ*xhat_1 | sum dZ*xhat_2]); outputs have 2C entries.
// sc = gamma / sqrt(running_var + eps), sh = beta - running_mean * sc, mu = running_mean, inv = 1 / sqrt(running_var + eps):
// the coefficient vectors of an eval-mode (or frozen) BatchNorm, in the form the GEMM operand transforms take
__global__ void bn_eval_coef_kernel(int C, const float *__restrict__ gamma, const float *__restrict__ beta,
                                    const float *__restrict__ rm, const float *__restrict__ rv, float eps,
                                    float *__restrict__ sc, float *__restrict__ sh, float *__restrict__ mu, float *__restrict__ inv)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double istd = 1.0 / sqrt((double)rv[c] + (double)eps);
    const double g = gamma ? (double)gamma[c] : 1.0, b = beta ? (double)beta[c] : 0.0;
    sc[c] = (float)(g * istd);
    sh[c] = (float)(b - (double)rm[c] * g * istd);
    mu[c] = rm[c];
    inv[c] = (float)istd;
}

__global__ void bn_bwd_coef_kernel(int C, double rows, const double *__restrict__ stats, int dual, int frozen,
                                   const float *__restrict__ sc, const float *__restrict__ mu,
                                   const float *__restrict__ inv, float *__restrict__ a, float *__restrict__ b,
                                   float *__restrict__ d, float *__restrict__ dgamma_over_g, float *__restrict__ dbeta)
{
    const int total = dual ? 2 * C : C;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % C, h = i / C;
        const double s1 = stats[c], s2 = stats[(1 + h) * C + c];
        const double m1 = s1 / rows, m2 = s2 / rows;
        const double scale = sc[i], mean = mu[i], istd = inv[i];
        a[i] = (float)scale;
        // frozen BatchNorm (running statistics): the statistics are constants, dY = scale * dZ
        b[i] = frozen ? 0.f : (float)(-scale * m2 * istd);
        d[i] = frozen ? 0.f : (float)(scale * (m2 * mean * istd - m1));
        dgamma_over_g[i] = (float)s2;     // d(loss)/d(gamma)
        dbeta[i] = (float)s1;             // d(loss)/d(beta)
    }
}

// dZ = dA where the layer's activation was positive (z = sc*Y + sh [+ second BatchNorm half] > 0), in place;
// stats += (sum dZ, sum dZ*xhat_1 [, sum dZ*xhat_2]) per channel.  Thread = one channel quad, grid-stride over
// rows: coalesced 16-byte accesses, register accumulation, one shared-memory reduction + fp64 atomics per block.
__global__ void __launch_bounds__(256) bn_relu_bwd_kernel(long rows, int C, const float *src, int lds, float *dA, int ldd,
                                                          const float *__restrict__ Y, int ldy,
                                                          const float *__restrict__ sc, const float *__restrict__ sh,
                                                          const float *__restrict__ mu, const float *__restrict__ inv,
                                                          int dual, double *__restrict__ stats, int rows_per_block)
{
    const int quads = C / 4;                         // C % 4 == 0 (checked by the launcher)
    const int q = threadIdx.x % quads;               // blockDim.x is a multiple of quads
    const int rlane = threadIdx.x / quads, rstep = blockDim.x / quads;
    const int c = q * 4;
    const float4 a1 = *reinterpret_cast<const float4 *>(sc + c), b1 = *reinterpret_cast<const float4 *>(sh + c);
    const float4 m1 = *reinterpret_cast<const float4 *>(mu + c), i1 = *reinterpret_cast<const float4 *>(inv + c);
    float4 a2 = make_float4(0, 0, 0, 0), b2 = a2, m2 = a2, i2 = a2;
    if (dual) {
        a2 = *reinterpret_cast<const float4 *>(sc + C + c); b2 = *reinterpret_cast<const float4 *>(sh + C + c);
        m2 = *reinterpret_cast<const float4 *>(mu + C + c); i2 = *reinterpret_cast<const float4 *>(inv + C + c);
    }
    float s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    const long r0 = (long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (long r = r0 + rlane; r < r1; r += rstep) {
        float4 d = *reinterpret_cast<const float4 *>(src + (size_t)r * lds + c);    // src == dA: in place
        const float4 y = __ldg(reinterpret_cast<const float4 *>(Y + (size_t)r * ldy + c));
        float4 y2 = make_float4(0, 0, 0, 0);
        float z[4] = {fmaf(y.x, a1.x, b1.x), fmaf(y.y, a1.y, b1.y), fmaf(y.z, a1.z, b1.z), fmaf(y.w, a1.w, b1.w)};
        if (dual) {
            y2 = __ldg(reinterpret_cast<const float4 *>(Y + (size_t)r * ldy + C + c));
            z[0] += fmaf(y2.x, a2.x, b2.x); z[1] += fmaf(y2.y, a2.y, b2.y); z[2] += fmaf(y2.z, a2.z, b2.z); z[3] += fmaf(y2.w, a2.w, b2.w);
        }
        d.x = z[0] > 0.f ? d.x : 0.f; d.y = z[1] > 0.f ? d.y : 0.f; d.z = z[2] > 0.f ? d.z : 0.f; d.w = z[3] > 0.f ? d.w : 0.f;
        *reinterpret_cast<float4 *>(dA + (size_t)r * ldd + c) = d;
        s0[0] += d.x; s0[1] += d.y; s0[2] += d.z; s0[3] += d.w;
        s1[0] = fmaf(d.x, (y.x - m1.x) * i1.x, s1[0]); s1[1] = fmaf(d.y, (y.y - m1.y) * i1.y, s1[1]);
        s1[2] = fmaf(d.z, (y.z - m1.z) * i1.z, s1[2]); s1[3] = fmaf(d.w, (y.w - m1.w) * i1.w, s1[3]);
        if (dual) {
            s2[0] = fmaf(d.x, (y2.x - m2.x) * i2.x, s2[0]); s2[1] = fmaf(d.y, (y2.y - m2.y) * i2.y, s2[1]);
            s2[2] = fmaf(d.z, (y2.z - m2.z) * i2.z, s2[2]); s2[3] = fmaf(d.w, (y2.w - m2.w) * i2.w, s2[3]);
        }
    }
    // block reduction over the row lanes that share a channel quad
    __shared__ float red[3][256][4];
#pragma unroll
    for (int e = 0; e < 4; e++) { red[0][threadIdx.x][e] = s0[e]; red[1][threadIdx.x][e] = s1[e]; red[2][threadIdx.x][e] = s2[e]; }
    __syncthreads();
    if (rlane == 0) {
        double t0[4] = {0, 0, 0, 0}, t1[4] = {0, 0, 0, 0}, t2[4] = {0, 0, 0, 0};
        for (int j = 0; j < rstep; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) { t0[e] += red[0][j * quads + q][e]; t1[e] += red[1][j * quads + q][e]; t2[e] += red[2][j * quads + q][e]; }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            atomicAdd(stats + c + e, t0[e]);
            atomicAdd(stats + C + c + e, t1[e]);
            if (dual) atomicAdd(stats + 2 * C + c + e, t2[e]);
        }
    }
}

// out = [relu](sc*Y + sh): materialises a BatchNorm(+ReLU) output where a consumer outside the fused chain needs it
__global__ void __launch_bounds__(256) bn_apply_kernel(long rows, int C, const float *__restrict__ Y, int ldy,
                                                       const float *__restrict__ sc, const float *__restrict__ sh,
                                                       int relu, float *__restrict__ out, int ldo)
{
    const int quads = C / 4;
    const long total = rows * quads;
    for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const long r = rsb_div(i, quads);
        const int c = (int)(i - r * quads) * 4;
        const float4 y = __ldg(reinterpret_cast<const float4 *>(Y + (size_t)r * ldy + c));
        const float4 a = __ldg(reinterpret_cast<const float4 *>(sc + c)), b = __ldg(reinterpret_cast<const float4 *>(sh + c));
        float4 o = make_float4(fmaf(y.x, a.x, b.x), fmaf(y.y, a.y, b.y), fmaf(y.z, a.z, b.z), fmaf(y.w, a.w, b.w));
        if (relu) { o.x = rsb_relu(o.x); o.y = rsb_relu(o.y); o.z = rsb_relu(o.z); o.w = rsb_relu(o.w); }
        *reinterpret_cast<float4 *>(out + (size_t)r * ldo + c) = o;
    }
}

// Max-pool backward + BatchNorm backward of the LAST shared-MLP layer, densified in place:
//   Y[r,c] := a[c] * (arg[g,c] == r - g*ns ? dm[g,c] : 0) + b[c] * Y[r,c] + d[c]      (g = r / ns)
// i.e. Y (stored pre-BN output, no longer needed) becomes dL/dY.  One streaming pass; the tensor-core wgrad and
// dgrad then read a plain matrix instead of re-evaluating the selection per element.
__global__ void __launch_bounds__(256) pool_bn_bwd_dense_kernel(long G, int ns, int C, const float *__restrict__ dm,
                                                                const int *__restrict__ arg, float *__restrict__ Y, int ldy,
                                                                const float *__restrict__ a, const float *__restrict__ b,
                                                                const float *__restrict__ d)
{
    const int quads = C / 4;
    const long total = G * quads;
    for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const long g = rsb_div(i, quads);
        const int c = (int)(i - g * quads) * 4;
        const float4 a4 = __ldg(reinterpret_cast<const float4 *>(a + c)), b4 = __ldg(reinterpret_cast<const float4 *>(b + c));
        const float4 d4 = __ldg(reinterpret_cast<const float4 *>(d + c));
        const float4 m4 = __ldg(reinterpret_cast<const float4 *>(dm + g * C + c));
        const int4 s4 = __ldg(reinterpret_cast<const int4 *>(arg + g * C + c));
        float *y = Y + (size_t)g * ns * ldy + c;
        for (int s = 0; s < ns; s++) {
            float4 v = *reinterpret_cast<float4 *>(y + (size_t)s * ldy);
            v.x = fmaf(a4.x, s4.x == s ? m4.x : 0.f, fmaf(b4.x, v.x, d4.x));
            v.y = fmaf(a4.y, s4.y == s ? m4.y : 0.f, fmaf(b4.y, v.y, d4.y));
            v.z = fmaf(a4.z, s4.z == s ? m4.z : 0.f, fmaf(b4.z, v.z, d4.z));
            v.w = fmaf(a4.w, s4.w == s ? m4.w : 0.f, fmaf(b4.w, v.w, d4.w));
            *reinterpret_cast<float4 *>(y + (size_t)s * ldy) = v;
        }
    }
}

}  // namespace

RSB_EXPORT int rsb_pool_bn_backward_dense(long G, int ns, int C, const float *dm, const int *arg, float *Y, int ldy,
                                          const float *a, const float *b, const float *d, cudaStream_t stream)
{
    RSB_REQUIRE(C >= 4 && C % 4 == 0 && ldy % 4 == 0, "channels / pitch must be multiples of 4");
    if (G == 0) return 0;
    long blocks = (G * (C / 4) + 255) / 256;
    if (blocks > (long)rsb_sm_count() * 16) blocks = (long)rsb_sm_count() * 16;
    pool_bn_bwd_dense_kernel<<<(int)blocks, 256, 0, stream>>>(G, ns, C, dm, arg, Y, ldy, a, b, d);
    RSB_CHECK_LAUNCH("pool_bn_bwd_dense_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_bn_apply(long rows, int C, const float *Y, int ldy, const float *sc, const float *sh, int relu,
                            float *out, int ldo, cudaStream_t stream)
{
    RSB_REQUIRE(C >= 4 && C % 4 == 0 && ldy % 4 == 0 && ldo % 4 == 0, "channels / pitches must be multiples of 4");
    if (rows == 0) return 0;
    long blocks = (rows * (C / 4) + 255) / 256;
    if (blocks > (long)rsb_sm_count() * 16) blocks = (long)rsb_sm_count() * 16;
    bn_apply_kernel<<<(int)blocks, 256, 0, stream>>>(rows, C, Y, ldy, sc, sh, relu, out, ldo);
    RSB_CHECK_LAUNCH("bn_apply_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_bn_relu_backward(long rows, int C, const float *src, int lds, float *dA, int ldd, const float *Y, int ldy,
                                    const float *sc, const float *sh, const float *mu, const float *inv, int dual, double *stats,
                                    cudaStream_t stream)
{
    if (!src) { src = dA; lds = ldd; }
    RSB_REQUIRE(lds % 4 == 0 && ((uintptr_t)src % 16 == 0), "source gradient must be 16-byte aligned with a pitch multiple of 4");
    RSB_REQUIRE(C >= 4 && C % 4 == 0 && C <= 1024 && ldd % 4 == 0 && ldy % 4 == 0, "channels / pitches must be multiples of 4");
    RSB_REQUIRE(((uintptr_t)dA % 16 == 0) && ((uintptr_t)Y % 16 == 0), "tensors must be 16-byte aligned");
    if (rows == 0) return 0;
    const int quads = C / 4;
    int threads = 256;
    if (quads > 256) threads = quads;                        // C up to 1024: one row lane
    else threads = (256 / quads) * quads;
    const int target_blocks = rsb_sm_count() * 8;
    long rpb = (rows + target_blocks - 1) / target_blocks;
    const long min_rpb = (long)(threads / quads) * 8;
    if (rpb < min_rpb) rpb = min_rpb;
    const int blocks = (int)((rows + rpb - 1) / rpb);
    RSB_REQUIRE(threads <= 256, "too many channels");
    bn_relu_bwd_kernel<<<blocks, threads, 0, stream>>>(rows, C, src, lds, dA, ldd, Y, ldy, sc, sh, mu, inv, dual, stats, (int)rpb);
    RSB_CHECK_LAUNCH("bn_relu_bwd_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_bn_finalize(int C, long rows, const double *stats, const float *gamma, const float *beta, float eps,
                               float momentum, float *running_mean, float *running_var, float *sc, float *sh,
                               float *mu, float *inv, cudaStream_t stream)
{
    RSB_REQUIRE(C >= 1 && rows >= 1, "bad sizes");
    bn_finalize_kernel<<<RSB_DIVUP(C, 128), 128, 0, stream>>>(C, (double)rows, stats, gamma, beta, eps, momentum,
                                                             running_mean, running_var, sc, sh, mu, inv);
    RSB_CHECK_LAUNCH("bn_finalize_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_pool_forward(long G, int ns, int C, const float *Y, int ldy, const float *sc, const float *sh,
                                float *out, int *arg, cudaStream_t stream)
{
    if (G == 0) return 0;
    long blocks = (G * C + 255) / 256;
    if (blocks > (long)rsb_sm_count() * 32) blocks = (long)rsb_sm_count() * 32;
    pool_fwd_kernel<<<(int)blocks, 256, 0, stream>>>(G, ns, C, Y, ldy, sc, sh, out, arg);
    RSB_CHECK_LAUNCH("pool_fwd_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_pool_backward_stats(long G, int ns, int C, const float *dOut, const int *arg, const float *Y, int ldy,
                                       const float *sc, const float *sh, const float *mu, const float *inv, float *dm,
                                       double *stats, cudaStream_t stream)
{
    if (G == 0) return 0;
    const int target_blocks = rsb_sm_count() * 8;
    int gpb = (int)((G + target_blocks - 1) / target_blocks);
    if (gpb < 1) gpb = 1;
    const int blocks = (int)((G + gpb - 1) / gpb);
    pool_bwd_stats_kernel<<<blocks, 256, 0, stream>>>(G, ns, C, dOut, arg, Y, ldy, sc, sh, mu, inv, dm, stats, gpb);
    RSB_CHECK_LAUNCH("pool_bwd_stats_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_bn_backward_coef(int C, long rows, const double *stats, int dual, const float *sc, const float *mu,
                                    const float *inv, float *a, float *b, float *d, float *dgamma, float *dbeta,
                                    cudaStream_t stream)
{
    RSB_REQUIRE(C >= 1 && rows >= 1, "bad sizes");
    const int frozen = (dual >> 1) & 1;     // bit 1 of `dual`: running-statistics BatchNorm
    dual &= 1;
    bn_bwd_coef_kernel<<<RSB_DIVUP(dual ? 2 * C : C, 128), 128, 0, stream>>>(C, (double)rows, stats, dual, frozen, sc, mu, inv, a,
                                                                            b, d, dgamma, dbeta);
    RSB_CHECK_LAUNCH("bn_bwd_coef_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_bn_eval_coef(int C, const float *gamma, const float *beta, const float *running_mean, const float *running_var,
                                float eps, float *sc, float *sh, float *mu, float *inv, cudaStream_t stream)
{
    RSB_REQUIRE(C >= 1 && running_mean && running_var && sc && sh && mu && inv, "bad arguments");
    bn_eval_coef_kernel<<<RSB_DIVUP(C, 128), 128, 0, stream>>>(C, gamma, beta, running_mean, running_var, eps, sc, sh, mu, inv);
    RSB_CHECK_LAUNCH("bn_eval_coef_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_bn_update_running(int C, long rows, const double *sum, const double *sumsq, float momentum,
                                     float *running_mean, float *running_var, long long *num_batches_tracked,
                                     cudaStream_t stream)
{
    RSB_REQUIRE(C >= 1 && rows >= 1 && sum && sumsq && running_mean && running_var, "bad arguments");
    bn_running_kernel<<<RSB_DIVUP(C, 128), 128, 0, stream>>>(C, (double)rows, sum, sumsq, momentum, running_mean, running_var,
                                                             num_batches_tracked);
    RSB_CHECK_LAUNCH("bn_running_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}
