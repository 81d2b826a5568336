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

// out[g,c] = max_s relu(sc[c]*Y[g*ns+s, c] + sh[c]); arg[g,c] = first s attaining it
__global__ void __launch_bounds__(256) pool_fwd_kernel(long G, int ns, int C, const float *__restrict__ Y, int ldy,
                                                       const float *__restrict__ sc, const float *__restrict__ sh,
                                                       float *__restrict__ out, int *__restrict__ arg)
{
    const long total = G * C;
    for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const long g = i / C;
        const int c = (int)(i - g * C);
        const float a = __ldg(sc + c), b = __ldg(sh + c);
        const float *y = Y + (size_t)g * ns * ldy + c;
        float best = -1.f;
        int bi = 0;
        for (int s = 0; s < ns; s++) {
            const float z = fmaxf(fmaf(__ldg(y + (size_t)s * ldy), a, b), 0.f);
            if (z > best) { best = z; bi = s; }
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
__global__ void bn_bwd_coef_kernel(int C, double rows, const double *__restrict__ stats, int dual,
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
        b[i] = (float)(-scale * m2 * istd);
        d[i] = (float)(scale * (m2 * mean * istd - m1));
        dgamma_over_g[i] = (float)s2;     // d(loss)/d(gamma)
        dbeta[i] = (float)s1;             // d(loss)/d(beta)
    }
}

}  // namespace

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
    bn_bwd_coef_kernel<<<RSB_DIVUP(dual ? 2 * C : C, 128), 128, 0, stream>>>(C, (double)rows, stats, dual, sc, mu, inv, a,
                                                                            b, d, dgamma, dbeta);
    RSB_CHECK_LAUNCH("bn_bwd_coef_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}
