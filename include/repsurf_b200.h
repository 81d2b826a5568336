/*
 * repsurf_b200.h — C-ABI of librepsurf_b200.so (sm_100a kernels for the RepSurf-U hot path).
 *
 * This is the drop-in boundary.  Each entry replaces one `extern "C" *_cuda_launcher` of the
 * reference's pointops extension (declared in its *_cuda_kernel.h files) with the same argument
 * order and meaning, plus:
 *   - an explicit cudaStream_t on EVERY entry (the reference launches most kernels on the legacy
 *     default stream; only ballquery_fast / knnquery / knnquery_heap take a stream), and
 *   - an int status return: 0 on success, otherwise a cudaError_t value with a message available
 *     from rsb_last_error() (the reference's `_fast` launchers print and exit(-1), the others do
 *     not check at all).
 * All pointers are DEVICE pointers to contiguous fp32 / int32 arrays on the current device.
 * Ownership is the reference's: the caller allocates every output and scratch buffer; kernels
 * write in place; nothing is retained after return; all work is asynchronous on `stream`.
 * Thread-safety: entries are re-entrant; rsb_last_error() is per host thread.
 *
 * Paths below are relative to the reference repository root; cls/po = classification/modules/pointops,
 * seg/po = segmentation/modules/pointops.
 */
#ifndef REPSURF_B200_H
#define REPSURF_B200_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

int rsb_abi_version(void);
const char *rsb_last_error(void);
unsigned long long rsb_launch_count(void);   /* kernels launched by this library since the last reset */
void rsb_reset_launch_count(void);

/* ------------------------------------------------------------------ dense layout (classification) */

/* FPS kernel generation for multi-CTA (cluster) plans: 0 (default) = per-sample exchange by st.async messages completing
 * on the destination's mbarrier; 1 = one barrier.cluster per sample (round-1 kernel).  Results are identical. */
void rsb_fps_set_generation(int gen);

/* replaces furthestsampling_cuda_launcher(b,n,m,dataset,temp,idxs)      cls/po/src/sampling/sampling_cuda_kernel.h:19
 * xyz [b,n,3]; temp [b,n] scratch or NULL (no need to pre-fill; final running minima are written back
 * when given); idx [b,m] int32, idx[:,0] = 0.  new_xyz: optional [b,m,3] fused gather of the sampled
 * coordinates (NULL to skip) — replaces the separate gathering call of cls/modules/repsurface_utils.py:30. */
int rsb_furthestsampling_dense(int b, int n, int m, const float *xyz, float *temp, int *idx, float *new_xyz,
                               cudaStream_t stream);

/* replaces gathering_forward_cuda_launcher / gathering_backward_cuda_launcher   cls/po/src/sampling/sampling_cuda_kernel.h:17-18
 * points [b,c,n], idx [b,m] -> out [b,c,m]; backward accumulates (atomicAdd) into a caller-zeroed grad_points [b,c,n]. */
int rsb_gathering_forward(int b, int c, int n, int m, const float *points, const int *idx, float *out,
                          cudaStream_t stream);
int rsb_gathering_backward(int b, int c, int n, int m, const float *grad_out, const int *idx, float *grad_points,
                           cudaStream_t stream);

/* replaces ballquery_cuda_launcher_fast(b,n,m,radius,nsample,new_xyz,xyz,idx,stream)   cls/po/src/ballquery/ballquery_cuda_kernel.h:17
 * idx [b,m,nsample]: first nsample indices (ascending) with d2 < radius^2, padded with the first hit, zeros if none.
 * Unlike the reference the output need not be pre-zeroed. */
int rsb_ballquery(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx,
                  cudaStream_t stream);

/* replaces knnquery_cuda_launcher(b,n,m,nsample,xyz,new_xyz,idx,dist2,stream)          cls/po/src/knnquery/knnquery_cuda_kernel.h:15
 * idx [b,m,nsample] sorted by (d2, index); dist2 [b,m,nsample] or NULL.  1 <= nsample <= 200. */
int rsb_knnquery_dense(int b, int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx,
                       float *dist2, cudaStream_t stream);

/* replaces knnquery_heap_cuda_launcher(...)                                            cls/po/src/knnquery_heap/knnquery_heap_cuda_kernel.h:15
 * same shapes; heap order semantics (ascending d2, reference tie order).  1 <= nsample <= 100. */
int rsb_knnquery_heap_dense(int b, int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx,
                            float *dist2, cudaStream_t stream);

/* replaces grouping_forward_cuda_launcher_fast / grouping_backward_cuda_launcher       cls/po/src/grouping/grouping_cuda_kernel.h:13-17
 * points [b,c,n], idx [b,m,nsample] -> out [b,c,m,nsample]. */
int rsb_grouping_forward(int b, int c, int n, int m, int nsample, const float *points, const int *idx, float *out,
                         cudaStream_t stream);
int rsb_grouping_backward(int b, int c, int n, int m, int nsample, const float *grad_out, const int *idx,
                          float *grad_points, cudaStream_t stream);

/* replaces grouping_int_forward_cuda_launcher_fast                                     cls/po/src/grouping_int/grouping_int_cuda_kernel.h:13 */
int rsb_grouping_int_forward(int b, int c, int n, int m, int nsample, const long long *points, const int *idx,
                             long long *out, cudaStream_t stream);

/* replaces nearestneighbor_cuda_launcher_fast(b,n,m,unknown,known,dist2,idx)           cls/po/src/interpolation/interpolation_cuda_kernel.h:19
 * unknown [b,n,3], known [b,m,3] -> dist2 [b,n,3] (squared), idx [b,n,3]. */
int rsb_nearestneighbor(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                        cudaStream_t stream);

/* replaces interpolation_forward_cuda_launcher_fast(b,c,m,n,...) / interpolation_backward_cuda_launcher(b,n,c,m,...)
 *                                                                                      cls/po/src/interpolation/interpolation_cuda_kernel.h:20-22
 * points [b,c,m], idx/weight [b,n,3] -> out [b,c,n]. */
int rsb_interpolation_forward(int b, int c, int m, int n, const float *points, const int *idx, const float *weight,
                              float *out, cudaStream_t stream);
int rsb_interpolation_backward(int b, int c, int n, int m, const float *grad_out, const int *idx,
                               const float *weight, float *grad_points, cudaStream_t stream);

/* ------------------------------------------------------------------ packed layout (segmentation) */

/* replaces furthestsampling_cuda_launcher(b,n,xyz,offset,new_offset,tmp,idx)           seg/po/src/sampling/sampling_cuda_kernel.h:11
 * xyz [sum n,3]; offset/new_offset [b] int32 cumulative ends (device); n_max = largest segment length
 * (host value, as in the reference: seg/po/functions/pointops.py:39-41); idx [sum m] GLOBAL row ids.
 * n_max_dev: optional device scalar with the exact largest segment length; when given, n_max only has to be an
 * upper bound (it sizes the launch) and the reference's tie rule is derived from *n_max_dev on the device, so
 * callers whose segment sizes are produced on the GPU (sectorized FPS) need no host synchronisation. */
int rsb_furthestsampling_packed(int b, int n_max, const int *n_max_dev, const float *xyz, const int *offset,
                                const int *new_offset, float *tmp, int *idx, float *new_xyz, cudaStream_t stream);
/* Same, when the host knows only bounds on the segment sizes (the sectorized FPS of seg/po/functions/pointops.py:52-111
 * computes its sector sizes on the device): segments of up to ~n_expect points take a launch planned for n_expect, the
 * rest a second launch planned for n_limit (>= every segment).  No host synchronisation; results identical to
 * rsb_furthestsampling_packed.  n_max_dev as above (may be NULL when every launch has a segment of >= 1024 points). */
int rsb_furthestsampling_packed_bounded(int b, int n_expect, int n_limit, const int *n_max_dev, const float *xyz,
                                        const int *offset, const int *new_offset, float *tmp, int *idx, float *new_xyz,
                                        cudaStream_t stream);

/* The sector split of sectorized_fps (seg/po/functions/pointops.py:61-93: per cloud angle = atan2(x, y), num_sectors
 * equal-width sectors over [min, max + 1e-4], points of a sector in ascending index order) on the device, no host round trip.
 * nsec [b] = sectors of each cloud (1 for clouds below min_points, else num_sectors), seg_first [b] = index of each cloud's
 * first segment, nseg = sum(nsec) <= 512.  Outputs: order [n] = the stable segment-major permutation, sector_xyz [n,3] =
 * xyz[order], sector_offset [nseg] = cumulative segment ends, count_max [1] = largest segment (may be NULL).
 * rsb_sector_map_back: out[i] = order[idx[i]] as int64 (pointops.py:105). */
long rsb_sector_split_workspace_bytes(int n, int b, int nseg);
int rsb_sector_split(int b, int n, int num_sectors, int nseg, const float *xyz, const int *offset, const int *nsec,
                     const int *seg_first, void *workspace, long workspace_bytes, int *order, float *sector_xyz,
                     int *sector_offset, int *count_max, cudaStream_t stream);
int rsb_sector_map_back(int m, const int *order, const int *idx, long long *out, cudaStream_t stream);

/* replaces subtraction_forward/backward_cuda_launcher(n,nsample,c,...)     seg/po/src/subtraction/subtraction_cuda_kernel.h:11-12
 * output[n,s,c] = input1[n,c] - input2[idx[n,s],c] (bit-exact).  Backward: grad_input1 is WRITTEN (fixed-order sum over the
 * samples), grad_input2 accumulated by atomics (zero it first), as the reference's Python does (pointops.py:211-214). */
int rsb_subtraction_forward(int n, int nsample, int c, const float *input1, const float *input2, const int *idx,
                            float *output, cudaStream_t stream);
int rsb_subtraction_backward(int n, int nsample, int c, const int *idx, const float *grad_output, float *grad_input1,
                             float *grad_input2, cudaStream_t stream);
/* replaces aggregation_forward/backward_cuda_launcher(n,nsample,c,w_c,...) seg/po/src/aggregation/aggregation_cuda_kernel.h:11-12
 * output[n,c] = sum_s (input[idx[n,s],c] + position[n,s,c]) * weight[n,s,c % w_c] (bit-exact: ascending s, one fma per
 * sample).  Backward: grad_position / grad_weight written, grad_input accumulated by atomics (zero it first). */
int rsb_aggregation_forward(int n, int nsample, int c, int w_c, const float *input, const float *position,
                            const float *weight, const int *idx, float *output, cudaStream_t stream);
int rsb_aggregation_backward(int n, int nsample, int c, int w_c, const float *input, const float *position,
                             const float *weight, const int *idx, const float *grad_output, float *grad_input,
                             float *grad_position, float *grad_weight, cudaStream_t stream);

/* replaces knnquery_cuda_launcher(m,nsample,xyz,new_xyz,offset,new_offset,idx,dist2)   seg/po/src/knnquery/knnquery_cuda_kernel.h:11
 * adds b (= number of clouds).  dist [m,nsample]: squared distances, or their square roots when sqrt_out != 0
 * (fuses the torch.sqrt of seg/po/functions/pointops.py:127).  1 <= nsample <= 100. */
int rsb_knnquery_packed(int b, int m, int nsample, const float *xyz, const float *new_xyz, const int *offset,
                        const int *new_offset, int *idx, float *dist, int sqrt_out, cudaStream_t stream);

/* Same results as rsb_knnquery_packed / rsb_knnquery_dense / rsb_knnquery_heap_dense (index-exact, same tie
 * semantics), computed through a uniform grid built per call (counting sort, ~8 points per cell) instead of the
 * reference's all-pairs scan.  packed != 0: segmentation layout (n_total/m_total rows, offset/new_offset, global ids);
 * packed == 0: dense [b,n,3]/[b,m,3], local ids.  heap != 0: heap-order tie semantics.  workspace: caller-allocated
 * device scratch of rsb_knn_grid_workspace_bytes(n_total, b) bytes. */
long rsb_knn_grid_workspace_bytes(int n_total, int b);
int rsb_knnquery_grid(int packed, int heap, int b, int n, int m, int n_total, int m_total, int nsample,
                      const float *xyz, const float *new_xyz, const int *offset, const int *new_offset, int *idx,
                      float *dist, int sqrt_out, void *workspace, long workspace_bytes, cudaStream_t stream);
/* measurement hook: a device buffer of 3 x uint64 that every grid search adds to (candidates evaluated, cell ranges
 * scanned, queries resolved by the exact tie replay); NULL disables (default). */
void rsb_knn_grid_set_counters(unsigned long long *dev_counters);

/* replaces grouping_forward/backward_cuda_launcher(m,nsample,c,...)                    seg/po/src/grouping/grouping_cuda_kernel.h:11-12
 * input [n,c], idx [m,nsample] -> output [m,nsample,c]. */
int rsb_grouping_packed_forward(int m, int nsample, int c, const float *input, const int *idx, float *output,
                                cudaStream_t stream);
int rsb_grouping_packed_backward(int m, int nsample, int c, const float *grad_output, const int *idx,
                                 float *grad_input, cudaStream_t stream);

/* replaces interpolation_forward/backward_cuda_launcher(n,c,k,...)                     seg/po/src/interpolation/interpolation_cuda_kernel.h:11-12
 * input [m,c], idx/weight [n,k] -> output [n,c] (accumulated onto the caller-zeroed output, like the reference). */
int rsb_interpolation_packed_forward(int n, int c, int k, const float *input, const int *idx, const float *weight,
                                     float *output, cudaStream_t stream);
int rsb_interpolation_packed_backward(int n, int c, int k, const float *grad_output, const int *idx,
                                      const float *weight, float *grad_input, cudaStream_t stream);

/* Row-matrix builder of a grouped level (input of the fused shared MLP), one pass, and its backward scatter:
 *   rows[r] = [xyz[idx[r]] - new_xyz[r/ns] (+ polar form) | pad to P4 | normal[idx[r]] (Cn) | feature[idx[r]] (Cf) | pad to ld]
 * idx [rows] GLOBAL row ids; replaces the gathers + subtraction + xyz2sphere + cat of
 * segmentation/modules/repsurface_utils.py:36-49 and classification/modules/repsurface_utils.py:37-57. */
int rsb_group_rows_forward(long rows, int ns, int polar, int P4, int Cn, int Cf, int ld, const float *xyz,
                           const float *new_xyz, const int *idx, const float *normal, const float *feature, float *out,
                           cudaStream_t stream);
int rsb_group_rows_backward(long rows, int P4, int Cn, int Cf, int ld, const float *drows, const int *idx, float *dnormal,
                            float *dfeature, cudaStream_t stream);
/* Per-point table of a grouped level: out[i] = [xyz[i] | 0 | normal[i] | feature[i] | 0-pad], row pitch ld (multiple of 4).
 * It is the U tensor of an RSB_OPND_GATHER operand: the first shared-MLP GEMM gathers its rows with TMA (gather4) and
 * subtracts the group centre while staging the tile, so the [centres x nsample, C] row matrix never exists in HBM. */
int rsb_point_table(long n, int Cn, int Cf, int ld, const float *xyz, const float *normal, const float *feature, float *out,
                    cudaStream_t stream);

/* ------------------------------------------------------------------ umbrella surface descriptors (both layouts)
 * One kernel for group_by_umbrella[_v2] + cal_normal + cal_center + xyz2sphere + cal_const + check_nan_umb
 * ({classification,segmentation}/modules/{repsurface,recons,polar}_utils.py).  xyz [rows,3]; idx [np,k] kNN lists with
 * GLOBAL row ids; flip [np] = +-1 per point (the per-cloud random inversion) or NULL; out [np, G, 10] with
 * G = k - (skip_first ? 1 : 0).  skip_first drops the query itself (classification); rotate_key sorts by the azimuth of
 * the rotated offsets (segmentation 'fix'); order_seg selects the channel order [polar,normal,pos,centroid] (else
 * [centroid,polar,normal,pos]).  out has row pitch ld >= 10 floats per triangle: the first `channels` (10, or 9 = the
 * classification tree without return_dist, cls/modules/repsurface_utils.py:291-292) descriptor channels, zeros behind. */
int rsb_umbrella_features(long np, int k, int skip_first, int rotate_key, int order_seg, const float *xyz, const int *idx,
                          const float *flip, float *out, int channels, int ld, cudaStream_t stream);

/* Umbrella MLP of the segmentation tree, Conv1d(10,10) + BatchNorm(train) + ReLU + Conv1d(10,10) + sum over the g
 * triangles of a point (segmentation/modules/repsurface_utils.py:297-302, :322-327), fused and recomputing: only the
 * rows X [rows, 10] and the output [rows/g, 10] touch HBM.  cin == c == 10 (anything else is refused).
 *   _stats   : stats[0..c) += column sums of Y1 = X W1^T + b1, stats[c..2c) += sums of squares (-> rsb_bn_finalize)
 *   _forward : out[p] = sum_{j<g} W2 relu(sc*Y1 + sh) + b2; g <= 256
 *   _backward: acc1 [c*c + 3c] += (dW2 | db2 | sum dZ = dbeta | sum dZ*xhat = dgamma), acc2 [c*cin + c] += (dW1 | db1);
 *              both zeroed by the caller, fp64.  train = 0: BatchNorm used fixed (running) statistics. */
int rsb_umbrella_mlp_stats(long rows, int cin, int c, const float *X, const float *W1, const float *b1, double *stats,
                           cudaStream_t stream);
int rsb_umbrella_mlp_forward(long rows, int g, int cin, int c, const float *X, const float *W1, const float *b1,
                             const float *W2, const float *b2, const float *sc, const float *sh, float *out,
                             cudaStream_t stream);
int rsb_umbrella_mlp_backward(long rows, int g, int cin, int c, int train, const float *X, const float *dOut, const float *W1,
                              const float *b1, const float *W2, const float *sc, const float *sh, const float *mu,
                              const float *inv, double *acc1, double *acc2, cudaStream_t stream);

/* ------------------------------------------------------------------ shared MLP on tcgen05 (both layouts)
 * Replaces the library GEMMs behind nn.Conv2d/Conv1d(1x1)/nn.Linear on the RepSurf path
 * (classification/modules/repsurface_utils.py:236-243, segmentation/modules/repsurface_utils.py:220-227,267-282)
 * with a 3xTF32 (fp32-faithful) tensor-core GEMM over rows that fuses the PREVIOUS layer's BatchNorm+ReLU into
 * the operand load and THIS layer's BatchNorm statistics into the epilogue.
 *   Y[rows,N] = act(X)[rows,K] @ W[N,K]^T + bias
 *   mode 0: act = identity; 1: relu(x*sc+sh) with sc/sh [K]; 2: relu(x[:, :K]*sc[:K]+sh[:K] + x[:, K:2K]*sc[K:]+sh[K:])
 *   stats: optional fp64 [2N] (caller-zeroed): += per-column sum and sum of squares of Y.
 * Weights are pre-split (hi/lo tf32) into the tensor-core operand layout once per step by rsb_linear_tc_prep_weight
 * into a buffer of rsb_linear_tc_weight_floats(N, K) floats; transposed != 0 takes W stored as [K, N]. */
long rsb_linear_tc_weight_floats(int N, int K);
int rsb_linear_tc_prep_weight(int N, int K, const float *W, int ldw, int transposed, float *Wp, cudaStream_t stream);
int rsb_linear_tc_forward(long rows, int K, int N, const float *X, int ldx, const float *Wp, const float *bias,
                          int mode, const float *sc, const float *sh, float *Y, double *stats, cudaStream_t stream);

/* General form used by the fused shared-MLP forward/backward.  An OPERAND is a logical [rows, K] matrix whose
 * element (r, k) is computed from stored tensors while the tile is staged (k below already includes k0): */
enum {
    RSB_OPND_RAW = 0,           /* U[r,k]                                                                     */
    RSB_OPND_BN_RELU = 1,       /* relu(a[k]*U[r,k] + d[k])                      previous BatchNorm + ReLU       */
    RSB_OPND_DUAL_BN_RELU = 2,  /* relu(a[k]*U[r,k]+d[k] + a[ku+k]*U[r,ku+k]+d[ku+k])   relu(bn_l(y_l)+bn_f(y_f)) */
    RSB_OPND_AFFINE2 = 3,       /* a[k]*U[r, k % ku] + b[k]*V[r,k] + d[k]        BatchNorm backward dY            */
    RSB_OPND_POOLED = 4,        /* a[k]*(arg[g,k]==r-g*ns ? U[g,k] : 0) + b[k]*V[r,k] + d[k],  g = r / ns        */
    RSB_OPND_GATHER = 5         /* U[arg[r], k] - (k < 3 ? V[r / ns, k] : 0): the grouped level's row matrix built while the
                                   tile is staged - U = per-point table [ku rows, ldu] = [xyz, 0 | normal, feature | 0-pad],
                                   arg = neighbour index of every row, V = group centres [rows/ns, 3]  (TMA gather4; replaces
                                   the gathers + subtraction + cat of seg/modules/repsurface_utils.py:36-49)               */
};
typedef struct {
    const float *U, *V;      /* U: [rows, ldu] (POOLED: [rows/ns, ldu] pooled gradient), V: [rows, ldv]            */
    const float *a, *b, *d;  /* per-channel coefficients                                                          */
    const int *arg;          /* POOLED: [rows/ns, ldu] sample index of the pooled maximum                         */
    int ldu, ldv;
    int K;                   /* logical width of the operand                                                      */
    int k0;                  /* channel offset added to every k                                                   */
    int ku;                  /* DUAL: half width; AFFINE2: width of U (k % ku)                                    */
    int kind, ns;
} rsb_opnd_t;

enum {
    RSB_EPI_BIAS_STATS = 0,  /* y = acc + bias[n]; stats[2N] += (sum y, sum y^2)                                  */
    RSB_EPI_RELU_MASK = 1    /* y = z > 0 ? acc : 0 with z = sc[n]*Yl[r,n]+sh[n] (+ dual: second half at N+n);
                                stats[2N or 3N] += (sum y, sum y*xhat1 [, sum y*xhat2]), xhat = (Yl-mu)*inv      */
};
typedef struct {
    float *Y;                /* [rows, ldy] or NULL                                                               */
    int ldy;
    const float *bias;       /* BIAS_STATS, may be NULL                                                           */
    double *stats;           /* caller-zeroed, may be NULL                                                        */
    const float *Yl;         /* RELU_MASK: stored pre-BatchNorm output of the layer below, [rows, ldl]             */
    int ldl;
    const float *sc, *sh, *mu, *inv;
    int kind, dual;
    const int *scatter;      /* BIAS_STATS without bias/stats: Y[scatter[r], :] += acc[r, :] (vector reductions) instead of
                                Y[r, :] = acc - the grouping backward fused into the input-gradient GEMM, or NULL         */
} rsb_epi_t;

/* Kernel generation of rsb_gemm_rows / rsb_gemm_wgrad: 0 (default) = the TMA-fed kernels (csrc/mlp_tc2.cu) whenever
 * every stored tensor of the operands is 16-byte aligned with a row pitch that is a multiple of 4 floats, else the
 * first-generation cp.async kernels (csrc/mlp_tc.cu); 1 = always the first generation (A/B comparisons, tests). */
void rsb_tc_set_generation(int gen);
/* SM budget of the persistent row GEMMs that follow (0 = all SMs): for launches the caller knows to overlap a long-running
 * cluster launch on another stream (the next level's FPS), so that no CTA has to wait for a second wave. */
void rsb_tc_set_sm_budget(int sms);
/* Y = A @ W^T with W pre-split by rsb_linear_tc_prep_weight(N, A->K, ...). */
int rsb_gemm_rows(long rows, int N, const rsb_opnd_t *A, const float *Wp, const rsb_epi_t *E, cudaStream_t stream);
/* dW[m, n] += sum_r G(r, m) * X(r, n);  dW is [G->K, ldw] fp32, accumulated with atomics (caller zeroes it). */
int rsb_gemm_wgrad(long rows, const rsb_opnd_t *G, const rsb_opnd_t *X, float *dW, int ldw, cudaStream_t stream);

/* Per-channel helpers of the fused shared MLP (train-mode BatchNorm, max-pool over nsample):
 * rsb_bn_finalize: stats (sum, sum^2 over `rows`) -> sc = gamma/sqrt(var+eps), sh = beta - mean*sc, mu, inv; updates
 *   running_mean/var in place when given (torch semantics: biased variance to normalise, unbiased for the running value).
 * rsb_pool_forward: out[g,c] = max_s relu(sc*Y[g*ns+s,c]+sh), arg = first maximising s.
 * rsb_pool_backward_stats: dm = dOut masked by (pooled value > 0); stats[2C] += (sum dm, sum dm*xhat at the arg-max row).
 * rsb_bn_backward_coef: coefficients of dY = a*dZ + b*Y + d (BatchNorm backward) + dgamma/dbeta, from
 *   stats = (sum dZ, sum dZ*xhat [, second xhat when dual]); bit 1 of `dual` marks a BatchNorm that normalised with
 *   its running statistics (eval mode / frozen): then dY = sc*dZ (b = d = 0), dgamma/dbeta unchanged.
 * rsb_bn_eval_coef: sc/sh/mu/inv of an eval-mode BatchNorm from its running statistics (gamma/beta may be NULL). */
int rsb_bn_eval_coef(int C, const float *gamma, const float *beta, const float *running_mean, const float *running_var,
                     float eps, float *sc, float *sh, float *mu, float *inv, cudaStream_t stream);
int rsb_bn_finalize(int C, long rows, const double *stats, const float *gamma, const float *beta, float eps,
                    float momentum, float *running_mean, float *running_var, float *sc, float *sh, float *mu,
                    float *inv, cudaStream_t stream);
/* running_mean/var (+ num_batches_tracked, int64, may be NULL) of one BatchNorm from its batch sums / sums of squares. */
int rsb_bn_update_running(int C, long rows, const double *sum, const double *sumsq, float momentum, float *running_mean,
                          float *running_var, long long *num_batches_tracked, cudaStream_t stream);
int rsb_pool_forward(long G, int ns, int C, const float *Y, int ldy, const float *sc, const float *sh, float *out,
                     int *arg, cudaStream_t stream);
int rsb_pool_backward_stats(long G, int ns, int C, const float *dOut, const int *arg, const float *Y, int ldy,
                            const float *sc, const float *sh, const float *mu, const float *inv, float *dm,
                            double *stats, cudaStream_t stream);
/* rsb_pool_bn_backward_dense: in place Y := a*(arg==sample ? dm : 0) + b*Y + d  (max-pool + BatchNorm backward of the
 *   last shared-MLP layer; Y, the stored pre-BN output, becomes dL/dY). */
int rsb_pool_bn_backward_dense(long G, int ns, int C, const float *dm, const int *arg, float *Y, int ldy, const float *a,
                               const float *b, const float *d, cudaStream_t stream);
/* rsb_bn_apply: out = [relu](sc*Y + sh) (materialised BatchNorm(+ReLU) output). */
int rsb_bn_apply(long rows, int C, const float *Y, int ldy, const float *sc, const float *sh, int relu, float *out,
                 int ldo, cudaStream_t stream);
/* rsb_bn_relu_backward: dA := src where (sc*Y+sh [+ second half when dual]) > 0 else 0  (ReLU backward of a layer whose
 *   pre-BatchNorm output Y was stored; src NULL = in place), and stats[2C | 3C] += (sum dZ, sum dZ*xhat_1 [, sum dZ*xhat_2]). */
int rsb_bn_relu_backward(long rows, int C, const float *src, int lds, float *dA, int ldd, const float *Y, int ldy,
                         const float *sc, const float *sh, const float *mu, const float *inv, int dual, double *stats,
                         cudaStream_t stream);
int rsb_bn_backward_coef(int C, long rows, const double *stats, int dual, const float *sc, const float *mu,
                         const float *inv, float *a, float *b, float *d, float *dgamma, float *dbeta,
                         cudaStream_t stream);

/* ------------------------------------------------------------------ whole-scene inference (segmentation/tool/test_s3dis.py)
 * rsb_scene_vote:   pred[idx[r], :] += softmax(logits[r, :]), count[idx[r]] += 1     test_s3dis.py:208-213 (logits row pitch ld)
 * rsb_scene_decide: label[p] = argmax_c pred[p, c] / count[p]                        test_s3dis.py:217
 * rsb_label_median: out[p] = lower median of label[nbr[p, 0..k)], k <= 128           util/utils.py:242-244 (torch.median) */
int rsb_scene_vote(long rows, int num_class, const float *logits, int ld, const long long *idx, float *pred, float *count,
                   cudaStream_t stream);
int rsb_scene_decide(long n, int num_class, const float *pred, const float *count, int *label, cudaStream_t stream);
int rsb_label_median(long n, int k, const int *nbr, const int *label, int *out, cudaStream_t stream);

/* Cross-entropy of the segmentation step (segmentation/tool/train.py: nn.CrossEntropyLoss(ignore_index)), one pass:
 * loss = mean over rows with target != ignore_index of (logsumexp(x) - x[target]); grad = softmax(x) - onehot(target) is written
 * by the forward (unscaled, zero on ignored rows) and scaled by upstream / #valid rows in the backward.  acc: fp64[2], pre-zeroed. */
int rsb_cross_entropy_forward(long rows, int num_class, const float *logits, int ld, const long long *target,
                              long long ignore_index, float *grad, int ldg, double *acc, float *loss, cudaStream_t stream);
int rsb_cross_entropy_backward(long n_elements, float *grad, const double *acc, const float *upstream, cudaStream_t stream);

/* ------------------------------------------------------------------ classification evaluation harness
 * replaces the torch-native FPS resampling of sample()            classification/modules/pointnet2_utils.py:62-75, :114-124
 * feat [b,c,n] channel-first (xyz = channels 0..2), start [b] int64 first picks; idx [b,m] int64, out [b,c,m] = feat[:, :, idx].
 * torch semantics: un-fused squared distance ((dx*dx + dy*dy) + dz*dz), strict-< running minimum, first-maximum picks. */
int rsb_fps_native_sample(int b, int c, int n, int m, const float *feat, const long long *start, long long *idx, float *out,
                          cudaStream_t stream);

/* ------------------------------------------------------------------ segmentation input pipeline (grid subsampling, crop)
 * rsb_coord_min:     out3 (pre-set to +inf) = column minima of coord [n,3]                         data_util.py:37 (coord - min)
 * rsb_voxel_keys:    key[i] = FNV64-1A(floor((coord[i] - cmin) / voxel_size)) ^ 2^63 (signed-sortable)   voxelize_utils.py:4-17, :40
 * rsb_voxel_runs:    runs of equal keys in the SORTED key array: start[r] = first position of run r, scalars[0] = number of
 *                    runs (= occupied voxels); scratch = ceil(n / 1024) ints                        voxelize_utils.py:50-51 (np.unique)
 * rsb_voxel_counts:  count[r], count_max[0] (pre-zeroed)
 * rsb_voxel_pick:    out[r] = order[start[r] + draw[r] % count[r]]                                  voxelize_utils.py:53-55
 * rsb_seed_distance: dist[i] = |coord[i] - coord[seed]|^2, un-fused fp32 like numpy                 data_util.py:47
 * rsb_coord_max:     out3 (pre-set to -inf) = column maxima of coord [n,3]
 * rsb_voxel_keys_ravel: key[i] = ravel_hash_vec(floor(coord / voxel_size))[i] (rank inside the occupied box)   voxelize_utils.py:20-35, :44
 * whole-scene crop planner (segmentation/tool/test_s3dis.py:143-158, data_process):
 * rsb_argmin_f64:    work[1] = index of the FIRST smallest element of v [n] (np.argmin; NaN-free input), work = 2 x uint64   :145
 * rsb_seed_distance_dev: rsb_seed_distance with the seed index read from device memory               :146
 * rsb_crop_update:   priority[crop] += (1 - dist[crop] / dist[crop[m-1]])^2 (fp32, added in fp64), covered[crop] = 1,
 *                    *n_covered += rows covered for the first time; crop = m distinct rows in ascending distance   :150-158 */
int rsb_coord_min(long n, const float *coord, float *out3, cudaStream_t stream);
int rsb_voxel_keys(long n, const float *coord, const float *cmin, float voxel_size, long long *key, cudaStream_t stream);
int rsb_voxel_runs(long n, const long long *sorted_key, int *scratch, int *start, int *scalars, cudaStream_t stream);
int rsb_voxel_counts(int n_runs, long n, const int *start, int *count, int *count_max, cudaStream_t stream);
int rsb_voxel_pick(int n_runs, const int *start, const int *count, const long long *draw, const long long *order,
                   long long *out, cudaStream_t stream);
int rsb_seed_distance(long n, const float *coord, long seed, float *dist, cudaStream_t stream);
int rsb_coord_max(long n, const float *coord, float *out3, cudaStream_t stream);
int rsb_voxel_keys_ravel(long n, const float *coord, const float *cmin, const float *cmax, float voxel_size, long long *key,
                         cudaStream_t stream);
int rsb_argmin_f64(long n, const double *v, unsigned long long *work, cudaStream_t stream);
int rsb_seed_distance_dev(long n, const float *coord, const unsigned long long *seed, float *dist, cudaStream_t stream);
int rsb_crop_update(int m, const long long *crop, const float *dist, double *priority, int *covered, int *n_covered,
                    cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* REPSURF_B200_H */
