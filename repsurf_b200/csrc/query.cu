// query.cu — ball query, k-nearest-neighbour (dense + packed) and 3-NN for sm_100a.
//
// Replaces (index-exact):
//   classification/modules/pointops/src/ballquery/ballquery_cuda_kernel.cu:47-101      (rule R3)
//   classification/modules/pointops/src/knnquery/knnquery_cuda_kernel.cu:6-72           (rule R4)
//   classification/modules/pointops/src/knnquery_heap/knnquery_heap_cuda_kernel.cu:53-110 (rule R5)
//   classification/modules/pointops/src/interpolation/interpolation_cuda_kernel.cu:134-176 (3-NN)
//   segmentation/modules/pointops/src/knnquery/knnquery_cuda_kernel.cu:65-116           (rule R5, offsets)
//
// The reference gives every query ONE thread that scans the whole cloud through L1 and keeps its
// top-k in local memory (double best[200] / float best_dist[100] heaps).  Here:
//   * kNN: a warp owns QW queries; candidate tiles are staged once per block in shared memory as
//     SoA and every lane evaluates 2 candidates x QW queries per step with Blackwell's packed
//     fp32x2 pipe (__fadd2_rn/__fmul2_rn/__ffma2_rn: exact IEEE per element, so rule R1 holds).
//     The top-k lives in REGISTERS, distributed over the warp as a sorted list (KPL slots / lane);
//     a candidate only leaves the fast loop if it beats the current k-th distance (~k ln(n/k)
//     times per query), then it is inserted with one redux.sync + shuffles.
//   * order semantics: the dense kernel is a stable insertion => sorted by (d2, index).  The heap
//     kernels return ascending d2 with a history-dependent order among EQUAL d2; the fast path is
//     provably identical unless two candidates tie on d2 at or inside the top-k boundary, which
//     is detected (tie flag) and replayed through an exact restatement of the heap (knn_heap_replay).
//   * ball query: a warp per query, ballot-ordered compaction keeps the ascending-index semantics,
//     early exit after nsample hits like the reference.
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int KNN_WARPS = 8;      // warps per block
constexpr int KNN_TILE = 2048;    // candidates per shared-memory tile (SoA, 24 KB)

struct QueryParams {
    const float *xyz;        // candidates: dense [b,n,3] | packed [sum n,3]
    const float *new_xyz;    // queries:    dense [b,m,3] | packed [sum m,3]
    const int *offset;       // packed: candidate segment ends [b]
    const int *new_offset;   // packed: query segment ends [b]
    int *idx;                // [.., k]
    float *dist2;            // [.., k] or nullptr
    int b, n, m, k;
    int packed;
    int sqrt_out;            // write sqrt(d2) instead of d2 (fuses the wrappers' torch.sqrt)
};

__device__ __forceinline__ bool lex_less(float d1, int i1, float d2, int i2)
{
    return d1 < d2 || (d1 == d2 && i1 < i2);
}

// Exact restatement of the reference max-heap (reheap / heap_sort) for ONE query, executed by a
// whole warp: distances in parallel, heap operations by lane 0 in shared memory.
// classification/.../knnquery_heap_cuda_kernel.cu:21-50, segmentation/.../knnquery_cuda_kernel.cu:21-48.
__device__ void knn_heap_replay(const float *__restrict__ cand, int start, int end, int index_base, int sentinel,
                                float qx, float qy, float qz, int k, float *hd, int *hi, int lane)
{
    for (int i = lane; i < k; i += 32) { hd[i] = 1e10f; hi[i] = sentinel; }
    __syncwarp();
    for (int base = start; base < end; base += 32) {
        const int c = base + lane;
        float d = CUDART_INF_F;
        if (c < end) d = rsb_sqdist(qx, qy, qz, cand[(size_t)c * 3], cand[(size_t)c * 3 + 1], cand[(size_t)c * 3 + 2]);
        unsigned mask = __ballot_sync(0xffffffffu, d < hd[0]);
        while (mask) {
            const int l = __ffs(mask) - 1;
            mask &= mask - 1;
            const float dl = __shfl_sync(0xffffffffu, d, l);
            if (lane == 0 && dl < hd[0]) {
                hd[0] = dl;
                hi[0] = base + l - index_base;
                int root = 0, child = 1;
                while (child < k) {
                    if (child + 1 < k && hd[child + 1] > hd[child]) child++;
                    if (hd[root] > hd[child]) break;
                    const float tf = hd[root]; hd[root] = hd[child]; hd[child] = tf;
                    const int ti = hi[root]; hi[root] = hi[child]; hi[child] = ti;
                    root = child;
                    child = root * 2 + 1;
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        for (int i = k - 1; i > 0; i--) {
            float tf = hd[0]; hd[0] = hd[i]; hd[i] = tf;
            int ti = hi[0]; hi[0] = hi[i]; hi[i] = ti;
            int root = 0, child = 1;
            while (child < i) {
                if (child + 1 < i && hd[child + 1] > hd[child]) child++;
                if (hd[root] > hd[child]) break;
                tf = hd[root]; hd[root] = hd[child]; hd[child] = tf;
                ti = hi[root]; hi[root] = hi[child]; hi[child] = ti;
                root = child;
                child = root * 2 + 1;
            }
        }
    }
    __syncwarp();
}

// KPL: list slots per lane (capacity 32*KPL >= k).  QW: queries per warp.  HEAP: heap-order semantics.
template <int KPL, int QW, bool HEAP>
__global__ void __launch_bounds__(KNN_WARPS * 32) knn_kernel(QueryParams P)
{
    __shared__ __align__(16) float sx[KNN_TILE], sy[KNN_TILE], sz[KNN_TILE];
    __shared__ float heap_d[HEAP ? KNN_WARPS * 100 : 1];
    __shared__ int heap_i[HEAP ? KNN_WARPS * 100 : 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int QPB = KNN_WARPS * QW;

    // ---- which segment / which queries ---------------------------------------------------------
    int cstart, cend, qstart, qend, chunk, index_base;
    if (P.packed) {
        int blk = blockIdx.x, s = 0, prev_m = 0;
        for (; s < P.b; s++) {
            const int ms = P.new_offset[s] - prev_m;
            const int nb = (ms + QPB - 1) / QPB;
            if (blk < nb) break;
            blk -= nb;
            prev_m = P.new_offset[s];
        }
        if (s >= P.b) return;
        cstart = s ? P.offset[s - 1] : 0;
        cend = P.offset[s];
        qstart = prev_m;
        qend = P.new_offset[s];
        chunk = blk;
        index_base = 0;  // packed indices are global row ids
    } else {
        const int bi = blockIdx.y;
        cstart = bi * P.n;
        cend = cstart + P.n;
        qstart = bi * P.m;
        qend = qstart + P.m;
        chunk = blockIdx.x;
        index_base = cstart;  // dense indices are local to the cloud
    }
    const int k = P.k;
    const int sentinel_idx = HEAP ? (P.packed ? cstart : 0) : 0;
    const float sentinel_d = HEAP ? 1e10f : CUDART_INF_F;

    // ---- queries of this warp (negated: d = c + (-q)) --------------------------------------------
    const int q0 = qstart + chunk * QPB + warp * QW;
    float2 nqx[QW], nqy[QW], nqz[QW];
    float kth_d[QW];
    int kth_i[QW];
    float ld[QW][KPL];
    int li[QW][KPL];
    bool tie[QW];
#pragma unroll
    for (int q = 0; q < QW; q++) {
        int qi = q0 + q;
        if (qi >= qend) qi = qend - 1;
        if (qi < qstart) qi = qstart;  // (empty chunk guard; such warps write nothing)
        const float x = P.new_xyz[(size_t)qi * 3], y = P.new_xyz[(size_t)qi * 3 + 1], z = P.new_xyz[(size_t)qi * 3 + 2];
        nqx[q] = make_float2(-x, -x);
        nqy[q] = make_float2(-y, -y);
        nqz[q] = make_float2(-z, -z);
        kth_d[q] = sentinel_d;
        kth_i[q] = sentinel_idx;
        tie[q] = false;
#pragma unroll
        for (int s = 0; s < KPL; s++) { ld[q][s] = sentinel_d; li[q][s] = sentinel_idx; }
    }
    const int kth_lane = (k - 1) / KPL, kth_slot = (k - 1) % KPL;

    // ---- tiles ----------------------------------------------------------------------------------
    for (int tbase = cstart; tbase < cend; tbase += KNN_TILE) {
        const int tn = min(KNN_TILE, cend - tbase);
        __syncthreads();
        for (int i = tid; i < KNN_TILE; i += KNN_WARPS * 32) {
            float x = 1e30f, y = 1e30f, z = 1e30f;  // padding: distance overflows to +inf
            if (i < tn) {
                const float *p = P.xyz + (size_t)(tbase + i) * 3;
                x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
            }
            sx[i] = x; sy[i] = y; sz[i] = z;
        }
        __syncthreads();
        const int tn_pad = (tn + 63) & ~63;
        for (int base = 0; base < tn_pad; base += 64) {
            const int c = base + 2 * lane;
            const float2 cx = *reinterpret_cast<const float2 *>(&sx[c]);
            const float2 cy = *reinterpret_cast<const float2 *>(&sy[c]);
            const float2 cz = *reinterpret_cast<const float2 *>(&sz[c]);
            float2 d[QW];
            bool anyhit = false;
#pragma unroll
            for (int q = 0; q < QW; q++) {
                const float2 dx = __fadd2_rn(cx, nqx[q]);
                const float2 dy = __fadd2_rn(cy, nqy[q]);
                const float2 dz = __fadd2_rn(cz, nqz[q]);
                float2 t = __fmul2_rn(dy, dy);
                t = __ffma2_rn(dx, dx, t);
                d[q] = __ffma2_rn(dz, dz, t);
                anyhit |= (d[q].x <= kth_d[q]) | (d[q].y <= kth_d[q]);
            }
            if (!__any_sync(0xffffffffu, anyhit)) continue;
            // ---- rare path: insert candidates that beat the current k-th -----------------------------
#pragma unroll
            for (int q = 0; q < QW; q++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const float dc = h ? d[q].y : d[q].x;
                    unsigned mask = __ballot_sync(0xffffffffu, dc <= kth_d[q]);
                    while (mask) {
                        const int l = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const float cd = __shfl_sync(0xffffffffu, dc, l);
                        const int ci = tbase + base + 2 * l + h - index_base;
                        if (HEAP) {
                            if (cd == kth_d[q]) { tie[q] = true; continue; }  // boundary tie -> replay
                            if (!(cd < kth_d[q])) continue;
                        } else {
                            if (!lex_less(cd, ci, kth_d[q], kth_i[q])) continue;
                        }
                        int cnt = 0;
                        bool eq = false;
#pragma unroll
                        for (int s = 0; s < KPL; s++) {
                            cnt += lex_less(ld[q][s], li[q][s], cd, ci) ? 1 : 0;
                            eq |= ld[q][s] == cd;
                        }
                        if (HEAP && __any_sync(0xffffffffu, eq)) tie[q] = true;
                        const int ins = __reduce_add_sync(0xffffffffu, cnt);
                        const float prev_d = __shfl_up_sync(0xffffffffu, ld[q][KPL - 1], 1);
                        const int prev_i = __shfl_up_sync(0xffffffffu, li[q][KPL - 1], 1);
#pragma unroll
                        for (int s = KPL - 1; s >= 0; s--) {
                            const int pos = lane * KPL + s;
                            const float sd = s > 0 ? ld[q][s > 0 ? s - 1 : 0] : prev_d;
                            const int si = s > 0 ? li[q][s > 0 ? s - 1 : 0] : prev_i;
                            if (pos > ins) { ld[q][s] = sd; li[q][s] = si; }
                            else if (pos == ins) { ld[q][s] = cd; li[q][s] = ci; }
                        }
                        float kd = ld[q][0];
                        int ki = li[q][0];
#pragma unroll
                        for (int s = 1; s < KPL; s++)
                            if (kth_slot == s) { kd = ld[q][s]; ki = li[q][s]; }
                        kth_d[q] = __shfl_sync(0xffffffffu, kd, kth_lane);
                        kth_i[q] = __shfl_sync(0xffffffffu, ki, kth_lane);
                    }
                }
            }
        }
    }

    // ---- write back ---------------------------------------------------------------------------------
#pragma unroll
    for (int q = 0; q < QW; q++) {
        const int qi = q0 + q;
        if (qi >= qend) continue;  // warp-uniform
        int *oi = P.idx + (size_t)qi * k;
        float *od = P.dist2 ? P.dist2 + (size_t)qi * k : nullptr;
        if (HEAP && tie[q]) {
            float *hd = heap_d + warp * 100;
            int *hi = heap_i + warp * 100;
            knn_heap_replay(P.xyz, cstart, cend, index_base, sentinel_idx, -nqx[q].x, -nqy[q].x, -nqz[q].x, k, hd, hi, lane);
            for (int i = lane; i < k; i += 32) {
                oi[i] = hi[i];
                if (od) od[i] = P.sqrt_out ? sqrtf(hd[i]) : hd[i];
            }
            __syncwarp();
            continue;
        }
#pragma unroll
        for (int s = 0; s < KPL; s++) {
            const int pos = lane * KPL + s;
            if (pos < k) {
                oi[pos] = li[q][s];
                if (od) od[pos] = P.sqrt_out ? sqrtf(ld[q][s]) : ld[q][s];
            }
        }
    }
}

template <int KPL, int QW, bool HEAP>
int knn_launch_t(const QueryParams &P, cudaStream_t stream)
{
    constexpr int QPB = KNN_WARPS * QW;
    dim3 grid;
    if (P.packed) grid = dim3(RSB_DIVUP(P.m, QPB) + P.b);
    else grid = dim3(RSB_DIVUP(P.m, QPB), P.b);
    if (grid.x == 0 || grid.y == 0) return 0;
    knn_kernel<KPL, QW, HEAP><<<grid, KNN_WARPS * 32, 0, stream>>>(P);
    RSB_CHECK_LAUNCH("knn_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

template <bool HEAP>
int knn_launch(const QueryParams &P, cudaStream_t stream)
{
    if (P.k <= 32) return knn_launch_t<1, 4, HEAP>(P, stream);
    if (P.k <= 64) return knn_launch_t<2, 4, HEAP>(P, stream);
    if (P.k <= 128) return knn_launch_t<4, 2, HEAP>(P, stream);
    return knn_launch_t<7, 2, HEAP>(P, stream);
}

// ---- ball query: one warp per query ---------------------------------------------------------------
__global__ void __launch_bounds__(256) ballquery_kernel(int b, int n, int m, float radius, int nsample,
                                                        const float *__restrict__ new_xyz,
                                                        const float *__restrict__ xyz, int *__restrict__ idx)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int bi = blockIdx.y;
    if (q >= m) return;
    const float *pts = xyz + (size_t)bi * n * 3;
    const float *qp = new_xyz + ((size_t)bi * m + q) * 3;
    int *out = idx + ((size_t)bi * m + q) * nsample;
    const float r2 = __fmul_rn(radius, radius);
    const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    int cnt = 0, first = 0;
    for (int base = 0; base < n && cnt < nsample; base += 32) {
        const int k = base + lane;
        bool hit = false;
        if (k < n) {
            const float d2 = rsb_sqdist(qx, qy, qz, __ldg(pts + (size_t)k * 3), __ldg(pts + (size_t)k * 3 + 1),
                                        __ldg(pts + (size_t)k * 3 + 2));
            hit = d2 < r2;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (mask) {
            if (cnt == 0) first = base + __ffs(mask) - 1;
            const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
            if (hit && pos < nsample) out[pos] = k;
            cnt += __popc(mask);
        }
    }
    // pad with the first hit; all-zero when the ball is empty (the reference relies on a pre-zeroed
    // buffer, classification/modules/pointops/functions/pointops.py:220 — written here instead)
    if (cnt > nsample) cnt = nsample;
    for (int i = cnt + lane; i < nsample; i += 32) out[i] = first;
}

}  // namespace

RSB_EXPORT int rsb_ballquery(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                             const float *xyz, int *idx, cudaStream_t stream)
{
    RSB_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample >= 1, "bad sizes");
    if (b == 0 || m == 0) return 0;
    dim3 grid(RSB_DIVUP(m, 8), b);
    ballquery_kernel<<<grid, 256, 0, stream>>>(b, n, m, radius, nsample, new_xyz, xyz, idx);
    RSB_CHECK_LAUNCH("ballquery_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

RSB_EXPORT int rsb_knnquery_dense(int b, int n, int m, int nsample, const float *xyz, const float *new_xyz,
                                  int *idx, float *dist2, cudaStream_t stream)
{
    RSB_REQUIRE(nsample >= 1 && nsample <= 200, "nsample must be in [1, 200] (reference local array bound)");
    RSB_REQUIRE(n >= 1, "empty cloud");
    if (b == 0 || m == 0) return 0;
    QueryParams P = {};
    P.xyz = xyz; P.new_xyz = new_xyz; P.idx = idx; P.dist2 = dist2;
    P.b = b; P.n = n; P.m = m; P.k = nsample; P.packed = 0; P.sqrt_out = 0;
    return knn_launch<false>(P, stream);
}

RSB_EXPORT int rsb_knnquery_heap_dense(int b, int n, int m, int nsample, const float *xyz, const float *new_xyz,
                                       int *idx, float *dist2, cudaStream_t stream)
{
    RSB_REQUIRE(nsample >= 1 && nsample <= 100, "nsample must be in [1, 100] (reference local array bound)");
    RSB_REQUIRE(n >= 1, "empty cloud");
    if (b == 0 || m == 0) return 0;
    QueryParams P = {};
    P.xyz = xyz; P.new_xyz = new_xyz; P.idx = idx; P.dist2 = dist2;
    P.b = b; P.n = n; P.m = m; P.k = nsample; P.packed = 0; P.sqrt_out = 0;
    return knn_launch<true>(P, stream);
}

RSB_EXPORT int rsb_knnquery_packed(int b, int m, int nsample, const float *xyz, const float *new_xyz,
                                   const int *offset, const int *new_offset, int *idx, float *dist,
                                   int sqrt_out, cudaStream_t stream)
{
    RSB_REQUIRE(nsample >= 1 && nsample <= 100, "nsample must be in [1, 100] (reference local array bound)");
    if (b == 0 || m == 0) return 0;
    QueryParams P = {};
    P.xyz = xyz; P.new_xyz = new_xyz; P.offset = offset; P.new_offset = new_offset; P.idx = idx; P.dist2 = dist;
    P.b = b; P.m = m; P.k = nsample; P.packed = 1; P.sqrt_out = sqrt_out;
    return knn_launch<true>(P, stream);
}

// 3-NN of `unknown` in `known`: the reference's 3-slot insertion is the stable (d2, index) order.
RSB_EXPORT int rsb_nearestneighbor(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                                   int *idx, cudaStream_t stream)
{
    RSB_REQUIRE(m >= 1, "empty known set");
    if (b == 0 || n == 0) return 0;
    QueryParams P = {};
    P.xyz = known; P.new_xyz = unknown; P.idx = idx; P.dist2 = dist2;
    P.b = b; P.n = m; P.m = n; P.k = 3; P.packed = 0; P.sqrt_out = 0;
    return knn_launch<false>(P, stream);
}
