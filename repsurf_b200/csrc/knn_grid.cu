// knn_grid.cu — exact k-nearest-neighbour search through a uniform grid (same results as the brute-force
// kernels in query.cu, which themselves are index-exact against the reference's
//   classification/modules/pointops/src/knnquery/knnquery_cuda_kernel.cu:6-72          (stable (d2, index) order)
//   segmentation/modules/pointops/src/knnquery/knnquery_cuda_kernel.cu:65-116          (heap order)
// The reference scans all n candidates for each of the m queries (13.4 G pairs for the umbrella search of
// one S3DIS batch).  Here every cloud is binned once into a uniform grid (~8 points per cell, counting sort
// into a float4 array carrying the original row id), and a warp per query visits the cells ring by ring around
// the query's cell, stopping when the k-th best SQUARED distance is provably smaller than the distance to the
// unexplored region.  Exactness:
//   * candidate distances use rsb_sqdist (rule R1), so every evaluated d2 equals the reference's bit for bit;
//   * the top-k list orders by (d2, original index) — independent of the visiting order;
//   * termination is conservative (strict inequality with a relative margin against the cube face distance), so
//     no unvisited candidate can have d2 <= the k-th distance;
//   * heap-order kernels: any tie on d2 at or inside the top-k boundary is detected and that query is replayed
//     with the exact heap restatement over the whole cloud (knn_heap_replay_grid), as in query.cu.
#include "common.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace {

constexpr int GRID_STAGE = 320;     // float4 candidates per warp that a pass may stage in shared memory (5 KB)
constexpr int GRID_WARPS = 8;

struct SegGrid {
    float ox, oy, oz, inv_h, h;
    int gx, gy, gz;
    int cell_base;      // offset of this segment's cells in the global cell arrays
    int cstart, cend;   // candidate rows of this segment
};

struct GridParams {
    const float *xyz;          // candidates
    const int *offset;         // packed: candidate segment ends (device) or nullptr (dense)
    int b, n;                  // segments; dense: points per segment
    SegGrid *seg;              // [b]
    float *bbox;               // [b][6]
    int *cell_cnt;             // [n_total + b + 1]  counts -> exclusive starts
    int *cursor;               // [n_total + b + 1]
    int *cid;                  // [n_total]
    float4 *sorted;            // [n_total]
    float target_per_cell;     // average points per grid cell the geometry aims for (depends on k, see rsb_knnquery_grid)
};

__device__ __forceinline__ void seg_range(const GridParams &P, int s, int &c0, int &c1)
{
    if (P.offset) { c0 = s ? P.offset[s - 1] : 0; c1 = P.offset[s]; }
    else { c0 = s * P.n; c1 = c0 + P.n; }
}

// one block per segment: bounding box, then grid geometry
__global__ void __launch_bounds__(1024) grid_setup_kernel(GridParams P)
{
    const int s = blockIdx.x;
    int c0, c1;
    seg_range(P, s, c0, c1);
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float v = P.xyz[(size_t)i * 3 + a];
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    }
    __shared__ float sm[6][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        for (int o = 16; o; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0) { sm[a][warp] = lo[a]; sm[3 + a][warp] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int a = 0; a < 3; a++) {
            float l = CUDART_INF_F, h = -CUDART_INF_F;
            for (int w = 0; w < nw; w++) { l = fminf(l, sm[a][w]); h = fmaxf(h, sm[3 + a][w]); }
            lo[a] = l; hi[a] = h;
        }
        SegGrid g;
        const int npts = c1 - c0;
        g.cstart = c0; g.cend = c1;
        g.cell_base = c0 + s;
        if (npts <= 0) { g.ox = g.oy = g.oz = 0.f; g.h = 1.f; g.inv_h = 1.f; g.gx = g.gy = g.gz = 1; P.seg[s] = g; return; }
        const float ex = fmaxf(hi[0] - lo[0], 1e-12f), ey = fmaxf(hi[1] - lo[1], 1e-12f), ez = fmaxf(hi[2] - lo[2], 1e-12f);
        float h = cbrtf(ex * ey * ez * P.target_per_cell / (float)npts);
        h = fmaxf(h, 1e-6f * fmaxf(ex, fmaxf(ey, ez)));
        int gx, gy, gz;
        for (int it = 0; it < 64; it++) {
            gx = max(1, min(1024, (int)ceilf(ex / h)));
            gy = max(1, min(1024, (int)ceilf(ey / h)));
            gz = max(1, min(1024, (int)ceilf(ez / h)));
            if ((long)gx * gy * gz <= (long)max(npts, 1)) break;
            h *= 1.26f;
        }
        g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
        g.h = h; g.inv_h = 1.f / h;
        g.gx = gx; g.gy = gy; g.gz = gz;
        P.seg[s] = g;
    }
}

__device__ __forceinline__ int cell_coord(float v, float o, float inv_h, int n)
{
    int c = (int)floorf((v - o) * inv_h);
    return max(0, min(n - 1, c));
}

__global__ void __launch_bounds__(256) grid_count_kernel(GridParams P, int n_total)
{
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n_total; i += gridDim.x * 256) {
        int s;
        if (P.offset) { s = 0; while (i >= P.offset[s]) s++; }
        else s = i / P.n;
        const SegGrid g = P.seg[s];
        const int cx = cell_coord(P.xyz[(size_t)i * 3], g.ox, g.inv_h, g.gx);
        const int cy = cell_coord(P.xyz[(size_t)i * 3 + 1], g.oy, g.inv_h, g.gy);
        const int cz = cell_coord(P.xyz[(size_t)i * 3 + 2], g.oz, g.inv_h, g.gz);
        const int cid = g.cell_base + (cz * g.gy + cy) * g.gx + cx;
        P.cid[i] = cid;
        atomicAdd(P.cell_cnt + cid, 1);
    }
}

// one block per segment: exclusive scan of its cell counts (in place), absolute positions in `sorted`
__global__ void __launch_bounds__(1024) grid_scan_kernel(GridParams P)
{
    const int s = blockIdx.x;
    const SegGrid g = P.seg[s];
    const int ncell = g.gx * g.gy * g.gz;
    int *cnt = P.cell_cnt + g.cell_base;
    __shared__ int warp_sum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = g.cstart;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < ncell + 1; base += 1024) {       // one extra entry: end of the last cell
        const int i = base + threadIdx.x;
        const int v = i < ncell ? cnt[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sum[lane];
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            warp_sum[lane] = w;
        }
        __syncthreads();
        const int excl = carry + (warp ? warp_sum[warp - 1] : 0) + x - v;
        if (i <= ncell) cnt[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) grid_scatter_kernel(GridParams P, int n_total)
{
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n_total; i += gridDim.x * 256) {
        const int cid = P.cid[i];
        const int pos = P.cell_cnt[cid] + atomicAdd(P.cursor + cid, 1);
        P.sorted[pos] = make_float4(P.xyz[(size_t)i * 3], P.xyz[(size_t)i * 3 + 1], P.xyz[(size_t)i * 3 + 2], __int_as_float(i));
    }
}

// ---- query ---------------------------------------------------------------------------------------------------
struct GridQuery {
    const float *xyz;          // original candidates (heap replay)
    const float *new_xyz;      // queries
    const int *new_offset;     // packed: query segment ends, or nullptr (dense)
    const SegGrid *seg;
    const int *cell_start;     // exclusive starts (cell_cnt after the scan)
    const float4 *sorted;
    int *idx;
    float *dist2;
    int b, m_dense, m_total, k, packed, sqrt_out;
    unsigned long long *counters;   // optional [3]: candidates evaluated, cell ranges scanned, tie replays (measurement only)
};

__device__ __forceinline__ bool lex_less_g(float d1, int i1, float d2, int i2) { return d1 < d2 || (d1 == d2 && i1 < i2); }

// exact heap restatement (same as query.cu::knn_heap_replay), reading the ORIGINAL candidate order
__device__ void knn_heap_replay_grid(const float *__restrict__ cand, int start, int end, int index_base, int sentinel,
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

// ---- k <= 32: the top-k list is ONE 64-bit key per lane, (distance bits << 32) | index: squared distances are non-negative
// floats, so unsigned key order == lexicographic (distance, index) order.  A batch of 32 candidates that beats the current
// k-th key is either inserted one by one (few survivors: ballot + popc rank, two shuffles) or, when many survive (the first
// batches of every query), sorted by a 15-stage bitonic network and merged with the list in 6 more stages - a fixed ~170
// instructions instead of ~40 per inserted candidate (ncu, profiles/r02_ncu_full_seg.md: the kernel is issue-bound, 66-73 %).
__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src)
{
    const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src), hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long shfl_xor64(unsigned long long v, int m)
{
    const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m), hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long shfl_up64(unsigned long long v, int d)
{
    const unsigned lo = __shfl_up_sync(0xffffffffu, (unsigned)v, d), hi = __shfl_up_sync(0xffffffffu, (unsigned)(v >> 32), d);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long knn_key(float d, int i) { return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i; }

template <int KPL, bool HEAP, bool STAGED>
__global__ void __launch_bounds__(GRID_WARPS * 32) knn_grid_kernel(GridQuery P)
{
    __shared__ float heap_d[HEAP ? GRID_WARPS * 100 : 1];
    __shared__ int heap_i[HEAP ? GRID_WARPS * 100 : 1];
    // per-warp staging area: the candidates of one pass (<= 32 cell ranges) are fetched into it by bulk copies (the TMA engine's
    // 1-D mode: every range of the cell-sorted float4 array is a 16-byte aligned run), completion on the warp's mbarrier
    // (STAGED; measured on the S3DIS step: 2.64 ms -> 2.87 ms per step - the extra registers and shared memory cost a resident
    // block per SM and the kernel is issue-bound, not latency-bound - so the default is the direct path: RSB_KNN_STAGE=1 selects this)
    __shared__ __align__(16) float4 stage[STAGED ? GRID_WARPS : 1][STAGED ? GRID_STAGE : 1];
    __shared__ __align__(8) unsigned long long stage_bar[GRID_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = blockIdx.x * GRID_WARPS + warp;
    if (qi >= P.m_total) return;
    const uint32_t bar = rsb_smem_addr(&stage_bar[warp]), stage_base = rsb_smem_addr(&stage[STAGED ? warp : 0][0]);
    if (STAGED && lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t stage_phase = 0;
    int s;
    if (P.packed) {
        // segment of the query: number of cumulative ends <= qi, 32 segments per ballot (one load instead of a dependent chain)
        s = 0;
        for (int b0 = 0; b0 < P.b; b0 += 32) {
            const int e = b0 + lane < P.b ? __ldg(P.new_offset + b0 + lane) : 0x7fffffff;
            s += __popc(__ballot_sync(0xffffffffu, qi >= e));
        }
    } else s = qi / P.m_dense;
    const SegGrid g = P.seg[s];
    const int k = P.k;
    const int index_base = P.packed ? 0 : g.cstart;
    const int sentinel_idx = HEAP ? (P.packed ? g.cstart : 0) : 0;
    const float sentinel_d = HEAP ? 1e10f : CUDART_INF_F;
    const float qx = P.new_xyz[(size_t)qi * 3], qy = P.new_xyz[(size_t)qi * 3 + 1], qz = P.new_xyz[(size_t)qi * 3 + 2];

    float ld[KPL];
    int li[KPL];
#pragma unroll
    for (int t = 0; t < KPL; t++) { ld[t] = sentinel_d; li[t] = sentinel_idx; }
    float kth_d = sentinel_d;
    int kth_i = sentinel_idx;
    bool tie = false;
    const int kth_lane = (k - 1) / KPL, kth_slot = (k - 1) % KPL;
    unsigned long long L = knn_key(sentinel_d, sentinel_idx), kth_key = L;      // KPL == 1: this lane's list entry, the k-th key

    const int cx = cell_coord(qx, g.ox, g.inv_h, g.gx), cy = cell_coord(qy, g.oy, g.inv_h, g.gy), cz = cell_coord(qz, g.oz, g.inv_h, g.gz);
    const int *cs = P.cell_start + g.cell_base;
    const int rmax = max(max(g.gx, g.gy), g.gz);

    unsigned n_cand = 0, n_ranges = 0;
    // One call scans up to 32 cell ranges, lane t holding range t = [my_s, my_e) of the cell-sorted point array: the ranges are
    // concatenated virtually and walked in full 32-candidate batches (a 27-cell neighbourhood is ~10 ranges of ~20 points: one
    // batch PER RANGE left a third of the lanes idle and chained two dependent global loads per range)
    auto scan_ranges = [&](int my_s, int my_e) {
        const int len = max(my_e - my_s, 0);
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        n_cand += (unsigned)total;
        n_ranges += (unsigned)__popc(__ballot_sync(0xffffffffu, len > 0));
        // passes that fit the staging area: one bulk copy per range, all in flight at once, then the candidates are read from
        // shared memory as one contiguous array (no per-batch global-load latency, no search for a position's range)
        const bool staged = STAGED && total > 0 && total <= GRID_STAGE;
        if (staged) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier reads of the area vs the copies that refill it
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)total * 16u) : "memory");
            __syncwarp();
            if (len > 0)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 stage_base + (uint32_t)(incl - len) * 16u), "l"(P.sorted + my_s), "r"((uint32_t)len * 16u), "r"(bar)
                             : "memory");
            uint32_t done;
            do {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(done) : "r"(bar), "r"(stage_phase) : "memory");
            } while (!done);
            stage_phase ^= 1u;
        }
        for (int base = 0; base < total; base += 32) {
            const int pos = base + lane;
            const int j1 = 1, j = pos < total ? 0 : 1;            // "j < j1"  <=>  this lane holds a candidate
            float d = CUDART_INF_F;
            int ci = 0;
            if (staged) {
                if (j < j1) {
                    const float4 p = stage[STAGED ? warp : 0][STAGED ? pos : 0];
                    d = rsb_sqdist(qx, qy, qz, p.x, p.y, p.z);
                    ci = __float_as_int(p.w) - index_base;
                }
            } else {
                // range of this position: first lane whose inclusive prefix exceeds it (binary search over the warp's prefixes)
                int slot = 0;
#pragma unroll
                for (int st = 16; st > 0; st >>= 1) {
                    const int v = __shfl_sync(0xffffffffu, incl, slot + st - 1);
                    if (v <= pos) slot += st;
                }
                slot = min(slot, 31);
                const int rs = __shfl_sync(0xffffffffu, my_s, slot), rexcl = __shfl_sync(0xffffffffu, incl - len, slot);
                if (j < j1) {
                    const float4 p = __ldg(P.sorted + rs + (pos - rexcl));
                    d = rsb_sqdist(qx, qy, qz, p.x, p.y, p.z);
                    ci = __float_as_int(p.w) - index_base;
                }
            }
            if constexpr (KPL == 1) {
                unsigned long long ck = j < j1 ? knn_key(d, ci) : ~0ull;
                // survivors: strictly better than the k-th key; HEAP semantics also needs to know about candidates that only
                // TIE with the k-th distance (they decide between the list and the exact replay)
                unsigned mask = __ballot_sync(0xffffffffu, ck < kth_key);
                if (HEAP && __any_sync(0xffffffffu, j < j1 && d == kth_d)) tie = true;
                const int ns = __popc(mask);
                if (ns == 0) continue;
                if (ns <= 4) {
                    while (mask) {
                        const int l = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const unsigned long long c = shfl64(ck, l);
                        if (!(c < kth_key)) {                                           // the k-th key moved meanwhile
                            if (HEAP && (unsigned)(c >> 32) == (unsigned)(kth_key >> 32)) tie = true;
                            continue;
                        }
                        if (HEAP && __any_sync(0xffffffffu, (unsigned)(L >> 32) == (unsigned)(c >> 32))) tie = true;
                        const int ins = __popc(__ballot_sync(0xffffffffu, L < c));      // list entries in front of it
                        const unsigned long long up = shfl_up64(L, 1);
                        L = lane > ins ? up : (lane == ins ? c : L);
                        kth_key = shfl64(L, k - 1);
                    }
                } else {
                    // bitonic sort of the batch (ascending), then merge: the 32 smallest of list + batch, sorted
                    if (!(ck < kth_key)) ck = ~0ull;
#pragma unroll
                    for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
                        for (int jj = k2 >> 1; jj > 0; jj >>= 1) {
                            const unsigned long long o = shfl_xor64(ck, jj);
                            const bool take_min = ((lane & jj) == 0) == ((lane & k2) == 0 || k2 == 32);
                            ck = take_min ? (ck < o ? ck : o) : (ck < o ? o : ck);
                        }
                    const unsigned long long rev = shfl64(ck, 31 - lane);                // descending batch
                    unsigned long long m = L < rev ? L : rev;
                    const unsigned long long dis = L < rev ? rev : L;                   // the 32 discarded keys
#pragma unroll
                    for (int jj = 16; jj > 0; jj >>= 1) {
                        const unsigned long long o = shfl_xor64(m, jj);
                        m = ((lane & jj) == 0) ? (m < o ? m : o) : (m < o ? o : m);
                    }
                    L = m;
                    kth_key = shfl64(L, k - 1);
                    if (HEAP) {
                        // equal distances next to each other among the first k + 1 entries, or a discarded key that ties with the
                        // new k-th distance: order / membership is decided by the heap's insertion history -> exact replay
                        const unsigned kd = (unsigned)(kth_key >> 32);
                        const unsigned prevd = __shfl_up_sync(0xffffffffu, (unsigned)(L >> 32), 1);
                        const bool adj = lane > 0 && lane <= k && lane < 32 && prevd == (unsigned)(L >> 32) && (unsigned)(L >> 32) != __float_as_uint(sentinel_d);
                        if (__any_sync(0xffffffffu, adj || (dis != ~0ull && (unsigned)(dis >> 32) == kd && kd != __float_as_uint(sentinel_d)))) tie = true;
                    }
                }
                kth_d = __uint_as_float((unsigned)(kth_key >> 32));
                kth_i = (int)(unsigned)kth_key;
                continue;
            }
            unsigned mask = __ballot_sync(0xffffffffu, j < j1 && d <= kth_d);
            while (mask) {
                const int l = __ffs(mask) - 1;
                mask &= mask - 1;
                const float cd = __shfl_sync(0xffffffffu, d, l);
                const int cidx = __shfl_sync(0xffffffffu, ci, l);
                if (HEAP) {
                    if (cd == kth_d) { tie = true; continue; }
                    if (!(cd < kth_d)) continue;
                } else {
                    if (!lex_less_g(cd, cidx, kth_d, kth_i)) continue;
                }
                int cnt = 0;
                bool eq = false;
#pragma unroll
                for (int t = 0; t < KPL; t++) {
                    cnt += lex_less_g(ld[t], li[t], cd, cidx) ? 1 : 0;
                    eq |= ld[t] == cd;
                }
                if (HEAP && __any_sync(0xffffffffu, eq)) tie = true;
                const int ins = __reduce_add_sync(0xffffffffu, cnt);
                const float prev_d = __shfl_up_sync(0xffffffffu, ld[KPL - 1], 1);
                const int prev_i = __shfl_up_sync(0xffffffffu, li[KPL - 1], 1);
#pragma unroll
                for (int t = KPL - 1; t >= 0; t--) {
                    const int pos = lane * KPL + t;
                    const float sd = t > 0 ? ld[t > 0 ? t - 1 : 0] : prev_d;
                    const int si = t > 0 ? li[t > 0 ? t - 1 : 0] : prev_i;
                    if (pos > ins) { ld[t] = sd; li[t] = si; }
                    else if (pos == ins) { ld[t] = cd; li[t] = cidx; }
                }
                float kd = ld[0];
                int ki = li[0];
#pragma unroll
                for (int t = 1; t < KPL; t++)
                    if (kth_slot == t) { kd = ld[t]; ki = li[t]; }
                kth_d = __shfl_sync(0xffffffffu, kd, kth_lane);
                kth_i = __shfl_sync(0xffffffffu, ki, kth_lane);
            }
        }
    };

    for (int r = 0; r <= rmax; r++) {
        const int z0 = max(cz - r, 0), z1 = min(cz + r, g.gz - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, g.gy - 1);
        const int x0 = max(cx - r, 0), x1 = min(cx + r, g.gx - 1);
        // every (z, y) row of the ring contributes two range slots: a shell row its whole x span (+ an empty slot), an inner
        // row its left and its right boundary cell; 32 slots per pass, each lane fetching the two cell starts of its own slot
        const int ny = y1 - y0 + 1, n_slots = 2 * (z1 - z0 + 1) * ny;
        for (int t0 = 0; t0 < n_slots; t0 += 32) {
            const int t = t0 + lane;
            int my_s = 0, my_e = 0;
            if (t < n_slots) {
                const int row = t >> 1, which = t & 1;
                const int z = z0 + row / ny, y = y0 + row % ny;
                const int rowbase = (z * g.gy + y) * g.gx;
                const bool shell_row = (abs(z - cz) == r) || (abs(y - cy) == r);
                if (shell_row) {
                    if (which == 0) { my_s = __ldg(cs + rowbase + x0); my_e = __ldg(cs + rowbase + x1 + 1); }
                } else if (which == 0) {
                    if (cx - r >= 0) { my_s = __ldg(cs + rowbase + cx - r); my_e = __ldg(cs + rowbase + cx - r + 1); }
                } else if (cx + r < g.gx && r > 0) {
                    my_s = __ldg(cs + rowbase + cx + r); my_e = __ldg(cs + rowbase + cx + r + 1);
                }
            }
            scan_ranges(my_s, my_e);
        }
        // conservative termination: distance to the nearest face of the explored cube that still has cells behind it
        float face = CUDART_INF_F;
        if (cx - r > 0) face = fminf(face, qx - (g.ox + (float)(cx - r) * g.h));
        if (cx + r < g.gx - 1) face = fminf(face, (g.ox + (float)(cx + r + 1) * g.h) - qx);
        if (cy - r > 0) face = fminf(face, qy - (g.oy + (float)(cy - r) * g.h));
        if (cy + r < g.gy - 1) face = fminf(face, (g.oy + (float)(cy + r + 1) * g.h) - qy);
        if (cz - r > 0) face = fminf(face, qz - (g.oz + (float)(cz - r) * g.h));
        if (cz + r < g.gz - 1) face = fminf(face, (g.oz + (float)(cz + r + 1) * g.h) - qz);
        if (face == CUDART_INF_F) break;                       // the cube covers the whole grid
        // cell membership was decided in fp32: allow an absolute slack of a few ulps of the coordinate magnitude
        // (queries outside the box, i.e. clamped cells, can have face <= slack: they keep expanding)
        const float slack = 1e-5f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + g.h * (float)rmax);
        if (face > slack && kth_d < (face - slack) * (face - slack) * 0.9999f) break;
    }

    if (P.counters && lane == 0) {
        atomicAdd(P.counters, (unsigned long long)n_cand);
        atomicAdd(P.counters + 1, (unsigned long long)n_ranges);
        if (HEAP && tie) atomicAdd(P.counters + 2, 1ull);
    }
    int *oi = P.idx + (size_t)qi * k;
    float *od = P.dist2 ? P.dist2 + (size_t)qi * k : nullptr;
    if (HEAP && tie) {
        float *hd = heap_d + warp * 100;
        int *hi = heap_i + warp * 100;
        knn_heap_replay_grid(P.xyz, g.cstart, g.cend, index_base, sentinel_idx, qx, qy, qz, k, hd, hi, lane);
        for (int i = lane; i < k; i += 32) {
            oi[i] = hi[i];
            if (od) od[i] = P.sqrt_out ? sqrtf(hd[i]) : hd[i];
        }
        return;
    }
    if constexpr (KPL == 1) {
        ld[0] = __uint_as_float((unsigned)(L >> 32));
        li[0] = (int)(unsigned)L;
    }
#pragma unroll
    for (int t = 0; t < KPL; t++) {
        const int pos = lane * KPL + t;
        if (pos < k) {
            oi[pos] = li[t];
            if (od) od[pos] = P.sqrt_out ? sqrtf(ld[t]) : ld[t];
        }
    }
}

template <bool HEAP>
int launch_query(const GridQuery &Q, cudaStream_t stream)
{
    const int blocks = RSB_DIVUP(Q.m_total, GRID_WARPS);
    if (blocks == 0) return 0;
    static const bool stage = [] { const char *v = getenv("RSB_KNN_STAGE"); return v && v[0] && v[0] != '0'; }();
    if (Q.k <= 32 && stage) knn_grid_kernel<1, HEAP, true><<<blocks, GRID_WARPS * 32, 0, stream>>>(Q);
    else if (Q.k <= 32) knn_grid_kernel<1, HEAP, false><<<blocks, GRID_WARPS * 32, 0, stream>>>(Q);
    else if (Q.k <= 64) knn_grid_kernel<2, HEAP, false><<<blocks, GRID_WARPS * 32, 0, stream>>>(Q);
    else if (Q.k <= 128) knn_grid_kernel<4, HEAP, false><<<blocks, GRID_WARPS * 32, 0, stream>>>(Q);
    else knn_grid_kernel<7, HEAP, false><<<blocks, GRID_WARPS * 32, 0, stream>>>(Q);
    RSB_CHECK_LAUNCH("knn_grid_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// bytes of scratch needed by rsb_knnquery_grid for n_total candidate points in b segments
// measurement hook (bench.py): device buffer of 3 x u64 that every grid search adds its work to (candidates evaluated, cell
// ranges scanned, queries that went through the exact tie replay); NULL (default) disables the counting
static unsigned long long *g_knn_counters = nullptr;
RSB_EXPORT void rsb_knn_grid_set_counters(unsigned long long *dev_counters) { g_knn_counters = dev_counters; }

RSB_EXPORT long rsb_knn_grid_workspace_bytes(int n_total, int b)
{
    size_t bytes = 0;
    bytes += align_up(sizeof(SegGrid) * (size_t)b, 256);
    bytes += align_up(sizeof(float) * 6 * (size_t)b, 256);
    bytes += 2 * align_up(sizeof(int) * ((size_t)n_total + b + 1), 256);
    bytes += align_up(sizeof(int) * (size_t)n_total, 256);
    bytes += align_up(sizeof(float4) * (size_t)n_total, 256);
    return (long)bytes;
}

// Exact kNN through a uniform grid.  packed != 0: xyz [n_total,3] with offset[b]/new_offset[b] (segmentation
// semantics, global row ids); packed == 0: dense [b,n,3] / [b,m,3] (classification semantics, local ids).
// heap != 0 selects the heap-order kernels' tie semantics (rule R5) else the stable (d2, index) order (R4).
RSB_EXPORT int rsb_knnquery_grid(int packed, int heap, int b, int n, int m, int n_total, int m_total, int nsample,
                                 const float *xyz, const float *new_xyz, const int *offset, const int *new_offset,
                                 int *idx, float *dist, int sqrt_out, void *workspace, long workspace_bytes,
                                 cudaStream_t stream)
{
    RSB_REQUIRE(nsample >= 1 && nsample <= (heap ? 100 : 200), "nsample out of range");
    RSB_REQUIRE(workspace && workspace_bytes >= rsb_knn_grid_workspace_bytes(n_total, b), "workspace too small");
    if (b == 0 || m_total == 0) return 0;
    unsigned char *w = static_cast<unsigned char *>(workspace);
    GridParams G = {};
    G.xyz = xyz; G.offset = packed ? offset : nullptr; G.b = b; G.n = n;
    // Cell size by k: the search stops after the 27-cell neighbourhood when the k-th neighbour is closer than the nearest face of
    // that cube (>= one cell edge h away).  In a uniform cloud with p points per cell the k nearest lie within
    // h (3 k / (4 pi p))^(1/3): p ~ k / 2 puts them at ~0.8 h.  Fewer points per cell for small k mean fewer candidates
    // (k = 3: ~70 instead of ~216), more for k = 32 avoid the 125-cell second ring (~1000 candidates).  The result is exact
    // for any p; RSB_KNN_PPC overrides (tuning).
    static const float ppc_env = [] { const char *v = getenv("RSB_KNN_PPC"); return v ? (float)atof(v) : 0.f; }();
    G.target_per_cell = ppc_env > 0.f ? ppc_env : fminf(fmaxf(0.5f * (float)nsample, 2.5f), 24.f);
    G.seg = reinterpret_cast<SegGrid *>(w); w += align_up(sizeof(SegGrid) * (size_t)b, 256);
    G.bbox = reinterpret_cast<float *>(w); w += align_up(sizeof(float) * 6 * (size_t)b, 256);
    const size_t cells = (size_t)n_total + b + 1;
    G.cell_cnt = reinterpret_cast<int *>(w); w += align_up(sizeof(int) * cells, 256);
    G.cursor = reinterpret_cast<int *>(w); w += align_up(sizeof(int) * cells, 256);
    G.cid = reinterpret_cast<int *>(w); w += align_up(sizeof(int) * (size_t)n_total, 256);
    G.sorted = reinterpret_cast<float4 *>(w);
    RSB_CUDA(cudaMemsetAsync(G.cell_cnt, 0, 2 * align_up(sizeof(int) * cells, 256), stream));
    grid_setup_kernel<<<b, 1024, 0, stream>>>(G);
    RSB_CHECK_LAUNCH("grid_setup_kernel");
    const int eb = (int)((n_total + 255) / 256 < rsb_sm_count() * 8 ? (n_total + 255) / 256 : rsb_sm_count() * 8);
    grid_count_kernel<<<eb, 256, 0, stream>>>(G, n_total);
    RSB_CHECK_LAUNCH("grid_count_kernel");
    grid_scan_kernel<<<b, 1024, 0, stream>>>(G);
    RSB_CHECK_LAUNCH("grid_scan_kernel");
    grid_scatter_kernel<<<eb, 256, 0, stream>>>(G, n_total);
    RSB_CHECK_LAUNCH("grid_scatter_kernel");
    RSB_COUNT_LAUNCH(4);
    GridQuery Q = {};
    Q.xyz = xyz; Q.new_xyz = new_xyz; Q.new_offset = packed ? new_offset : nullptr; Q.seg = G.seg;
    Q.cell_start = G.cell_cnt; Q.sorted = G.sorted; Q.idx = idx; Q.dist2 = dist;
    Q.b = b; Q.m_dense = m; Q.m_total = m_total; Q.k = nsample; Q.packed = packed; Q.sqrt_out = sqrt_out;
    Q.counters = g_knn_counters;
    return heap ? launch_query<true>(Q, stream) : launch_query<false>(Q, stream);
}
