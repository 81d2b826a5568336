// fps.cu — farthest point sampling for sm_100a: one thread-block CLUSTER per cloud / segment,
// the cloud resident in registers for the whole run, one cluster barrier per sample.
//
// Replaces (same results, index-exact):
//   classification/modules/pointops/src/sampling/sampling_cuda_kernel.cu:59-210  (dense  [B,N,3])
//   segmentation/modules/pointops/src/sampling/sampling_cuda_kernel.cu:15-171    (packed [ΣN,3]+offset)
//
// Design (B200-first, not a translation):
//   * The reference runs ONE block per cloud, re-reads xyz (stride-3 AoS) and read-modify-writes the
//     running min-distance array in global memory on every one of the m iterations, then does an
//     11-barrier shared-memory tree.  Here each thread owns PPT points (x,y,z,min-dist) in REGISTERS
//     for the whole kernel; HBM/L2 is touched once at load time.  A cluster of CS CTAs (up to 16 =
//     the non-portable maximum) splits one cloud, so a 40 960-point cloud keeps 10 points/thread.
//   * Per iteration: register scan -> warp arg-max with redux.sync -> every warp leader publishes
//     (key, x, y, z) of its winner straight into the shared memory of EVERY CTA of the cluster
//     (st.shared::cluster), ONE barrier.cluster (which also orders the block), then every warp
//     reduces the CS*warps candidates locally.  No second exchange for the winner's coordinates.
//   * Exact tie rule R2: the reference's winner among equal maxima is the point
//     whose (bit-reversed (k mod BS), k) is smallest, BS = its block size.  Points are therefore
//     loaded in that PRIORITY ORDER (position p <-> original index k), so "smallest position wins"
//     reproduces the reference bit for bit with a plain (value, ~position) max key.
//   * Distances use rsb_sqdist (rule R1).
#include "common.cuh"

namespace {

struct FpsParams {
    const float *xyz;        // dense: [b, n, 3]; packed: [sum n, 3]
    const int *offset;       // packed only: cumulative ends [b]
    const int *new_offset;   // packed only
    float *temp;             // optional running-min scratch (reference ABI); final values written back
    int *idx;                // dense: [b, m]; packed: [sum m]
    float *new_xyz;          // optional fused gather of the sampled coordinates (same layout as idx, x3)
    int n, m;                // dense sizes (unused when packed)
    int packed;
    int bs_ref;              // the REFERENCE's block size (power of two) -> tie rule
    int log2_bs;
    const int *n_max_dev;    // optional: device scalar holding the largest segment length; when set the
                             // reference block size is derived from it in-kernel (no host sync needed)
    int cs;                  // cluster size
    int pos_lo, pos_hi;      // only segments whose priority-position count lies in (pos_lo, pos_hi] are processed
};

__device__ __forceinline__ int pos2k(uint32_t p, int q_per_thread, int bs_ref, int log2_bs, bool &in_range)
{
    // position p = r * Q + q  <->  reference thread t = bitrev(r), q-th point of that thread: k = q*BS + t
    const uint32_t r = p / (uint32_t)q_per_thread;
    const uint32_t q = p - r * (uint32_t)q_per_thread;
    const uint32_t t = log2_bs ? (__brev(r) >> (32 - log2_bs)) : 0u;
    in_range = r < (uint32_t)bs_ref;
    return (int)(q * (uint32_t)bs_ref + t);
}

// PPT > 0: points in registers.  PPT == 0: streaming fallback (min-dist in P.temp, xyz re-read through L2).
// WIDE: 1024-thread CTAs (64 registers per thread: PPT <= 12) - lets ONE CTA hold a 12k-point segment and skip the
// cluster barrier.  Measured SLOWER than a 4-CTA cluster at 10k points (1.31 vs 0.95 us/sample: one SM's issue rate
// bounds the 13-instruction/point scan), so no default plan selects it; reachable through RSB_FPS_PLAN="1,1024".
template <int PPT, bool CLUSTER, bool WIDE = false>
__global__ void __launch_bounds__((PPT == 0 || WIDE) ? 1024 : 512, 1) fps_kernel(FpsParams P)
{
    extern __shared__ __align__(16) unsigned char fps_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const uint32_t cs = CLUSTER ? (uint32_t)P.cs : 1u;
    const uint32_t rank = CLUSTER ? rsb_cluster_ctarank() : 0u;
    const int seg = CLUSTER ? (int)rsb_cluster_id_x() : (int)blockIdx.x;
    const int E = (int)cs * nwarps;  // candidates per iteration
    if (P.n_max_dev) {  // floor(log2(n_max)) capped at 10 == the reference's libm expression (checked for n < 5000)
        const int nm = max(__ldg(P.n_max_dev), 1);
        P.log2_bs = min(31 - __clz(nm), 10);
        P.bs_ref = 1 << P.log2_bs;
    }
    uint2 *keys = reinterpret_cast<uint2 *>(fps_smem);                       // [2][E]
    float4 *crd = reinterpret_cast<float4 *>(fps_smem + sizeof(uint2) * 2 * E);  // [2][E]

    // ---- segment descriptor ------------------------------------------------------------------
    int start_n, n_seg, start_m, m_seg, index_base;
    if (P.packed) {
        start_n = seg ? P.offset[seg - 1] : 0;
        n_seg = P.offset[seg] - start_n;
        start_m = seg ? P.new_offset[seg - 1] : 0;
        m_seg = P.new_offset[seg] - start_m;
        index_base = start_n;
    } else {
        start_n = seg * P.n;
        n_seg = P.n;
        start_m = seg * P.m;
        m_seg = P.m;
        index_base = 0;
    }
    const float *xyz = P.xyz + (size_t)start_n * 3;
    int *out = P.idx + start_m;
    float *out_xyz = P.new_xyz ? P.new_xyz + (size_t)start_m * 3 : nullptr;
    float *temp = P.temp ? P.temp + start_n : nullptr;
    if (n_seg <= 0 || m_seg <= 0) return;  // whole cluster takes this branch together

    const uint32_t T = cs * blockDim.x;
    const uint32_t g = rank * blockDim.x + tid;
    const int Q = (n_seg + P.bs_ref - 1) / P.bs_ref;
    const uint32_t npos = (uint32_t)Q * (uint32_t)P.bs_ref;
    // bounded launches (rsb_furthestsampling_packed_bounded): a segment belongs to exactly one of the two launches
    if ((long)npos <= (long)P.pos_lo || (long)npos > (long)P.pos_hi) return;   // whole cluster takes this branch together

    // ---- load (priority order) -----------------------------------------------------------------
    constexpr int R = PPT > 0 ? PPT : 1;
    float px[R], py[R], pz[R], md[R];
    if (PPT > 0) {
#pragma unroll
        for (int j = 0; j < R; j++) {
            const uint32_t p = g + (uint32_t)j * T;
            bool in_range;
            const int k = pos2k(p, Q, P.bs_ref, P.log2_bs, in_range);
            const bool valid = in_range && k < n_seg;
            px[j] = valid ? __ldg(xyz + (size_t)k * 3 + 0) : 0.f;
            py[j] = valid ? __ldg(xyz + (size_t)k * 3 + 1) : 0.f;
            pz[j] = valid ? __ldg(xyz + (size_t)k * 3 + 2) : 0.f;
            md[j] = valid ? 1e10f : -1.f;  // fminf(d, -1) stays -1: never beats a real point
        }
    } else {
        for (uint32_t p = g; p < npos; p += T) {
            bool in_range;
            const int k = pos2k(p, Q, P.bs_ref, P.log2_bs, in_range);
            if (in_range && k < n_seg) temp[k] = 1e10f;
        }
        if (CLUSTER) { rsb_cluster_arrive_release(); rsb_cluster_wait_acquire(); } else __syncthreads();
    }

    float cx = __ldg(xyz + 0), cy = __ldg(xyz + 1), cz = __ldg(xyz + 2);  // first sample: row 0 (k = 0)
    if (g == 0) {
        out[0] = index_base;
        if (out_xyz) { out_xyz[0] = cx; out_xyz[1] = cy; out_xyz[2] = cz; }
    }

    for (int it = 1; it < m_seg; it++) {
        // ---- thread-local scan: first strict maximum in position order ---------------------------
        float best = -1.f, bx = 0.f, by = 0.f, bz = 0.f;
        uint32_t bp = g;
        if (PPT > 0) {
            int bj = 0;
#pragma unroll
            for (int j = 0; j < R; j++) {
                const float d = rsb_sqdist(px[j], py[j], pz[j], cx, cy, cz);
                const float d2 = fminf(d, md[j]);
                md[j] = d2;
                const bool gt = d2 > best;
                best = gt ? d2 : best;
                bj = gt ? j : bj;
                bx = gt ? px[j] : bx;
                by = gt ? py[j] : by;
                bz = gt ? pz[j] : bz;
            }
            bp = g + (uint32_t)bj * T;
        } else {
            for (uint32_t p = g; p < npos; p += T) {
                bool in_range;
                const int k = pos2k(p, Q, P.bs_ref, P.log2_bs, in_range);
                if (!(in_range && k < n_seg)) continue;
                const float x = xyz[(size_t)k * 3], y = xyz[(size_t)k * 3 + 1], z = xyz[(size_t)k * 3 + 2];
                const float d2 = fminf(rsb_sqdist(x, y, z, cx, cy, cz), temp[k]);
                temp[k] = d2;
                if (d2 > best) { best = d2; bp = p; bx = x; by = y; bz = z; }
            }
        }
        // key: value bits (+1 so that "no point" = 0 sorts below d2 = +0), then ~position (smaller wins)
        const uint32_t hi = best >= 0.f ? __float_as_uint(best) + 1u : 0u;
        const uint32_t lo = ~bp;
        // ---- warp arg-max ------------------------------------------------------------------------
        const uint32_t whi = __reduce_max_sync(0xffffffffu, hi);
        const uint32_t wlo = __reduce_max_sync(0xffffffffu, hi == whi ? lo : 0u);
        const int src = __ffs(__ballot_sync(0xffffffffu, hi == whi && lo == wlo)) - 1;
        const float wx = __shfl_sync(0xffffffffu, bx, src);
        const float wy = __shfl_sync(0xffffffffu, by, src);
        const float wz = __shfl_sync(0xffffffffu, bz, src);
        // ---- publish to every CTA of the cluster, one barrier --------------------------------------
        const int par = it & 1;
        const int slot = par * E + (int)rank * nwarps + warp;
        if (CLUSTER) {
            if ((uint32_t)lane < cs) {
                rsb_st_cluster_v2(rsb_mapa(rsb_smem_addr(&keys[slot]), lane), whi, wlo);
                rsb_st_cluster_v4(rsb_mapa(rsb_smem_addr(&crd[slot]), lane), __float_as_uint(wx),
                                  __float_as_uint(wy), __float_as_uint(wz), 0u);
            }
            rsb_cluster_arrive_release();
            rsb_cluster_wait_acquire();
        } else {
            if (lane == 0) {
                keys[slot] = make_uint2(whi, wlo);
                crd[slot] = make_float4(wx, wy, wz, 0.f);
            }
            __syncthreads();
        }
        // ---- every warp reduces the E candidates ---------------------------------------------------
        uint32_t khi = 0, klo = 0;
        int ke = 0;
        for (int e = lane; e < E; e += 32) {
            const uint2 k = keys[par * E + e];
            const bool gt = k.x > khi || (k.x == khi && k.y > klo);
            khi = gt ? k.x : khi;
            klo = gt ? k.y : klo;
            ke = gt ? e : ke;
        }
        const uint32_t fhi = __reduce_max_sync(0xffffffffu, khi);
        const uint32_t flo = __reduce_max_sync(0xffffffffu, khi == fhi ? klo : 0u);
        const int wsrc = __ffs(__ballot_sync(0xffffffffu, khi == fhi && klo == flo)) - 1;
        const int we = __shfl_sync(0xffffffffu, ke, wsrc);
        const float4 c = crd[par * E + we];
        cx = c.x; cy = c.y; cz = c.z;
        if (g == 0) {
            bool in_range;
            const int k = pos2k(~flo, Q, P.bs_ref, P.log2_bs, in_range);
            out[it] = index_base + k;
            if (out_xyz) { out_xyz[it * 3] = cx; out_xyz[it * 3 + 1] = cy; out_xyz[it * 3 + 2] = cz; }
        }
    }

    // reference ABI: `temp` holds the final running minima (callers discard it)
    if (PPT > 0 && temp) {
#pragma unroll
        for (int j = 0; j < R; j++) {
            const uint32_t p = g + (uint32_t)j * T;
            bool in_range;
            const int k = pos2k(p, Q, P.bs_ref, P.log2_bs, in_range);
            if (in_range && k < n_seg) temp[k] = md[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Second-generation cluster kernel: the per-sample exchange is a message, not a barrier.
//   * round 1 measured one barrier.cluster per sample at ~0.45 us of the 0.95 us (profiles/r01_fps_plan_sweep.txt).
//     Here every warp's winner (key, x, y, z) goes to every CTA of the cluster with st.async, which completes bytes on
//     the DESTINATION's mbarrier; a CTA waits on its own mbarrier (hardware-suspended try_wait) until the E messages of
//     the sample have landed.  No cluster barrier and no __syncthreads inside the loop: two mbarriers alternate by
//     sample parity (a warp can only be one sample ahead of the slowest warp of the cluster, because finishing sample
//     s needs everybody's message of sample s).
//   * the scan keeps only (value, local index) per thread - the coordinates of the winner are read back from a shared
//     memory copy of the CTA's points by the one lane that won - and evaluates two points per instruction on the
//     packed fp32x2 pipe (__fadd2_rn / __fmul2_rn / __ffma2_rn are exact per element, so rule R1 holds): 7 instead
//     of 13 instructions per point.
// Same priority-order layout, key and tie rule as fps_kernel: results are bit-identical.
__device__ __forceinline__ void fps_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(0x989680u)
            : "memory");
    } while (!done);
}

template <int PP>   // point PAIRS per thread
__global__ void __launch_bounds__(512, 1) fps2_kernel(FpsParams P)
{
    extern __shared__ __align__(16) unsigned char fps_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nt = blockDim.x, nwarps = nt >> 5;
    const uint32_t cs = (uint32_t)P.cs;
    const uint32_t rank = rsb_cluster_ctarank();
    const int seg = (int)rsb_cluster_id_x();
    const int E = (int)cs * nwarps;
    if (P.n_max_dev) {
        const int nm = max(__ldg(P.n_max_dev), 1);
        P.log2_bs = min(31 - __clz(nm), 10);
        P.bs_ref = 1 << P.log2_bs;
    }
    uint64_t *bars = reinterpret_cast<uint64_t *>(fps_smem);                           // [2]
    uint4 *msg = reinterpret_cast<uint4 *>(fps_smem + 16);                              // [2][E]  (hi, lo, x, y)
    float *msgz = reinterpret_cast<float *>(fps_smem + 16 + sizeof(uint4) * 2 * E);     // [2][E]
    float4 *pts = reinterpret_cast<float4 *>(fps_smem + 16 + (sizeof(uint4) + 4) * 2 * E + ((16 - (8 * E) % 16) % 16));

    int start_n, n_seg, start_m, m_seg, index_base;
    if (P.packed) {
        start_n = seg ? P.offset[seg - 1] : 0;
        n_seg = P.offset[seg] - start_n;
        start_m = seg ? P.new_offset[seg - 1] : 0;
        m_seg = P.new_offset[seg] - start_m;
        index_base = start_n;
    } else {
        start_n = seg * P.n; n_seg = P.n; start_m = seg * P.m; m_seg = P.m; index_base = 0;
    }
    const float *xyz = P.xyz + (size_t)start_n * 3;
    int *out = P.idx + start_m;
    float *out_xyz = P.new_xyz ? P.new_xyz + (size_t)start_m * 3 : nullptr;
    float *temp = P.temp ? P.temp + start_n : nullptr;
    if (n_seg <= 0 || m_seg <= 0) return;
    const uint32_t T = cs * (uint32_t)nt;
    const uint32_t g = rank * (uint32_t)nt + tid;
    const int Q = (n_seg + P.bs_ref - 1) / P.bs_ref;
    const uint32_t npos = (uint32_t)Q * (uint32_t)P.bs_ref;
    if ((long)npos <= (long)P.pos_lo || (long)npos > (long)P.pos_hi) return;   // whole cluster takes this branch together

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rsb_smem_addr(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rsb_smem_addr(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- load (priority order): registers hold pairs (2i, 2i+1) -> positions g + (2i)T, g + (2i+1)T ----------
    float2 px[PP], py[PP], pz[PP], md[PP];
#pragma unroll
    for (int i = 0; i < PP; i++) {
        float v[2][4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t p = g + (uint32_t)(2 * i + h) * T;
            bool in_range;
            const int k = pos2k(p, Q, P.bs_ref, P.log2_bs, in_range);
            const bool valid = in_range && k < n_seg;
            v[h][0] = valid ? __ldg(xyz + (size_t)k * 3 + 0) : 0.f;
            v[h][1] = valid ? __ldg(xyz + (size_t)k * 3 + 1) : 0.f;
            v[h][2] = valid ? __ldg(xyz + (size_t)k * 3 + 2) : 0.f;
            v[h][3] = valid ? 1e10f : -1.f;
            pts[(2 * i + h) * nt + tid] = make_float4(v[h][0], v[h][1], v[h][2], 0.f);
        }
        px[i] = make_float2(v[0][0], v[1][0]);
        py[i] = make_float2(v[0][1], v[1][1]);
        pz[i] = make_float2(v[0][2], v[1][2]);
        md[i] = make_float2(v[0][3], v[1][3]);
    }
    float cx = __ldg(xyz + 0), cy = __ldg(xyz + 1), cz = __ldg(xyz + 2);
    if (g == 0) {
        out[0] = index_base;
        if (out_xyz) { out_xyz[0] = cx; out_xyz[1] = cy; out_xyz[2] = cz; }
    }
    // mbarrier initialisation visible to the whole cluster before the first remote completion
    rsb_cluster_arrive_release();
    rsb_cluster_wait_acquire();

    const uint32_t bar0 = rsb_smem_addr(&bars[0]);
    uint32_t phases = 0u;                       // bit p: parity to wait for on mbarrier p
    const uint32_t tx_bytes = (uint32_t)E * 20u;
    for (int it = 1; it < m_seg; it++) {
        const int par = it & 1;
        const uint32_t bar = bar0 + 8u * (uint32_t)par;
        if (tid == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx_bytes) : "memory");
        // ---- thread-local scan: first strict maximum in position order (two points per packed instruction) ----
        const float2 ncx = make_float2(-cx, -cx), ncy = make_float2(-cy, -cy), ncz = make_float2(-cz, -cz);
        float best = -1.f;
        int bj = 0;
#pragma unroll
        for (int i = 0; i < PP; i++) {
            const float2 dx = __fadd2_rn(px[i], ncx), dy = __fadd2_rn(py[i], ncy), dz = __fadd2_rn(pz[i], ncz);
            float2 t = __fmul2_rn(dy, dy);
            t = __ffma2_rn(dx, dx, t);
            t = __ffma2_rn(dz, dz, t);
            const float m0 = fminf(t.x, md[i].x), m1 = fminf(t.y, md[i].y);
            md[i] = make_float2(m0, m1);
            const bool g0 = m0 > best;
            best = g0 ? m0 : best;
            bj = g0 ? 2 * i : bj;
            const bool g1 = m1 > best;
            best = g1 ? m1 : best;
            bj = g1 ? 2 * i + 1 : bj;
        }
        const uint32_t bp = g + (uint32_t)bj * T;
        const uint32_t hi = best >= 0.f ? __float_as_uint(best) + 1u : 0u;
        const uint32_t lo = ~bp;
        const uint32_t whi = __reduce_max_sync(0xffffffffu, hi);
        const uint32_t wlo = __reduce_max_sync(0xffffffffu, hi == whi ? lo : 0u);
        const int src = __ffs(__ballot_sync(0xffffffffu, hi == whi && lo == wlo)) - 1;
        // ---- the winning lane sends (key, x, y, z) to every CTA of the cluster -------------------------------
        if (lane == src) {
            const float4 c = pts[bj * nt + tid];
            const int slot = par * E + (int)rank * nwarps + warp;
            const uint32_t a_msg = rsb_smem_addr(&msg[slot]), a_z = rsb_smem_addr(&msgz[slot]);
            for (uint32_t r = 0; r < cs; r++) {
                const uint32_t rbar = rsb_mapa(bar, r);
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(rsb_mapa(a_msg, r)),
                             "r"(whi), "r"(wlo), "r"(__float_as_uint(c.x)), "r"(__float_as_uint(c.y)), "r"(rbar)
                             : "memory");
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(rsb_mapa(a_z, r)),
                             "r"(__float_as_uint(c.z)), "r"(rbar)
                             : "memory");
            }
        }
        // ---- wait for the E messages of this sample, reduce them ----------------------------------------------
        fps_mbar_wait(bar, (phases >> par) & 1u);
        phases ^= 1u << par;
        uint32_t khi = 0, klo = 0;
        float kx = 0.f, ky = 0.f, kz = 0.f;
        for (int e = lane; e < E; e += 32) {
            const uint4 k = msg[par * E + e];
            const bool gt = k.x > khi || (k.x == khi && k.y > klo);
            if (gt) { khi = k.x; klo = k.y; kx = __uint_as_float(k.z); ky = __uint_as_float(k.w); kz = msgz[par * E + e]; }
        }
        const uint32_t fhi = __reduce_max_sync(0xffffffffu, khi);
        const uint32_t flo = __reduce_max_sync(0xffffffffu, khi == fhi ? klo : 0u);
        const int wsrc = __ffs(__ballot_sync(0xffffffffu, khi == fhi && klo == flo)) - 1;
        cx = __shfl_sync(0xffffffffu, kx, wsrc);
        cy = __shfl_sync(0xffffffffu, ky, wsrc);
        cz = __shfl_sync(0xffffffffu, kz, wsrc);
        if (g == 0) {
            bool in_range;
            const int k = pos2k(~flo, Q, P.bs_ref, P.log2_bs, in_range);
            out[it] = index_base + k;
            if (out_xyz) { out_xyz[it * 3] = cx; out_xyz[it * 3 + 1] = cy; out_xyz[it * 3 + 2] = cz; }
        }
    }
    if (temp) {
#pragma unroll
        for (int i = 0; i < PP; i++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t p = g + (uint32_t)(2 * i + h) * T;
                bool in_range;
                const int k = pos2k(p, Q, P.bs_ref, P.log2_bs, in_range);
                if (in_range && k < n_seg) temp[k] = h ? md[i].y : md[i].x;
            }
        }
    }
    // no CTA leaves while a peer may still have messages for it in flight
    rsb_cluster_arrive_release();
    rsb_cluster_wait_acquire();
}

struct FpsPlan {
    int cs, nt, ppt;
};

// Choose (cluster size, threads, points/thread).  `n_max` here is the number of PRIORITY POSITIONS the
// largest segment spans (BS * ceil(n / BS) >= n: the position space is padded to whole reference threads).
FpsPlan fps_plan(int n_max)
{
    FpsPlan pl;
    if (n_max <= 256) { pl.cs = 1; pl.nt = 64; }
    else if (n_max <= 1024) { pl.cs = 1; pl.nt = 128; }
    else if (n_max <= 4096) { pl.cs = 1; pl.nt = 256; }
    else if (n_max <= 8192) { pl.cs = 1; pl.nt = 512; }     // one CTA: ~0.55 us/sample; any cluster costs >= 0.9 us
    else if (n_max <= 16384) { pl.cs = 4; pl.nt = 256; }    // measured sweep: profiles/r01_fps_plan_sweep.txt
    else if (n_max <= 65536) { pl.cs = 8; pl.nt = 512; }
    else { pl.cs = 16; pl.nt = 512; }
    const long cap = (long)pl.cs * pl.nt;
    pl.ppt = (int)((n_max + cap - 1) / cap);
    if (pl.ppt > 16) { pl.cs = 16; pl.nt = 1024; pl.ppt = 0; }  // streaming fallback
    return pl;
}

template <int PPT>
int fps_launch_ppt(const FpsParams &P, int nseg, const FpsPlan &pl, cudaStream_t stream)
{
    const size_t smem = (size_t)(sizeof(uint2) + sizeof(float4)) * 2 * pl.cs * (pl.nt / 32);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(pl.nt);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (pl.cs > 1) {
        auto kern = fps_kernel<PPT, true>;
        if (pl.cs > 8) RSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cfg.gridDim = dim3(nseg * pl.cs);
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = pl.cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        RSB_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
    } else {
        cfg.gridDim = dim3(nseg);
        cfg.attrs = nullptr;
        cfg.numAttrs = 0;
        if (pl.nt > 512) {
            if constexpr (PPT >= 1 && PPT <= 12) RSB_CUDA(cudaLaunchKernelEx(&cfg, fps_kernel<PPT, false, true>, P));
            else { rsb_set_error("fps: 1024-thread CTAs hold at most 12 points per thread"); return (int)cudaErrorInvalidValue; }
        } else {
            RSB_CUDA(cudaLaunchKernelEx(&cfg, fps_kernel<PPT, false>, P));
        }
    }
    RSB_COUNT_LAUNCH(1);
    return 0;
}

// positions per thread the plan provides times its threads: the largest position count it can hold in registers
long fps_capacity(const FpsPlan &pl) { return pl.ppt > 0 ? (long)pl.cs * pl.nt * pl.ppt : 0x7fffffffL; }

int g_fps_generation = [] { const char *v = getenv("RSB_FPS_V1"); return (v && v[0] && v[0] != '0') ? 1 : 0; }();

template <int PP>
int fps2_launch_pp(const FpsParams &P, int nseg, const FpsPlan &pl, cudaStream_t stream)
{
    const int E = pl.cs * (pl.nt / 32);
    const size_t smem = 16 + (sizeof(uint4) + 4) * 2 * E + 16 + (size_t)2 * PP * pl.nt * sizeof(float4);
    auto kern = fps2_kernel<PP>;
    static bool attr_set = false;
    if (!attr_set) {
        RSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        RSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(pl.nt);
    cfg.gridDim = dim3(nseg * pl.cs);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl.cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RSB_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
    RSB_COUNT_LAUNCH(1);
    return 0;
}

int fps_launch(FpsParams P, int nseg, int n_max_pts, cudaStream_t stream, long *capacity = nullptr)
{
    // positions spanned by the largest segment; with a device-side n_max the reference block size may be
    // smaller than the host's bound, so allow one extra block of padding
    int n_max = RSB_DIVUP(n_max_pts, P.bs_ref) * P.bs_ref;
    if (P.n_max_dev) n_max = n_max_pts + 1024;
    FpsPlan pl = fps_plan(n_max);
    if (const char *e = getenv("RSB_FPS_PLAN")) {  // tuning hook: "cs,nt"
        int cs = 0, nt = 0;
        if (sscanf(e, "%d,%d", &cs, &nt) == 2 && cs >= 1 && cs <= 16 && nt >= 32 && nt <= 1024 && nt % 32 == 0 && (nt <= 512 || cs == 1)) {
            long cap = (long)cs * nt;
            int ppt = (int)((n_max + cap - 1) / cap);
            if (ppt <= (nt > 512 ? 12 : 16)) { pl.cs = cs; pl.nt = nt; pl.ppt = ppt; }
        }
    }
    P.cs = pl.cs;
    if (capacity) {                      // bounded launch: take what fits this plan, leave the rest to the follow-up launch
        *capacity = fps_capacity(pl);
        P.pos_hi = (int)(*capacity < 0x7fffffffL ? *capacity : 0x7fffffffL);
    }
    if (pl.ppt == 0) RSB_REQUIRE(P.temp != nullptr, "segments beyond the register-resident capacity need the temp scratch buffer");
    if (pl.cs > 1 && pl.ppt > 0 && !g_fps_generation) {
        // cluster plans: message-passing kernel (capacity is unchanged: 2 * ceil(ppt / 2) >= ppt points per thread)
        switch ((pl.ppt + 1) / 2) {
#define RSB_CASE2(N) case N: return fps2_launch_pp<N>(P, nseg, pl, stream);
            RSB_CASE2(1) RSB_CASE2(2) RSB_CASE2(3) RSB_CASE2(4) RSB_CASE2(5) RSB_CASE2(6) RSB_CASE2(7) RSB_CASE2(8)
#undef RSB_CASE2
        }
    }
    switch (pl.ppt) {
#define RSB_CASE(N) case N: return fps_launch_ppt<N>(P, nseg, pl, stream);
        RSB_CASE(0) RSB_CASE(1) RSB_CASE(2) RSB_CASE(3) RSB_CASE(4) RSB_CASE(5) RSB_CASE(6) RSB_CASE(7) RSB_CASE(8)
        RSB_CASE(9) RSB_CASE(10) RSB_CASE(11) RSB_CASE(12) RSB_CASE(13) RSB_CASE(14) RSB_CASE(15) RSB_CASE(16)
#undef RSB_CASE
    }
    rsb_set_error("fps: no plan for n_max=%d", n_max);
    return (int)cudaErrorInvalidValue;
}

// classification/modules/pointops/src/cuda_utils.h:15-18 (same libm expression => same block size)
int ref_opt_n_threads(int work_size)
{
    const int pow_2 = (int)(std::log(static_cast<double>(work_size)) / std::log(2.0));
    int v = 1 << pow_2;
    if (v > 1024) v = 1024;
    if (v < 1) v = 1;
    return v;
}

int ilog2(int v)
{
    int l = 0;
    while ((1 << (l + 1)) <= v) l++;
    return l;
}

}  // namespace

RSB_EXPORT void rsb_fps_set_generation(int gen) { g_fps_generation = gen == 1 ? 1 : 0; }

RSB_EXPORT int rsb_furthestsampling_dense(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                          float *new_xyz, cudaStream_t stream)
{
    RSB_REQUIRE(b >= 0 && n >= 1 && m >= 0, "bad sizes");
    if (b == 0 || m == 0) return 0;
    FpsParams P = {};
    P.xyz = xyz; P.temp = temp; P.idx = idx; P.new_xyz = new_xyz;
    P.n = n; P.m = m; P.packed = 0;
    P.pos_lo = -1; P.pos_hi = 0x7fffffff;
    P.bs_ref = ref_opt_n_threads(n);
    P.log2_bs = ilog2(P.bs_ref);
    return fps_launch(P, b, n, stream);
}

RSB_EXPORT int rsb_furthestsampling_packed(int b, int n_max, const int *n_max_dev, const float *xyz,
                                           const int *offset, const int *new_offset, float *tmp, int *idx,
                                           float *new_xyz, cudaStream_t stream)
{
    RSB_REQUIRE(b >= 0 && n_max >= 1, "bad sizes");
    if (b == 0) return 0;
    FpsParams P = {};
    P.xyz = xyz; P.offset = offset; P.new_offset = new_offset; P.temp = tmp; P.idx = idx; P.new_xyz = new_xyz;
    P.packed = 1;
    P.pos_lo = -1; P.pos_hi = 0x7fffffff;
    P.n_max_dev = n_max_dev;
    P.bs_ref = ref_opt_n_threads(n_max);
    P.log2_bs = ilog2(P.bs_ref);
    return fps_launch(P, b, n_max, stream);
}

// Packed FPS when the host knows only BOUNDS on the segment sizes (sectorized FPS: the sector sizes are computed on the
// device).  Two launches, no host synchronisation: the first is planned for n_expect points per segment and processes
// every segment that fits its register capacity; the second is planned for n_limit (>= every segment) and processes
// only the segments the first one left.  With well-balanced sectors the second launch exits immediately.
RSB_EXPORT int rsb_furthestsampling_packed_bounded(int b, int n_expect, int n_limit, const int *n_max_dev, const float *xyz,
                                                   const int *offset, const int *new_offset, float *tmp, int *idx,
                                                   float *new_xyz, cudaStream_t stream)
{
    RSB_REQUIRE(b >= 0 && n_expect >= 1 && n_limit >= 1, "bad sizes");
    if (b == 0) return 0;
    if (n_expect > n_limit) n_expect = n_limit;
    FpsParams P = {};
    P.xyz = xyz; P.offset = offset; P.new_offset = new_offset; P.temp = tmp; P.idx = idx; P.new_xyz = new_xyz;
    P.packed = 1;
    P.n_max_dev = n_max_dev;
    // reference block size: from the device-side maximum when given, else from n_expect (the caller guarantees that the
    // true maximum and n_expect are both >= 1024, where the reference caps the block size)
    P.bs_ref = ref_opt_n_threads(n_expect);
    P.log2_bs = ilog2(P.bs_ref);
    P.pos_lo = -1; P.pos_hi = 0x7fffffff;
    long cap = 0;
    int rc = fps_launch(P, b, n_expect, stream, &cap);
    if (rc) return rc;
    if (cap >= (long)n_limit + 1024) return 0;      // the first plan already holds the largest possible segment
    P.pos_lo = (int)cap; P.pos_hi = 0x7fffffff;
    return fps_launch(P, b, n_limit, stream);
}
