// umbrella.cu — umbrella-surface descriptors of RepSurf-U in ONE kernel (one thread per point).
//
// Replaces the ~25 small torch kernels the reference launches per forward for
//   {classification,segmentation}/modules/repsurface_utils.py  group_by_umbrella[_v2]  (argsort by azimuth, roll, cat)
//   {classification,segmentation}/modules/recons_utils.py       cal_normal, cal_center, cal_const, check_nan_umb
//   {classification,segmentation}/modules/polar_utils.py        xyz2sphere
// Input: point coordinates + the kNN index list of every point (global row ids); output: the 10-channel
// descriptor of every triangle of the point's umbrella, [points, G, 10], channel order of the chosen tree:
//   cls: [centroid(3), polar(3), normal(3), pos(1)]   classification/modules/repsurface_utils.py:290
//   seg: [polar(3), normal(3), pos(1), centroid(3)]   segmentation/modules/repsurface_utils.py:320
// Semantics kept on purpose (SURVEY.md §8a note †): the segmentation tree keeps the query itself among the k
// neighbours (zero offset -> two degenerate triangles -> NaN normals -> repaired from the first valid triangle),
// sorts by the azimuth of the ROTATED offsets but builds triangles from the unrotated ones; the sign of all
// normals follows the x component of triangle 0 (NaN counts as "not positive"); the polar form is computed before
// the NaN repair.  Forward only: coordinates carry no gradient on the RepSurf path.
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int MAXG = 32;
constexpr float INV_2PI = 0.15915494309189535f;   // 1 / (2*pi) as torch computes phi / (2*np.pi) in fp32
constexpr float PI_F = 3.14159265358979323846f;
constexpr float TWO_PI_F = 6.28318530717958647692f;

__device__ __forceinline__ float azimuth01(float y, float x) { return atan2f(y, x) / TWO_PI_F + 0.5f; }

__global__ void __launch_bounds__(128) umbrella_kernel(long np, int k, int skip_first, int rotate, int order_seg,
                                                        const float *__restrict__ xyz, const int *__restrict__ idx,
                                                        const float *__restrict__ flip, float *__restrict__ out, int channels,
                                                        int ld)
{
    const long p = blockIdx.x * 128L + threadIdx.x;
    if (p >= np) return;
    const int G = k - (skip_first ? 1 : 0);
    const float cx = xyz[p * 3], cy = xyz[p * 3 + 1], cz = xyz[p * 3 + 2];
    float ox[MAXG], oy[MAXG], oz[MAXG], key[MAXG];
    int ord[MAXG];
    for (int i = 0; i < G; i++) {
        const long j = idx[p * k + i + (skip_first ? 1 : 0)];
        ox[i] = xyz[j * 3] - cx;
        oy[i] = xyz[j * 3 + 1] - cy;
        oz[i] = xyz[j * 3 + 2] - cz;
        float kx = ox[i], ky = oy[i];
        if (rotate) {   // offsets @ [[0.5,-0.5,0.7071],[0.7071,0.7071,0],[-0.5,0.5,0.7071]]  (segmentation 'fix' sort)
            kx = ox[i] * 0.5f + oy[i] * 0.7071f + oz[i] * -0.5f;
            ky = ox[i] * -0.5f + oy[i] * 0.7071f + oz[i] * 0.5f;
        }
        key[i] = azimuth01(ky, kx);
        ord[i] = i;
    }
    // stable insertion sort of the neighbour order by azimuth
    for (int i = 1; i < G; i++) {
        const int oi = ord[i];
        const float ki = key[oi];
        int j = i - 1;
        while (j >= 0 && key[ord[j]] > ki) { ord[j + 1] = ord[j]; j--; }
        ord[j + 1] = oi;
    }
    const float fl = flip ? flip[p] : 1.f;
    float *o = out + (size_t)p * G * ld;      // triangle rows of pitch ld (>= 10)
    float sign = 0.f;
    int first_ok = -1;
    // pass 1: raw descriptors; remember the first triangle whose normal is finite
    for (int i = 0; i < G; i++) {
        const int a = ord[i], b = ord[(i + 1 == G) ? 0 : i + 1];
        const float ax = ox[a], ay = oy[a], az = oz[a], bx = ox[b], by = oy[b], bz = oz[b];
        float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
        const float nrm = sqrtf(nx * nx + ny * ny + nz * nz);
        nx /= nrm; ny /= nrm; nz /= nrm;
        if (i == 0) sign = (nx > 0.f ? 1.f : -1.f) * fl;
        nx *= sign; ny *= sign; nz *= sign;
        const float mx = (0.f + ax + bx) / 3.f, my = (0.f + ay + by) / 3.f, mz = (0.f + az + bz) / 3.f;
        const float rho = sqrtf(mx * mx + my * my + mz * mz);
        float theta = acosf(mz / rho);
        if (rho == 0.f) theta = 0.f;
        theta /= PI_F;
        const float phi = azimuth01(my, mx);
        const float pos = (nx * mx + ny * my + nz * mz) / 1.7320508075688772f;
        const bool bad = isnan(nx) || isnan(ny) || isnan(nz);
        if (!bad && first_ok < 0) first_ok = i;
        float *t = o + i * ld;
        if (order_seg) { t[0] = rho; t[1] = theta; t[2] = phi; t[3] = nx; t[4] = ny; t[5] = nz; t[6] = pos; t[7] = mx; t[8] = my; t[9] = mz; }
        else { t[0] = mx; t[1] = my; t[2] = mz; t[3] = rho; t[4] = theta; t[5] = phi; t[6] = nx; t[7] = ny; t[8] = nz; t[9] = pos; }
    }
    if (first_ok < 0) first_ok = 0;   // argmax over an all-false mask is 0 in the reference
    // pass 2: NaN repair (normal, centroid, pos of degenerate triangles <- first valid triangle); polar untouched
    const int no = order_seg ? 3 : 6, co = order_seg ? 7 : 0, po = order_seg ? 6 : 9;
    const float *f = o + first_ok * ld;
    const float fn0 = f[no], fn1 = f[no + 1], fn2 = f[no + 2], fc0 = f[co], fc1 = f[co + 1], fc2 = f[co + 2], fp = f[po];
    for (int i = 0; i < G; i++) {
        float *t = o + i * ld;
        if (isnan(t[no]) || isnan(t[no + 1]) || isnan(t[no + 2])) {
            t[no] = fn0; t[no + 1] = fn1; t[no + 2] = fn2;
            t[co] = fc0; t[co + 1] = fc1; t[co + 2] = fc2;
            t[po] = fp;
        }
        for (int c = channels; c < ld; c++) t[c] = 0.f;    // dropped channel (return_dist=False) and alignment padding
    }
}

}  // namespace

// xyz [rows,3]; idx [np,k] global row ids; flip [np] (+1/-1) or NULL; out [np, G, ld], G = k - (skip_first ? 1 : 0):
// the first `channels` (9 or 10) descriptor channels of every triangle, columns channels..ld-1 zero.
RSB_EXPORT int rsb_umbrella_features(long np, int k, int skip_first, int rotate_key, int order_seg, const float *xyz,
                                     const int *idx, const float *flip, float *out, int channels, int ld, cudaStream_t stream)
{
    RSB_REQUIRE(k >= 2 && k - (skip_first ? 1 : 0) <= MAXG, "group size out of range");
    RSB_REQUIRE((channels == 9 || channels == 10) && ld >= 10, "channels must be 9 or 10 and ld >= 10");
    if (np == 0) return 0;
    umbrella_kernel<<<(unsigned)((np + 127) / 128), 128, 0, stream>>>(np, k, skip_first, rotate_key, order_seg, xyz, idx, flip, out, channels, ld);
    RSB_CHECK_LAUNCH("umbrella_kernel");
    RSB_COUNT_LAUNCH(1);
    return 0;
}
