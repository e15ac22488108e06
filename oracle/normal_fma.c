/* Un-normalised normal (forward-difference cross product) and its length, rounded exactly like the
 * reference's CPU build.  TEST INFRASTRUCTURE (oracle), not product code.
 *
 * gradslam/structures/rgbdimages.py:733-734 calls torch.cross and Tensor.norm on float32.  On the
 * AVX2 / AVX-512 builds of torch (checked against torch 2.11 in the build container, bit for bit on
 * every element) these evaluate as
 *     cross.x = fma(a.y, b.z, -(a.z * b.y))     (first product exact, second rounded; y, z likewise)
 *     norm    = sqrt(fma(c.z, c.z, fma(c.y, c.y, c.x * c.x)))
 * fmaf() is exact by definition (one rounding), so this file IS that arithmetic; it is compiled with
 * -ffp-contract=off so nothing else is fused.
 */
#include <math.h>
#include <stdint.h>

void gsx_oracle_cross_norm_fma(const float *a, const float *b, int64_t n, float *cross, float *norm) {
  for (int64_t i = 0; i < n; ++i) {
    const float ax = a[3 * i], ay = a[3 * i + 1], az = a[3 * i + 2];
    const float bx = b[3 * i], by = b[3 * i + 1], bz = b[3 * i + 2];
    const float p0 = az * by, p1 = ax * bz, p2 = ay * bx;
    const float cx = fmaf(ay, bz, -p0);
    const float cy = fmaf(az, bx, -p1);
    const float cz = fmaf(ax, by, -p2);
    cross[3 * i] = cx;
    cross[3 * i + 1] = cy;
    cross[3 * i + 2] = cz;
    const float xx = cx * cx;
    norm[i] = sqrtf(fmaf(cz, cz, fmaf(cy, cy, xx)));
  }
}
