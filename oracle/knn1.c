/* Exact brute-force 1-nearest-neighbour, CPU.  TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Restates the algorithm of chamferdist.chamfer.knn_points for K=1 (third-party, pinned
 * chamferdist==1.0.0 in the reference's requirements.txt:2; call site
 * gradslam/odometry/icputils.py:200-201; its source is NOT under /root/reference).  The published
 * algorithm (PyTorch3D-style knn) is: for every query point scan every target point, squared L2
 * distance accumulated over the 3 coordinates, keep the smallest; the first minimum wins, i.e. ties
 * resolve to the LOWEST target index.  Distances are rounded op by op in float32
 * ((dx*dx + dy*dy) + dz*dz); compile with -ffp-contract=off so gcc does not fuse them.
 *
 * threads <= 0: use all OpenMP threads (the CPU baseline is given every host core).
 */
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void gsx_oracle_knn1(const float *src, int64_t ns, const float *tgt, int64_t nt,
                     float *out_d2, int64_t *out_idx, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < ns; ++i) {
    const float sx = src[3 * i], sy = src[3 * i + 1], sz = src[3 * i + 2];
    float best = 0.0f;
    int64_t bi = -1;
    for (int64_t j = 0; j < nt; ++j) {
      const float dx = sx - tgt[3 * j], dy = sy - tgt[3 * j + 1], dz = sz - tgt[3 * j + 2];
      const float d = (dx * dx + dy * dy) + dz * dz;
      if (bi < 0 || d < best) {
        best = d;
        bi = j;
      }
    }
    out_d2[i] = best;
    out_idx[i] = bi;
  }
}
