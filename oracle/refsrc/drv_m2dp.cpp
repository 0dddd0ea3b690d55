// ORACLE -- test infrastructure only.  Calls the reference's own M2DP class (M2DP.cpp, compiled unchanged) and
// align_points_PCA (pts_align.h) the way test_m2dp.cpp:38-70 does: four direction variants per scan.
#include "place_recognition/generate_signatures/src/M2DP/M2DP.h"
#include "place_recognition/generate_signatures/src/utils/pts_align.h"
#include <cstdint>

extern "C" void ref_m2dp_generate(const double* xyz, const float* inten, const int64_t* off, int nscan, double max_rho,
                                  double* hist) {
  M2DP m2dp(max_rho);
  const int w = 2 * (int)m2dp.getSignatureSize();
  for (int s = 0; s < nscan; s++) {
    std::vector<std::pair<Eigen::Vector3d, float>> pts, aligned;
    for (int64_t i = off[s]; i < off[s + 1]; i++)
      pts.push_back({Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), inten[i]});
    align_points_PCA(pts, aligned);
    int subrow = 0;
    for (int direction_x = -1; direction_x < 2; direction_x += 2) {
      for (int direction_y = -1; direction_y < 2; direction_y += 2) {
        std::vector<std::pair<Eigen::Vector3d, float>> directed;
        for (auto& pc : aligned) {
          Eigen::Vector3d p;
          p << direction_x * pc.first[0], direction_y * pc.first[1], (direction_x * direction_y) * pc.first[2];
          directed.push_back({p, pc.second});
        }
        Eigen::VectorXd ct, ci;
        m2dp.getSignature(directed, ct, ci);
        Eigen::VectorXd signature(w);
        signature << ct, ci;
        double* row = hist + ((size_t)4 * s + subrow++) * w;
        for (int k = 0; k < w; k++) row[k] = signature(k);
      }
    }
  }
}

