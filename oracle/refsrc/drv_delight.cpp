// ORACLE -- test infrastructure only.  Calls the reference's own DELIGHT class (DELIGHT.cpp, compiled unchanged)
// the way test_delight.cpp:38-56 does.  Points whose int(intensity) falls outside [0, 255] would index out of the
// 16 x 256 Eigen matrix in the reference (undefined behaviour); callers keep intensities inside that range.
#include "place_recognition/generate_signatures/src/DELIGHT/DELIGHT.h"
#include <cstdint>

extern "C" void ref_delight_generate(const double* xyz, const float* inten, const int64_t* off, int nscan,
                                     double* hist) {
  DELIGHT delight;
  const int w = (int)delight.getSignatureSize();
  for (int s = 0; s < nscan; s++) {
    std::vector<std::pair<Eigen::Vector3d, float>> pts;
    for (int64_t i = off[s]; i < off[s + 1]; i++)
      pts.push_back({Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), inten[i]});
    Eigen::MatrixXd signature;
    delight.getSignature(pts, signature);
    for (int r = 0; r < 16; r++)
      for (int c = 0; c < w; c++) hist[((size_t)16 * s + r) * w + c] = signature(r, c);
  }
}
