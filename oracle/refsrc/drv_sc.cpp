// ORACLE -- test infrastructure only.  Calls the reference's own SC class (SC.cpp, compiled unchanged from
// /root/reference against oracle/eigen_shim) the way test_sc.cpp:35-56 does, from flat arrays.
#include "place_recognition/generate_signatures/src/SC/SC.h"
#include <cstdint>

extern "C" void ref_sc_generate(const double* xyz, const float* inten, const int64_t* off, int nscan, double max_rho,
                                double* hist) {
  SC sc(max_rho);
  const int w = (int)sc.getSignatureSize();
  for (int s = 0; s < nscan; s++) {
    std::vector<std::pair<Eigen::Vector3d, float>> pts;
    for (int64_t i = off[s]; i < off[s + 1]; i++)
      pts.push_back({Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), inten[i]});
    Eigen::VectorXd structure, intensity;
    sc.getSignature(pts, structure, intensity);
    Eigen::VectorXd signature(2 * w);                  // test_sc.cpp:52-54
    signature << structure, intensity;
    for (int k = 0; k < 2 * w; k++) hist[(size_t)s * 2 * w + k] = signature(k);
  }
}
