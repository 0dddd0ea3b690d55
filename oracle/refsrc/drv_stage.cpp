// ORACLE -- test infrastructure only.  Runs the reference's own pts_preprocess (utils/pts_preprocess.h, compiled
// unchanged): reads poses_history_file / pts_history_file, writes incoming_id_file, returns the staged scans.
#include "place_recognition/generate_signatures/src/utils/pts_preprocess.h"
#include <cstdint>
#include <cstring>

namespace {
struct Staged {
  std::vector<std::vector<std::pair<Eigen::Vector3d, float>>> scans;
};
}  // namespace

extern "C" {

void* ref_stage_run(const char* poses_file, const char* pts_file, const char* incoming_id_file, double lidar_range,
                    int polar_filter) {
  auto* st = new Staged();
  std::string a(poses_file), b(pts_file), c(incoming_id_file);
  pts_preprocess(a, b, c, lidar_range, st->scans, polar_filter != 0);
  return st;
}

int ref_stage_num_scans(void* h) { return (int)((Staged*)h)->scans.size(); }

int64_t ref_stage_num_points(void* h) {
  int64_t n = 0;
  for (auto& s : ((Staged*)h)->scans) n += (int64_t)s.size();
  return n;
}

void ref_stage_copy(void* h, int64_t* off, double* xyz, float* inten) {
  auto* st = (Staged*)h;
  int64_t at = 0;
  off[0] = 0;
  for (size_t s = 0; s < st->scans.size(); s++) {
    for (auto& p : st->scans[s]) {
      xyz[3 * at + 0] = p.first(0);
      xyz[3 * at + 1] = p.first(1);
      xyz[3 * at + 2] = p.first(2);
      inten[at] = p.second;
      at++;
    }
    off[s + 1] = at;
  }
}

void ref_stage_free(void* h) { delete (Staged*)h; }

}  // extern "C"
