// place_recognition: the reference's three file-coupled stages in one process, everything heavy on the GPU:
//   pts_preprocess (pts_preprocess.h:169-232)  ->  test_sc.cpp:36-57 / test_m2dp.cpp:37-67  ->  run_test.m:1-85
//   place_recognition sc|m2dp <poses_history_file> <pts_history_file> <mask_width> <out_loops.txt>
//                     [--gt gt_file --loop-diff 10] [--lidar-range 45] [--history out_history.txt]
// poses / points: the SO-DSO outputs (PosesPts.h:12-24,35-39).  gt_file: KITTI ground-truth poses (12 numbers per
// line, row = incoming id, position = columns 4, 8, 12: test_kitti.m:23-25) or one "x y z" line per incoming id.
// out_loops.txt: one line per staged scan "incoming_id diff_idx(1-based) diff_v".  With --gt the precision-recall
// evaluation of run_test.m:58-85 is printed (AUC, top recall).  The scans never leave HBM between the stages; inside
// a scan the staged points are ordered by voxel index (see sodso_stage_points), not by libstdc++'s hash order.
#include <chrono>
#include <cstring>

#include "sodso_host.hpp"

namespace {
double secs(std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
  return std::chrono::duration<double>(b - a).count();
}
}  // namespace

int main(int argc, char **argv) {
  using namespace sodso_host;
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s sc|m2dp poses_file pts_file mask_width out_loops [--gt file --loop-diff d] "
                         "[--lidar-range r] [--history file]\n", argv[0]);
    return 1;
  }
  const std::string type = argv[1];
  if (type != "sc" && type != "m2dp") {
    std::fprintf(stderr, "unknown descriptor type %s\n", type.c_str());
    return 1;
  }
  const int mask_width = std::atoi(argv[4]);
  double lidar_range = 45.0, loop_diff = 10.0;   // test_sc.cpp:28, test_kitti.m:20
  std::string gt_file, hist_file;
  for (int i = 6; i < argc; i++) {
    if (!std::strcmp(argv[i], "--gt") && i + 1 < argc) gt_file = argv[++i];
    else if (!std::strcmp(argv[i], "--loop-diff") && i + 1 < argc) loop_diff = std::atof(argv[++i]);
    else if (!std::strcmp(argv[i], "--lidar-range") && i + 1 < argc) lidar_range = std::atof(argv[++i]);
    else if (!std::strcmp(argv[i], "--history") && i + 1 < argc) hist_file = argv[++i];
  }
  try {
    std::vector<IDPose> poses;
    std::vector<IDPtIntensity> pts;
    read_poses_pts(argv[2], argv[3], poses, pts);
    std::vector<int32_t> pose_id(poses.size()), pt_id(pts.size());
    std::vector<double> w2c(12 * poses.size()), pt_xyz(3 * pts.size());
    std::vector<float> pt_int(pts.size());
    for (size_t i = 0; i < poses.size(); i++) {
      pose_id[i] = poses[i].incoming_id;
      std::memcpy(&w2c[12 * i], poses[i].w2c, sizeof(double) * 12);
    }
    for (size_t i = 0; i < pts.size(); i++) {
      pt_id[i] = pts[i].incoming_id;
      for (int k = 0; k < 3; k++) pt_xyz[3 * i + k] = pts[i].pt[k];
      pt_int[i] = pts[i].intensity;
    }
    Context ctx(0);
    // ---- stage 1: pts_preprocess on the GPU
    auto t0 = std::chrono::steady_clock::now();
    sodso_staged *st = nullptr;
    check(sodso_stage_points(ctx.get(), pose_id.data(), w2c.data(), (int)poses.size(), pt_id.data(), pt_xyz.data(),
                             pt_int.data(), (int64_t)pts.size(), lidar_range, type == "m2dp" ? 1 : 0, &st),
          "sodso_stage_points");
    auto t1 = std::chrono::steady_clock::now();
    const int n = sodso_staged_num_scans(st);
    std::vector<int32_t> ids((size_t)n);
    check(sodso_staged_copy(st, ids.data(), nullptr, nullptr, nullptr), "sodso_staged_copy");
    std::printf("generate_spherical_points average time: %.4f ms average points: %.1f (%d scans)\n",
                n ? 1e3 * secs(t0, t1) / n : 0.0, n ? double(sodso_staged_num_points(st)) / n : 0.0, n);
    if (n == 0) throw std::runtime_error("no scans staged (fewer than 31 poses since the last reset)");
    // ---- stages 2 + 3: signatures and loop candidates
    std::vector<int32_t> idx((size_t)n);
    std::vector<double> score((size_t)n), hist;
    const size_t rows = type == "sc" ? n : 4 * (size_t)n, cols = type == "sc" ? 2 * SODSO_SC_SIZE : 2 * SODSO_M2DP_SIZE;
    if (!hist_file.empty() || type == "m2dp") hist.resize(rows * cols);
    auto t2 = std::chrono::steady_clock::now();
    if (type == "sc") {
      check(sodso_sc_scans_to_loops(ctx.get(), sodso_staged_xyz(st), sodso_staged_inten(st), sodso_staged_scan_off(st), n,
                                    lidar_range, mask_width, 2.0, hist.empty() ? nullptr : hist.data(), idx.data(),
                                    score.data(), nullptr, nullptr),
            "sodso_sc_scans_to_loops");
    } else {
      check(sodso_m2dp_generate(ctx.get(), sodso_staged_xyz(st), sodso_staged_inten(st), sodso_staged_scan_off(st), n,
                                lidar_range, hist.data()),
            "sodso_m2dp_generate");
      check(sodso_loop_top1(ctx.get(), SODSO_TYPE_M2DP, hist.data(), n, hist.data(), n, mask_width, 2.0, idx.data(),
                            score.data(), nullptr, nullptr),
            "sodso_loop_top1");
    }
    auto t3 = std::chrono::steady_clock::now();
    sodso_staged_destroy(st);
    std::printf("%s\ntm = %.6f ms per query (signatures + %d x %d pairs)\n", type.c_str(), 1e3 * secs(t2, t3) / n, n, n);
    if (!hist_file.empty()) write_history_auto(hist_file, hist.data(), rows, cols);
    {
      std::ofstream f(argv[5]);
      f << std::setprecision(17);
      for (int i = 0; i < n; i++) f << ids[i] << " " << idx[i] + 1 << " " << score[i] << "\n";
    }
    // ---- evaluation (run_test.m:2-22, 58-85) against ground-truth positions
    if (!gt_file.empty()) {
      size_t gr, gc;
      std::vector<double> g = read_history(gt_file, gr, gc);
      if (gc != 12 && gc != 3) throw std::runtime_error("gt file must have 12 (KITTI pose) or 3 (x y z) columns");
      std::vector<double> gt(3 * (size_t)n);
      for (int i = 0; i < n; i++) {
        const size_t row = (size_t)ids[i];                      // incoming_id + 1 in MATLAB's 1-based rows
        if (row >= gr) throw std::runtime_error("gt file has fewer rows than incoming ids");
        for (int k = 0; k < 3; k++) gt[3 * i + k] = gc == 12 ? g[row * 12 + 3 + 4 * k] : g[row * 3 + k];
      }
      std::vector<int32_t> nearest((size_t)n);
      int n_loops = 0;
      check(sodso_gt_loops(ctx.get(), gt.data(), n, gt.data(), n, loop_diff, mask_width, nearest.data(), nullptr, &n_loops),
            "sodso_gt_loops");
      double auc = 0, top_recall = 0;
      int top_count = 0;
      check(sodso_pr_curve(score.data(), idx.data(), gt.data(), n, gt.data(), n, loop_diff, n_loops, &auc, &top_recall,
                           &top_count, nullptr, nullptr, nullptr),
            "sodso_pr_curve");
      std::printf("total_lp = %d\nAUC = %.6f\ntop_recall = %.6f\nlp_detected = %d\n", n_loops, auc, top_recall, top_count);
    }
  } catch (const std::exception &e) {
    std::fprintf(stderr, "place_recognition: %s\n", e.what());
    return 2;
  }
  return 0;
}
