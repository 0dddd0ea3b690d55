// gen_signatures: the batch drivers test_sc.cpp / test_m2dp.cpp / test_delight.cpp without ROS.
//   gen_signatures sc|m2dp|delight <poses_history_file> <pts_history_file> <out_history_file> <incoming_id_file>
//                  [lidarRange=45] [--full-precision] [--stage-only]
// An output name ending in .bin selects the binary signature container (sodso_host.hpp) instead of text.
// Same inputs / outputs as lidar.launch:15-21: reads the two SO-DSO text files, writes
// incoming_id_file.txt and history_sc.txt (N x 2400) or history_m2dp.txt (4N x 384).
// --stage-only stops after pts_preprocess (no GPU needed): used by the incoming-id known-answer test.
#include <chrono>
#include <cstring>

#include "sodso_host.hpp"

int main(int argc, char **argv) {
  using namespace sodso_host;
  if (argc == 4 && !std::strcmp(argv[1], "reformat")) {   // read a text matrix, write it back Eigen-style
    size_t r, c;
    std::vector<double> m = read_history(argv[2], r, c);
    write_history_auto(argv[3], m.data(), r, c);   // .bin -> binary container, else Eigen-style text
    return 0;
  }
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s sc|m2dp|delight poses_file pts_file out_file incoming_id_file [lidarRange] "
                         "[--full-precision] [--stage-only]\n", argv[0]);
    return 1;   // test_sc.cpp:19-25: missing parameters -> return 1
  }
  const std::string type = argv[1];
  double lidarRange = 45.0;   // test_sc.cpp:28
  bool full = false, stage_only = false;
  for (int i = 6; i < argc; i++) {
    if (!std::strcmp(argv[i], "--full-precision")) full = true;
    else if (!std::strcmp(argv[i], "--stage-only")) stage_only = true;
    else lidarRange = std::atof(argv[i]);
  }
  if (type != "sc" && type != "m2dp" && type != "delight") {
    std::fprintf(stderr, "unknown descriptor type %s\n", type.c_str());
    return 1;
  }
  try {
    std::vector<Scan> scans;
    auto t0 = std::chrono::steady_clock::now();
    pts_preprocess(argv[2], argv[3], argv[5], lidarRange, scans, /*polar_filter=*/type != "sc");   // test_sc.cpp:32-33,
                                                                                                  // test_m2dp.cpp:33-34, test_delight.cpp:33-34
    auto t1 = std::chrono::steady_clock::now();
    size_t npts = 0;
    for (auto &s : scans) npts += s.size();
    std::printf("staged %zu scans, %.1f points on average, %.3f ms per frame\n", scans.size(),
                scans.empty() ? 0.0 : double(npts) / scans.size(),
                scans.empty() ? 0.0 : 1e3 * std::chrono::duration<double>(t1 - t0).count() / scans.size());
    if (stage_only) return 0;
    Context ctx(0);
    std::vector<double> hist;
    size_t rows, cols;
    auto g0 = std::chrono::steady_clock::now();
    if (type == "sc") {
      SC sc(ctx, lidarRange);
      hist = sc.getSignatures(scans);
      rows = scans.size();
      cols = 2 * sc.getSignatureSize();
    } else if (type == "delight") {
      DELIGHT delight(ctx);
      hist = delight.getHistory(scans);
      rows = 16 * scans.size();
      cols = delight.getSignatureSize();
    } else {
      M2DP m2dp(ctx, lidarRange);
      hist = m2dp.getHistory(scans);
      rows = 4 * scans.size();
      cols = 2 * m2dp.getSignatureSize();
    }
    auto g1 = std::chrono::steady_clock::now();
    std::printf("%s average time: %.6f ms (kernel %s: %.3f ms for the batch)\n", type == "sc" ? "SC" : type == "m2dp" ? "M2DP" : "DELIGHT",
                scans.empty() ? 0.0 : 1e3 * std::chrono::duration<double>(g1 - g0).count() / scans.size(),
                sodso_ctx_last_kernel_name(ctx.get()), sodso_ctx_last_kernel_ms(ctx.get()));
    write_history_auto(argv[4], hist.data(), rows, cols, full);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "gen_signatures: %s\n", e.what());
    return 2;
  }
  return 0;
}
