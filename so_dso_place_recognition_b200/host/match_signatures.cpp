// match_signatures: run_test.m:25-57 (match + fuse + mask + arg-min) on two history files.
//   match_signatures sc|m2dp|delight <hist1.txt> <hist2.txt> <mask_width> <out_loops.txt> [p_weight=2] [--matrices prefix]
// out_loops.txt: one line per query "diff_idx(1-based, like MATLAB) diff_v d_p d_i".
// --matrices prefix additionally writes prefix_p.txt / prefix_i.txt (processSC / processM2DP outputs).
#include <chrono>
#include <cstring>

#include "sodso_host.hpp"

int main(int argc, char **argv) {
  using namespace sodso_host;
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s sc|m2dp|delight hist1 hist2 mask_width out_loops [p_weight] [--matrices prefix]\n", argv[0]);
    return 1;
  }
  const std::string type = argv[1];
  const bool delight = type == "delight";
  const int t = type == "sc" ? SODSO_TYPE_SC : type == "m2dp" ? SODSO_TYPE_M2DP : -1;
  if (t < 0 && !delight) {
    std::fprintf(stderr, "unknown descriptor type %s\n", type.c_str());
    return 1;
  }
  const int mask_width = std::atoi(argv[4]);
  double p_weight = 2.0;   // run_test.m:39
  std::string prefix;
  for (int i = 6; i < argc; i++) {
    if (!std::strcmp(argv[i], "--matrices") && i + 1 < argc) prefix = argv[++i];
    else p_weight = std::atof(argv[i]);
  }
  try {
    size_t r1, c1, r2, c2;
    std::vector<double> h1 = read_history(argv[2], r1, c1), h2 = read_history(argv[3], r2, c2);
    if (delight) {   // run_test.m:31-32: one distance matrix, no fusion
      if (c1 != 256 || c2 != 256 || r1 % 16 || r2 % 16) throw std::runtime_error("history matrix has the wrong shape");
      const int m = (int)(r1 / 16), n = (int)(r2 / 16);
      Context ctx(0);
      std::vector<double> dist((size_t)m * n), score(m);
      std::vector<int32_t> idx(m);
      auto t0 = std::chrono::steady_clock::now();
      check(sodso_delight_match(ctx.get(), h1.data(), m, h2.data(), n, dist.data()), "sodso_delight_match");
      check(sodso_top1_single(ctx.get(), dist.data(), m, n, mask_width, idx.data(), score.data()), "sodso_top1_single");
      auto t1 = std::chrono::steady_clock::now();
      std::printf("%s\ntm = %.6f ms per query (%d x %d pairs)\n", type.c_str(),
                  1e3 * std::chrono::duration<double>(t1 - t0).count() / std::max(m, 1), m, n);
      std::ofstream f(argv[5]);
      f << std::setprecision(17);
      for (int i = 0; i < m; i++) f << idx[i] + 1 << " " << score[i] << "\n";
      if (!prefix.empty()) write_history(prefix + "_d.txt", dist.data(), m, n, true);
      return 0;
    }
    const size_t width = t == SODSO_TYPE_SC ? 2 * SODSO_SC_SIZE : 2 * SODSO_M2DP_SIZE, per = t == SODSO_TYPE_SC ? 1 : 4;
    if (c1 != width || c2 != width || r1 % per || r2 % per) throw std::runtime_error("history matrix has the wrong shape");
    const int m = (int)(r1 / per), n = (int)(r2 / per);
    Context ctx(0);
    std::vector<int32_t> idx(m);
    std::vector<double> score(m), dpa(m), dia(m);
    auto t0 = std::chrono::steady_clock::now();   // tic (run_test.m:25)
    check(sodso_loop_top1(ctx.get(), t, h1.data(), m, h2.data(), n, mask_width, p_weight, idx.data(), score.data(),
                          dpa.data(), dia.data()), "sodso_loop_top1");
    auto t1 = std::chrono::steady_clock::now();   // toc (run_test.m:42-44)
    std::printf("%s\ntm = %.6f ms per query (%d x %d pairs)\n", type.c_str(),
                1e3 * std::chrono::duration<double>(t1 - t0).count() / std::max(m, 1), m, n);
    std::ofstream f(argv[5]);
    f << std::setprecision(17);
    for (int i = 0; i < m; i++) f << idx[i] + 1 << " " << score[i] << " " << dpa[i] << " " << dia[i] << "\n";
    if (!prefix.empty()) {
      std::vector<double> dp((size_t)m * n), di((size_t)m * n);
      check(t == SODSO_TYPE_SC ? sodso_sc_match(ctx.get(), h1.data(), m, h2.data(), n, dp.data(), di.data())
                               : sodso_m2dp_match(ctx.get(), h1.data(), m, h2.data(), n, dp.data(), di.data()),
            "match");
      write_history(prefix + "_p.txt", dp.data(), m, n, true);
      write_history(prefix + "_i.txt", di.data(), m, n, true);
    }
  } catch (const std::exception &e) {
    std::fprintf(stderr, "match_signatures: %s\n", e.what());
    return 2;
  }
  return 0;
}
