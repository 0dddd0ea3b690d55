// match_signatures_sharded: run_test.m:25-57 against a signature database that is row-sharded over the GPUs of one
// box, from a C++ host: one thread per GPU, each with its own sodso_ctx and the library's NCCL
// communicator (sodso_comm_init); queries are streamed in batches through sodso_db_query_sharded.
//   match_signatures_sharded sc|m2dp <history_db> <history_queries> <mask_width> <k> <out_topk.txt>
//                            [--gpus N] [--batch 128] [--q-row0 R] [--p-weight 2]
// history files: the reference's text matrices (test_sc.cpp:63-66 / test_kitti.m:26) or the binary container.
// --q-row0: global row number of query 0 for the temporal mask |q - j| < mask_width (run_test.m:47-53); for a self-match
// (history_queries == history_db, test_kitti.m:28) it is 0, for queries newer than the database pass its row count.
// out_topk.txt: one line per query: k x "idx(1-based, like MATLAB; 0 = none) fused_score d_p d_i".
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

#include "sodso_host.hpp"

int main(int argc, char **argv) {
  using namespace sodso_host;
  if (argc < 7) {
    std::fprintf(stderr, "usage: %s sc|m2dp history_db history_queries mask_width k out_topk [--gpus N] [--batch B] "
                         "[--q-row0 R] [--p-weight W]\n", argv[0]);
    return 1;
  }
  const std::string type = argv[1];
  const int t = type == "sc" ? SODSO_TYPE_SC : type == "m2dp" ? SODSO_TYPE_M2DP : -1;
  if (t < 0) {
    std::fprintf(stderr, "unknown descriptor type %s\n", type.c_str());
    return 1;
  }
  const int mask_width = std::atoi(argv[4]), k = std::atoi(argv[5]);
  int gpus = 1, batch = 128;
  long long q_row0 = 0;
  double p_weight = 2.0;   // run_test.m:39
  for (int i = 7; i < argc; i++) {
    if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--batch") && i + 1 < argc) batch = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--q-row0") && i + 1 < argc) q_row0 = std::atoll(argv[++i]);
    else if (!std::strcmp(argv[i], "--p-weight") && i + 1 < argc) p_weight = std::atof(argv[++i]);
  }
  if (gpus < 1 || gpus > 16 || batch < 1 || k < 1) {
    std::fprintf(stderr, "bad --gpus / --batch / k\n");
    return 1;
  }
  try {
    size_t r1, c1, r2, c2;
    const std::vector<double> hq = read_history(argv[3], r1, c1), hdb = read_history(argv[2], r2, c2);
    const size_t width = t == SODSO_TYPE_SC ? 2 * SODSO_SC_SIZE : 2 * SODSO_M2DP_SIZE, per = t == SODSO_TYPE_SC ? 1 : 4;
    if (c1 != width || c2 != width || r1 % per || r2 % per) throw std::runtime_error("history matrix has the wrong shape");
    const int m = (int)(r1 / per), n = (int)(r2 / per);
    if (n < gpus) throw std::runtime_error("fewer database rows than GPUs");
    unsigned char id[SODSO_COMM_ID_BYTES] = {0};
    if (gpus > 1) check(sodso_comm_unique_id(id), "sodso_comm_unique_id");
    std::vector<int64_t> idx((size_t)m * k);
    std::vector<double> score((size_t)m * k), dpa((size_t)m * k), dia((size_t)m * k);
    std::vector<std::string> errors((size_t)gpus);
    std::atomic<int> failed{0};
    double secs = 0.0;
    auto worker = [&](int rank) {
      try {
        Context ctx(rank);
        check(sodso_comm_init(ctx.get(), id, gpus, rank), "sodso_comm_init");
        // contiguous block partition of the database rows (rank r: base rows, the first n % R ranks one more)
        const int base = n / gpus, rem = n % gpus;
        const int n_local = base + (rank < rem ? 1 : 0), row0 = rank * base + std::min(rank, rem);
        sodso_db *db = nullptr;
        check(sodso_db_create(ctx.get(), t, hdb.data() + (size_t)row0 * per * width, n_local, row0, &db), "sodso_db_create");
        // every rank receives the same merged lists; rank 0 writes them into the shared result
        std::vector<int64_t> li((size_t)batch * k);
        std::vector<double> ls((size_t)batch * k), lp((size_t)batch * k), ld((size_t)batch * k);
        auto t0 = std::chrono::steady_clock::now();   // tic (run_test.m:25)
        for (int b = 0; b < m; b += batch) {
          const int mb = std::min(batch, m - b);
          const bool out0 = rank == 0;
          check(sodso_db_query_sharded(db, hq.data() + (size_t)b * per * width, mb, q_row0 + b, mask_width, p_weight, k,
                                       out0 ? idx.data() + (size_t)b * k : li.data(),
                                       out0 ? score.data() + (size_t)b * k : ls.data(),
                                       out0 ? dpa.data() + (size_t)b * k : lp.data(),
                                       out0 ? dia.data() + (size_t)b * k : ld.data()),
                "sodso_db_query_sharded");
        }
        auto t1 = std::chrono::steady_clock::now();   // toc (run_test.m:42-44)
        if (rank == 0) secs = std::chrono::duration<double>(t1 - t0).count();
        sodso_db_destroy(db);
        check(sodso_comm_finalize(ctx.get()), "sodso_comm_finalize");
      } catch (const std::exception &e) {
        errors[(size_t)rank] = e.what();
        failed++;
      }
    };
    std::vector<std::thread> th;
    for (int r = 1; r < gpus; r++) th.emplace_back(worker, r);
    worker(0);
    for (auto &x : th) x.join();
    if (failed) {
      for (int r = 0; r < gpus; r++)
        if (!errors[(size_t)r].empty()) std::fprintf(stderr, "rank %d: %s\n", r, errors[(size_t)r].c_str());
      return 2;
    }
    std::printf("%s\ntm = %.6f ms per query (%d queries x %d database rows on %d GPU(s), batches of %d, top-%d)\n",
                type.c_str(), 1e3 * secs / std::max(m, 1), m, n, gpus, batch, k);
    std::ofstream f(argv[6]);
    f << std::setprecision(17);
    for (int i = 0; i < m; i++) {
      for (int r = 0; r < k; r++) {
        const size_t o = (size_t)i * k + r;
        f << (r ? " " : "") << idx[o] + 1 << " " << score[o] << " " << dpa[o] << " " << dia[o];
      }
      f << "\n";
    }
  } catch (const std::exception &e) {
    std::fprintf(stderr, "match_signatures_sharded: %s\n", e.what());
    return 2;
  }
  return 0;
}
