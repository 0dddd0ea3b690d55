// Host-side C++ mirror of the reference interface for the descriptor hot path, over the C ABI
// (include/sodso_pr.h).  Eigen-free: a 3-vector is std::array<double,3>.
//
//   IDPose / IDPtIntensity + text I/O   <- PosesPts.h:5-40 (same token order / default precision)
//   read_poses_pts, pts_preprocess      <- place_recognition/generate_signatures/src/utils/pts_preprocess.h
//                                          (CPU staging in the reference's own container order, so that the point
//                                          order and incoming_id_file.txt are byte-identical; the GPU staging is
//                                          sodso_stage_points, used by host/place_recognition.cpp)
//   class SC, class M2DP, class DELIGHT <- .../src/SC/SC.h:10-23, .../src/M2DP/M2DP.h:12-30, .../src/DELIGHT/DELIGHT.h:12-20
//   align_points_PCA                    <- .../src/utils/pts_align.h:7-9
//   write_history / read_history        <- test_sc.cpp:63-66 (Eigen operator<<), test_kitti.m:26 (load)
//   write/read/append_history_bin       <- (new) mmap-able binary container for the same matrices, SURVEY §8f N2
//
// All descriptor arithmetic happens in libsodso_pr.so on the GPU; this header only marshals.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/sodso_pr.h"

namespace sodso_host {

using Vec3 = std::array<double, 3>;
using PtI = std::pair<Vec3, float>;   // stands in for std::pair<Eigen::Vector3d, float>
using Scan = std::vector<PtI>;

struct IDPose {              // PosesPts.h:5-25
  int incoming_id = 0;
  double w2c[3][4] = {};
};
struct IDPtIntensity {       // PosesPts.h:27-40
  int incoming_id = 0;
  Vec3 pt{};
  float intensity = 0;
};

inline std::ostream &operator<<(std::ostream &os, const IDPose &p) {
  os << p.incoming_id << " ";
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 4; j++) os << p.w2c[i][j] << " ";
  os << std::endl;
  return os;
}
inline std::ostream &operator<<(std::ostream &os, const IDPtIntensity &p) {
  os << p.incoming_id << " " << p.pt[0] << " " << p.pt[1] << " " << p.pt[2] << " " << p.intensity << std::endl;
  return os;
}

inline void read_poses_pts(const std::string &poses_file, const std::string &pts_file, std::vector<IDPose> &poses,
                           std::vector<IDPtIntensity> &pts) {
  std::ifstream fp(poses_file);
  for (IDPose p; fp >> p.incoming_id;) {
    bool ok = true;
    for (int i = 0; i < 3 && ok; i++)
      for (int j = 0; j < 4 && ok; j++) ok = bool(fp >> p.w2c[i][j]);
    poses.push_back(p);
  }
  std::ifstream fq(pts_file);
  for (IDPtIntensity q; fq >> q.incoming_id >> q.pt[0] >> q.pt[1] >> q.pt[2] >> q.intensity;) pts.push_back(q);
}

// ---- staging (pts_preprocess.h:51-232) -----------------------------------------------------
namespace detail {
constexpr int kInitFrame = 30;         // INIT_FRAME
constexpr double kResGrid = 30;        // RES_GRID
inline double res_polar() { return 1.0 / 180.0 * M_PI; }   // RES_POLAR

struct LocalPt {
  Vec3 p;
  float it;
};

// "highest point per voxel": keep the point with the smallest y (camera y points down)
inline void filter_grid(const std::vector<LocalPt> &in, double range, Scan &out) {
  const double res[3] = {kResGrid, 2 * kResGrid, kResGrid};
  double step[3];
  int dim[3];
  for (int a = 0; a < 3; a++) {
    step[a] = 1.0 / (range / res[a]);
    dim[a] = static_cast<int>(std::floor(2 * range * step[a]) + 1);
  }
  std::unordered_map<int, int> best;   // voxel -> index into `in`; iteration order = reference's (T16)
  for (int i = 0; i < (int)in.size(); i++) {
    const Vec3 &p = in[i].p;
    const int xi = static_cast<int>(std::floor((p[0] + range) * step[0]));
    const int yi = static_cast<int>(std::floor((p[1] + range) * step[1]));
    const int zi = static_cast<int>(std::floor((p[2] + range) * step[2]));
    const int loc = xi + yi * dim[0] + zi * dim[0] * dim[1];
    auto it = best.find(loc);
    if (it == best.end())
      best[loc] = i;
    else if (-in[it->second].p[1] < -p[1])
      it->second = i;
  }
  for (auto &kv : best) out.push_back({in[kv.second].p, in[kv.second].it});
}

// "closest point per 1-degree polar cell"
inline void filter_polar(const std::vector<LocalPt> &in, Scan &out) {
  const double inv = 1.0 / res_polar();
  const int azi_bins = static_cast<int>(std::floor(2 * M_PI * inv) + 1);
  auto norm = [](const Vec3 &p) { return std::sqrt((p[0] * p[0] + p[1] * p[1]) + p[2] * p[2]); };
  std::unordered_map<int, int> best;
  for (int i = 0; i < (int)in.size(); i++) {
    const Vec3 &p = in[i].p;
    const double xz = std::sqrt(p[0] * p[0] + p[2] * p[2]);
    const int azi = static_cast<int>(std::floor((std::atan2(p[2], p[0]) + M_PI) * inv));
    const int ele = static_cast<int>(std::floor((std::atan2(p[1], xz) + M_PI / 2) * inv));
    const int loc = azi + ele * azi_bins;
    auto it = best.find(loc);
    if (it == best.end())
      best[loc] = i;
    else if (norm(in[it->second].p) > norm(p))
      it->second = i;
  }
  for (auto &kv : best) out.push_back({in[kv.second].p, in[kv.second].it});
}
}  // namespace detail

// pts_preprocess (pts_preprocess.h:169-232): writes incoming_id_file, fills one Scan per selected frame.
inline void pts_preprocess(const std::string &poses_file, const std::string &pts_file, const std::string &id_file,
                           double lidarRange, std::vector<Scan> &scans, bool polar_filter,
                           std::vector<int> *ids_out = nullptr) {
  std::vector<IDPose> poses;
  std::vector<IDPtIntensity> pts;
  read_poses_pts(poses_file, pts_file, poses, pts);
  std::ofstream ids;
  if (!id_file.empty()) ids.open(id_file);
  std::vector<int> nearby;   // indices into pts
  size_t next_pt = 0;
  int since_reset = 0;
  for (const IDPose &pose : poses) {
    const double tn = std::sqrt((pose.w2c[0][3] * pose.w2c[0][3] + pose.w2c[1][3] * pose.w2c[1][3]) +
                                pose.w2c[2][3] * pose.w2c[2][3]);
    if (tn < 1.0) {   // VO re-initialisation (pts_preprocess.h:189-193)
      since_reset = 0;
      nearby.clear();
    }
    while (next_pt < pts.size() && pts[next_pt].incoming_id <= pose.incoming_id) nearby.push_back((int)next_pt++);
    if (since_reset < detail::kInitFrame) {
      since_reset++;
      continue;
    }
    std::vector<detail::LocalPt> local;
    std::vector<int> keep;
    for (int k : nearby) {   // world -> camera, range crop (pts_preprocess.h:140-149)
      const Vec3 &g = pts[k].pt;
      detail::LocalPt l;
      for (int r = 0; r < 3; r++)
        l.p[r] = ((pose.w2c[r][0] * g[0] + pose.w2c[r][1] * g[1]) + pose.w2c[r][2] * g[2]) + pose.w2c[r][3] * 1.0;
      l.it = pts[k].intensity;
      if (std::sqrt((l.p[0] * l.p[0] + l.p[1] * l.p[1]) + l.p[2] * l.p[2]) < lidarRange) {
        local.push_back(l);
        keep.push_back(k);
      }
    }
    Scan s;
    if (polar_filter)
      detail::filter_polar(local, s);
    else
      detail::filter_grid(local, lidarRange, s);
    nearby.swap(keep);
    scans.push_back(std::move(s));
    if (ids.is_open()) ids << pose.incoming_id << std::endl;
    if (ids_out) ids_out->push_back(pose.incoming_id);
  }
}

// ---- flattening + history text I/O ------------------------------------------------------------
struct FlatScans {
  std::vector<double> xyz;
  std::vector<float> inten;
  std::vector<int64_t> off{0};
  int nscan() const { return (int)off.size() - 1; }
};
inline FlatScans flatten(const std::vector<Scan> &scans) {
  FlatScans f;
  for (const Scan &s : scans) {
    for (const PtI &p : s) {
      f.xyz.insert(f.xyz.end(), p.first.begin(), p.first.end());
      f.inten.push_back(p.second);
    }
    f.off.push_back((int64_t)f.inten.size());
  }
  return f;
}

// `stream << Eigen::MatrixXd` with the default IOFormat (test_sc.cpp:65, test_m2dp.cpp:85): stream
// precision (6 significant digits), every coefficient right-aligned to the widest one, " " between
// columns, "\n" between rows, no trailing newline.  full_precision = 17 digits (extension).
inline void write_history(const std::string &path, const double *m, size_t rows, size_t cols, bool full_precision = false) {
  std::vector<std::string> cell(rows * cols);
  size_t width = 0;
  for (size_t i = 0; i < rows * cols; i++) {
    std::ostringstream ss;
    if (full_precision) ss << std::setprecision(17);
    ss << m[i];
    cell[i] = ss.str();
    width = std::max(width, cell[i].size());
  }
  std::ofstream f(path);
  for (size_t r = 0; r < rows; r++) {
    if (r) f << "\n";
    for (size_t c = 0; c < cols; c++) {
      if (c) f << " ";
      f << std::setw((int)width) << cell[r * cols + c];
    }
  }
}

// ---- binary signature container ---------------------------------------------------------------
// The text hand-over costs 6 significant digits and, at 50 k scans, a gigabyte of decimal text.  The container is the
// same matrix as raw little-endian doubles behind a 64-byte header, so that it can be mmap-ed (the payload is 64-byte
// aligned) and appended to scan by scan:
//   bytes 0..7 "SODSOHST", u32 version = 1, u32 dtype = 0 (f64), u64 rows, u64 cols, 32 bytes reserved (zero).
struct HistoryBinHeader {
  char magic[8];
  uint32_t version, dtype;
  uint64_t rows, cols;
  unsigned char reserved[32];
};
static_assert(sizeof(HistoryBinHeader) == 64, "header layout");
constexpr char kHistoryMagic[9] = "SODSOHST";

inline bool is_history_bin(const std::string &path) {
  std::ifstream f(path, std::ios::binary);
  char m[8] = {};
  return f.read(m, 8) && !std::memcmp(m, kHistoryMagic, 8);
}

inline void write_history_bin(const std::string &path, const double *m, size_t rows, size_t cols) {
  HistoryBinHeader h{};
  std::memcpy(h.magic, kHistoryMagic, 8);
  h.version = 1;
  h.dtype = 0;
  h.rows = rows;
  h.cols = cols;
  std::ofstream f(path, std::ios::binary | std::ios::trunc);
  if (!f) throw std::runtime_error("cannot open " + path);
  f.write(reinterpret_cast<const char *>(&h), sizeof(h));
  f.write(reinterpret_cast<const char *>(m), (std::streamsize)(rows * cols * sizeof(double)));
}

// append rows (a new scan's signature rows) to an existing container, or create it
inline void append_history_bin(const std::string &path, const double *m, size_t rows, size_t cols) {
  if (!is_history_bin(path)) {
    write_history_bin(path, m, rows, cols);
    return;
  }
  std::fstream f(path, std::ios::binary | std::ios::in | std::ios::out);
  HistoryBinHeader h{};
  f.read(reinterpret_cast<char *>(&h), sizeof(h));
  if (h.version != 1 || h.dtype != 0 || h.cols != cols) throw std::runtime_error("append: container shape mismatch in " + path);
  f.seekp((std::streamoff)(sizeof(h) + h.rows * h.cols * sizeof(double)));
  f.write(reinterpret_cast<const char *>(m), (std::streamsize)(rows * cols * sizeof(double)));
  h.rows += rows;
  f.seekp(0);
  f.write(reinterpret_cast<const char *>(&h), sizeof(h));
}

inline std::vector<double> read_history_bin(const std::string &path, size_t &rows, size_t &cols) {
  std::ifstream f(path, std::ios::binary);
  HistoryBinHeader h{};
  if (!f.read(reinterpret_cast<char *>(&h), sizeof(h)) || std::memcmp(h.magic, kHistoryMagic, 8) || h.version != 1 || h.dtype != 0)
    throw std::runtime_error("not a signature container: " + path);
  rows = h.rows;
  cols = h.cols;
  std::vector<double> v(rows * cols);
  if (!f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(v.size() * sizeof(double))))
    throw std::runtime_error("truncated signature container: " + path);
  return v;
}

inline bool has_suffix(const std::string &s, const std::string &suf) {
  return s.size() >= suf.size() && !s.compare(s.size() - suf.size(), suf.size(), suf);
}
// text (Eigen layout) unless the file name ends in ".bin"
inline void write_history_auto(const std::string &path, const double *m, size_t rows, size_t cols, bool full_precision = false) {
  if (has_suffix(path, ".bin")) write_history_bin(path, m, rows, cols);
  else write_history(path, m, rows, cols, full_precision);
}

// MATLAB load() of a whitespace text matrix (test_kitti.m:26); a binary container is recognised by its magic
inline std::vector<double> read_history(const std::string &path, size_t &rows, size_t &cols) {
  if (is_history_bin(path)) return read_history_bin(path, rows, cols);
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::vector<double> v;
  std::string line;
  rows = cols = 0;
  while (std::getline(f, line)) {
    std::istringstream ss(line);
    size_t c = 0;
    for (double x; ss >> x; c++) v.push_back(x);
    if (!c) continue;
    if (cols && c != cols) throw std::runtime_error("ragged matrix in " + path);
    cols = c;
    rows++;
  }
  return v;
}

// ---- C-ABI wrappers --------------------------------------------------------------------------------
inline void check(int rc, const char *what) {
  if (rc != SODSO_OK) throw std::runtime_error(std::string(what) + ": " + sodso_last_error());
}

class Context {
 public:
  explicit Context(int device = 0) { check(sodso_ctx_create(device, &c_), "sodso_ctx_create"); }
  ~Context() { sodso_ctx_destroy(c_); }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  sodso_ctx *get() const { return c_; }

 private:
  sodso_ctx *c_ = nullptr;
};

// pts_align.h:7-9
inline void align_points_PCA(Context &ctx, const Scan &in, Scan &out) {
  FlatScans f = flatten({in});
  std::vector<double> o(f.xyz.size());
  check(sodso_align_pca(ctx.get(), f.xyz.data(), f.off.data(), 1, o.data(), nullptr), "sodso_align_pca");
  out.clear();
  for (size_t i = 0; i < in.size(); i++) out.push_back({{o[3 * i], o[3 * i + 1], o[3 * i + 2]}, in[i].second});
}

// SC.h:10-23
class SC {
 public:
  SC(Context &ctx, double max_rho) : ctx_(ctx), max_rho_(max_rho) {}
  unsigned int getSignatureSize() const { return (unsigned)sodso_sc_signature_size(); }
  void getSignature(const Scan &pts_clr_raw, std::vector<double> &structure_output, std::vector<double> &intensity_output) {
    std::vector<double> h = getSignatures({pts_clr_raw});
    structure_output.assign(h.begin(), h.begin() + getSignatureSize());
    intensity_output.assign(h.begin() + getSignatureSize(), h.end());
  }
  // whole batch in one launch: history_sc (test_sc.cpp:36-57), nscan x 2400 row-major
  std::vector<double> getSignatures(const std::vector<Scan> &scans) {
    FlatScans f = flatten(scans);
    std::vector<double> hist((size_t)f.nscan() * 2 * getSignatureSize());
    check(sodso_sc_generate(ctx_.get(), f.xyz.data(), f.inten.data(), f.off.data(), f.nscan(), max_rho_, hist.data()),
          "sodso_sc_generate");
    return hist;
  }

 private:
  Context &ctx_;
  double max_rho_;
};

// M2DP.h:12-30 (getSignature expects aligned + sign-flipped points, like the reference class)
class M2DP {
 public:
  M2DP(Context &ctx, double max_rho) : ctx_(ctx), max_rho_(max_rho) {}
  unsigned int getSignatureSize() const { return (unsigned)sodso_m2dp_signature_size(); }
  void getSignature(const Scan &pts_clr, std::vector<double> &count_output, std::vector<double> &intensity_output) {
    FlatScans f = flatten({pts_clr});
    std::vector<double> sig(2 * getSignatureSize());
    check(sodso_m2dp_signature(ctx_.get(), f.xyz.data(), f.inten.data(), f.off.data(), 1, max_rho_, sig.data()),
          "sodso_m2dp_signature");
    count_output.assign(sig.begin(), sig.begin() + getSignatureSize());
    intensity_output.assign(sig.begin() + getSignatureSize(), sig.end());
  }
  // test_m2dp.cpp:37-67 for a batch: PCA + 4 variants per scan, 4*nscan x 384 row-major
  std::vector<double> getHistory(const std::vector<Scan> &scans) {
    FlatScans f = flatten(scans);
    std::vector<double> hist((size_t)f.nscan() * 4 * 2 * getSignatureSize());
    check(sodso_m2dp_generate(ctx_.get(), f.xyz.data(), f.inten.data(), f.off.data(), f.nscan(), max_rho_, hist.data()),
          "sodso_m2dp_generate");
    return hist;
  }

 private:
  Context &ctx_;
  double max_rho_;
};

// DELIGHT.h:12-20 (getSignature aligns by PCA itself, DELIGHT.cpp:10-11)
class DELIGHT {
 public:
  explicit DELIGHT(Context &ctx) : ctx_(ctx) {}
  unsigned int getSignatureSize() const { return (unsigned)sodso_delight_signature_size(); }
  // 16 x 256 row-major (Eigen::MatrixXd output(16, BINS))
  void getSignature(const Scan &pts_clr_raw, std::vector<double> &output) {
    FlatScans f = flatten({pts_clr_raw});
    output.assign((size_t)16 * getSignatureSize(), 0.0);
    check(sodso_delight_generate(ctx_.get(), f.xyz.data(), f.inten.data(), f.off.data(), 1, output.data()),
          "sodso_delight_generate");
  }
  // test_delight.cpp:38-56 for a batch: 16*nscan x 256 row-major
  std::vector<double> getHistory(const std::vector<Scan> &scans) {
    FlatScans f = flatten(scans);
    std::vector<double> hist((size_t)f.nscan() * 16 * getSignatureSize());
    check(sodso_delight_generate(ctx_.get(), f.xyz.data(), f.inten.data(), f.off.data(), f.nscan(), hist.data()),
          "sodso_delight_generate");
    return hist;
  }

 private:
  Context &ctx_;
};

}  // namespace sodso_host
