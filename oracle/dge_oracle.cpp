// ============================================================================
// dge_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY -- NOT PRODUCT CODE)
// See dge_oracle.hpp for scope and parity status ("parity unpinned").
// All file:line citations are into /root/reference.
// ============================================================================
#include "dge_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <cassert>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>

namespace orc {

static const double kPi = 3.14159265358979323846;
// ANALYSIS KNOB (default 0 = the pinned restatement): a shift of the vehicle position inside the nearest-frontier query only
// (exploration_env.py:327).  The reference starts every episode at integer coordinates and its four forced reset steps
// (1, 1, pi/2) walk a closed square, so the first decision is taken AT the integer start pose, where several frontier cell centres
// are at exactly the same distance (e.g. 5 = |(0,5)| = |(4,3)|); `dist < min_dist` then picks whichever the rounding noise of the
// pose estimate (1e-13) favours -- the reference's own choice is noise there.  tests/golden/scan_guided.py --knife tries the four
// sign combinations (+-1e-9, +-1e-9) to measure how many of the early stops these ties cause.  Process-global on purpose: it is
// a property of a scan run, not of an env.
double g_knife_dx = 0.0, g_knife_dy = 0.0;

// ------------------------------------------------------------ SE(2) algebra
// gtsam-4.0 Pose2 semantics restated: rotation normalised to
// (-pi, pi] through atan2(sin, cos) like Rot2::theta().
static inline double wrap_pi(double t) { return std::atan2(std::sin(t), std::cos(t)); }

static inline Pose compose(const Pose &a, const Pose &b) {
  const double c = std::cos(a.th), s = std::sin(a.th);
  return Pose{a.x + c * b.x - s * b.y, a.y + s * b.x + c * b.y, wrap_pi(a.th + b.th)};
}

// between(p1,p2) = p1^-1 o p2 ; H1 = d between / d p1 in body-frame tangents (H2 = I)
static inline Pose between(const Pose &p1, const Pose &p2, double *H1 /*9 or null*/) {
  const double c1 = std::cos(p1.th), s1 = std::sin(p1.th);
  const double c2 = std::cos(p2.th), s2 = std::sin(p2.th);
  const double c = c1 * c2 + s1 * s2, s = -s1 * c2 + c1 * s2;
  const double x = p2.x - p1.x, y = p2.y - p1.y;
  Pose r{c1 * x + s1 * y, -s1 * x + c1 * y, std::atan2(s, c)};
  if (H1) {
    const double dt1 = -s2 * x + c2 * y, dt2 = -c2 * x - s2 * y;
    H1[0] = -c; H1[1] = -s; H1[2] = dt1;
    H1[3] = s;  H1[4] = -c; H1[5] = dt2;
    H1[6] = 0;  H1[7] = 0;  H1[8] = -1;
  }
  return r;
}

// first-order chart (stock gtsam without SLOW_BUT_CORRECT_EXPMAP):
// retract(p, v) = p o Pose2(v), local(p, q) = between(p, q) as a 3-vector.
static inline Pose retract(const Pose &p, const double *v) { return compose(p, Pose{v[0], v[1], v[2]}); }

// BearingRangeSensorModel::measure (Simulator2D.cpp:113-132) without noise:
// bearing = atan2 of the body-frame point, range = |l - t|; Jacobians as
// assembled at Simulator2D.cpp:124-130 (gtsam Pose2::bearing / Pose2::range).
static inline void predict_br(const Pose &p, double lx, double ly, double &bearing, double &range,
                              double *Hx /*2x3 or null*/, double *Hl /*2x2 or null*/) {
  const double c = std::cos(p.th), s = std::sin(p.th);
  const double dx = lx - p.x, dy = ly - p.y;
  const double qx = c * dx + s * dy, qy = -s * dx + c * dy;
  range = std::sqrt(dx * dx + dy * dy);
  bearing = std::atan2(qy, qx);
  if (Hx) {
    const double r2 = qx * qx + qy * qy;
    const double bx = -qy / r2, by = qx / r2;       // d bearing / d q
    const double rx = qx / range, ry = qy / range;  // d range / d q
    // dq/dpose = [[-1,0,qy],[0,-1,-qx]] ; dq/dl = R^T
    Hx[0] = -bx; Hx[1] = -by; Hx[2] = bx * qy - by * qx;
    Hx[3] = -rx; Hx[4] = -ry; Hx[5] = rx * qy - ry * qx;
    Hl[0] = bx * c - by * s;  Hl[1] = bx * s + by * c;
    Hl[2] = rx * c - ry * s;  Hl[3] = rx * s + ry * c;
  }
}

// --------------------------------------------------------- small matrices ---
// Utils.h:29-33 : inverse<N> = m.llt().solve(Identity)
static bool chol_inv(int n, const double *A, double *Ainv) {
  // fixed-size path (2x2 / 3x3, like Eigen's fixed-size LLT in the reference): no heap traffic
  double Lsmall[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ysmall[3];
  std::vector<double> Lbig, ybig;
  double *L = Lsmall, *y = ysmall;
  if (n > 3) { Lbig.assign(static_cast<size_t>(n) * n, 0.0); ybig.resize(n); L = Lbig.data(); y = ybig.data(); }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      if (i == j) {
        if (!(s > 0)) return false;
        L[i * n + i] = std::sqrt(s);
      } else {
        L[i * n + j] = s / L[j * n + j];
      }
    }
  for (int col = 0; col < n; ++col) {
    for (int i = 0; i < n; ++i) {
      double s = (i == col) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s -= L[i * n + k] * y[k];
      y[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = y[i];
      for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * Ainv[k * n + col];
      Ainv[i * n + col] = s / L[i * n + i];
    }
  }
  return true;
}
static inline M3 inv3(const M3 &m) { M3 r; if (!chol_inv(3, m.a, r.a)) for (double &v : r.a) v = std::numeric_limits<double>::quiet_NaN(); return r; }
static inline M2 inv2(const M2 &m) { M2 r; if (!chol_inv(2, m.a, r.a)) for (double &v : r.a) v = std::numeric_limits<double>::quiet_NaN(); return r; }
static inline double det3(const M3 &m) {
  const double *a = m.a;
  return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
}
static inline double det2(const M2 &m) { return m.a[0] * m.a[3] - m.a[1] * m.a[2]; }

// OccupancyMap.h:10-19 (incl. the MAX_LOGODDS quirk q7)
static inline double prob2logodds(double p) { return std::log(p / (1.0 - p)); }
static inline double logodds2prob(double l) { return std::exp(l) / (1.0 + std::exp(l)); }
static const double LOGODDS_FREE = prob2logodds(0.3);
static const double LOGODDS_OCCUPIED = prob2logodds(0.7);
static const double MIN_LOGODDS = prob2logodds(0.05);
static const double MAX_LOGODDS = logodds2prob(0.95);
static const double OCCUPIED_THRESH = prob2logodds(0.5);
#define ORC_DEG2RAD(x) ((x)*0.01745329251994329575)   // Utils.h:11

// ------------------------------------------------------------ virtual map ---
// VirtualMap.cpp:364-377
static inline M2 covariance_intersection(const M2 &m1, const M2 &m2) {
  const double a = det2(m1), b = det2(m2);
  const M2 i1 = inv2(m1);  // m1.llt().solve(m2).trace() = tr(m1^-1 m2)
  const double tr = i1.a[0] * m2.a[0] + i1.a[1] * m2.a[2] + i1.a[2] * m2.a[1] + i1.a[3] * m2.a[3];
  const double c = a * tr;
  const double d = a + b - c;
  double w = 0.5 * (2 * b - c) / d;
  if ((w < 0 && d < 0) || (w > 1 && d > 0)) w = 0.0;
  else if ((w < 0 && d > 0) || (w > 1 && d < 0)) w = 1.0;
  M2 r;
  for (int i = 0; i < 4; ++i) r.a[i] = w * m1.a[i] + (1.0 - w) * m2.a[i];
  return r;
}

// VirtualMap.cpp:213-229 predictVirtualLandmark
static inline bool predict_virtual_landmark(const Config &cfg, const Pose &p, const M3 &info, double cx, double cy, M2 &out) {
  double b, r, Hx[6], Hl[4];
  predict_br(p, cx, cy, b, r, Hx, Hl);
  // BearingRangeSensorModel::check  Simulator2D.cpp:100-105
  if (!(b < cfg.max_bearing && b > cfg.min_bearing && r < cfg.max_range && r > cfg.min_range)) return false;
  const double R0 = cfg.bearing_noise * cfg.bearing_noise, R1 = cfg.range_noise * cfg.range_noise;
  // Hl <- (Hl^T Hl)^-1 Hl^T   (general inverse; LU-based .inverse() in Eigen, closed form 2x2 here)
  M2 HtH{{Hl[0] * Hl[0] + Hl[2] * Hl[2], Hl[0] * Hl[1] + Hl[2] * Hl[3],
          Hl[1] * Hl[0] + Hl[3] * Hl[2], Hl[1] * Hl[1] + Hl[3] * Hl[3]}};
  const double dd = det2(HtH);
  M2 Hi{{HtH.a[3] / dd, -HtH.a[1] / dd, -HtH.a[2] / dd, HtH.a[0] / dd}};
  double G[4] = {Hi.a[0] * Hl[0] + Hi.a[1] * Hl[1], Hi.a[0] * Hl[2] + Hi.a[1] * Hl[3],
                 Hi.a[2] * Hl[0] + Hi.a[3] * Hl[1], Hi.a[2] * Hl[2] + Hi.a[3] * Hl[3]};
  // S = R + Hx * info^-1 * Hx^T   (state.information.llt().solve(Hx^T))
  const M3 Sx = inv3(info);
  double HS[6];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) HS[i * 3 + j] = Hx[i * 3 + 0] * Sx.a[0 * 3 + j] + Hx[i * 3 + 1] * Sx.a[1 * 3 + j] + Hx[i * 3 + 2] * Sx.a[2 * 3 + j];
  double S[4];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) S[i * 2 + j] = HS[i * 3 + 0] * Hx[j * 3 + 0] + HS[i * 3 + 1] * Hx[j * 3 + 1] + HS[i * 3 + 2] * Hx[j * 3 + 2];
  S[0] += R0; S[3] += R1;
  double GS[4] = {G[0] * S[0] + G[1] * S[2], G[0] * S[1] + G[1] * S[3], G[2] * S[0] + G[3] * S[2], G[2] * S[1] + G[3] * S[3]};
  M2 cov{{GS[0] * G[0] + GS[1] * G[1], GS[0] * G[2] + GS[1] * G[3], GS[2] * G[0] + GS[3] * G[1], GS[2] * G[2] + GS[3] * G[3]}};
  // symmetrise round-off like an SPD llt() would only read one triangle
  cov.a[2] = cov.a[1];
  out = inv2(cov);
  return true;
}

void virtual_map_rebuild(const Config &cfg, int T, const double *pose, const double *info, int L, const double *lm,
                         int rows, int cols, double *prob, double *vinfo, int32_t *seen) {
  const int V = rows * cols;
  const double res = cfg.resolution, min_x = cfg.map_min_x, min_y = cfg.map_min_y;
  // ---- a6: OccupancyMap::update(map)  OccupancyMap.cpp:122-138 ----
  std::vector<double> lo(V, 0.0);  // LOGODDS_UNKNOWN = log(1) = 0
  std::vector<int32_t> cnt(V, 0);
  auto upd = [&](int row, int col, bool free_) {  // OccupancyMap.cpp:55-62
    if (row >= rows || row < 0 || col >= cols || col < 0) return;
    double l = lo[row * cols + col] + (free_ ? LOGODDS_FREE : LOGODDS_OCCUPIED);
    l = std::min(MAX_LOGODDS, std::max(MIN_LOGODDS, l));
    lo[row * cols + col] = l;
  };
  for (int j = 0; j < L; ++j) {
    const int orow = static_cast<int>(std::floor((lm[2 * j + 1] - min_y) / res));
    const int ocol = static_cast<int>(std::floor((lm[2 * j + 0] - min_x) / res));
    upd(orow, ocol, false);
    if (!(orow >= rows || orow < 0 || ocol >= cols || ocol < 0)) cnt[orow * cols + ocol] = -1;
  }
  for (int k = 0; k < T; ++k) {  // OccupancyMap.cpp:64-120
    const Pose p{pose[3 * k], pose[3 * k + 1], pose[3 * k + 2]};
    const int origin_row = static_cast<int>(std::floor((p.y - min_y) / res));
    const int origin_col = static_cast<int>(std::floor((p.x - min_x) / res));
    int min_row = std::min(std::max(0, origin_row), rows - 1), max_row = min_row;
    int min_col = std::min(std::max(0, origin_col), cols - 1), max_col = min_col;
    for (double b = cfg.min_bearing; b < cfg.max_bearing + 1e-5; b += ORC_DEG2RAD(3)) {
      const double x = p.x + cfg.max_range * std::cos(p.th + b);
      const double y = p.y + cfg.max_range * std::sin(p.th + b);
      const int row = std::min(std::max(0, static_cast<int>(std::floor((y - min_y) / res))), rows - 1);
      const int col = std::min(std::max(0, static_cast<int>(std::floor((x - min_x) / res))), cols - 1);
      min_row = std::min(min_row, row); max_row = std::max(max_row, row);
      min_col = std::min(min_col, col); max_col = std::max(max_col, col);
    }
    for (int row = min_row; row <= max_row; ++row)
      for (int col = min_col; col <= max_col; ++col) {
        const bool saturated = std::fabs(lo[row * cols + col] - MIN_LOGODDS) < 1e-5;
        const double x = min_x + res * (col + 0.5), y = min_y + res * (row + 0.5);
        double b, r;
        predict_br(p, x, y, b, r, nullptr, nullptr);
        // checkWithoutMinRange  Simulator2D.cpp:107-111
        if (!(b < cfg.max_bearing && b > cfg.min_bearing && r < cfg.max_range)) continue;
        if (cnt[row * cols + col] >= 0) cnt[row * cols + col]++;
        if (saturated) continue;  // OccupancyMap.cpp:104-105 ("speed up"; no effect on the value)
        if (lo[row * cols + col] > OCCUPIED_THRESH + 1e-8) upd(row, col, false);
        else upd(row, col, true);
      }
  }
  for (int i = 0; i < V; ++i) prob[i] = logodds2prob(lo[i]);  // VirtualMap.cpp:69-84 with num_samples = 1
  if (seen) std::memcpy(seen, cnt.data(), sizeof(int32_t) * V);

  // ---- a7: VirtualMap::updateInformation(map)  VirtualMap.cpp:256-271 ----
  std::vector<uint8_t> updated(V, 0);
  const double i0 = 1.0 / (cfg.sigma0 * cfg.sigma0);
  for (int i = 0; i < V; ++i) { vinfo[4 * i] = i0; vinfo[4 * i + 1] = 0; vinfo[4 * i + 2] = 0; vinfo[4 * i + 3] = i0; }
  for (int k = 0; k < T; ++k) {  // VirtualMap.cpp:290-316
    const Pose p{pose[3 * k], pose[3 * k + 1], pose[3 * k + 2]};
    M3 inf; std::memcpy(inf.a, info + 9 * k, sizeof(double) * 9);
    if (det3(inf) < 1e-10) continue;
    for (int n = 0; n < V; ++n) {  // KDTreeR2::queryRadiusNeighbors  Distance.cpp:78-97 (linear scan)
      const double cx = (n % cols + 0.5) * res + min_x, cy = (n / cols + 0.5) * res + min_y;  // VirtualMap.cpp:329-330
      const double dx = p.x - cx, dy = p.y - cy;
      if (!(std::sqrt(dx * dx + dy * dy) < cfg.max_range)) continue;
      M2 neu;
      if (!predict_virtual_landmark(cfg, p, inf, cx, cy, neu)) continue;
      M2 cur; std::memcpy(cur.a, vinfo + 4 * n, sizeof(double) * 4);
      if (updated[n]) cur = covariance_intersection(cur, neu);
      else { cur = neu; updated[n] = 1; }
      std::memcpy(vinfo + 4 * n, cur.a, sizeof(double) * 4);
    }
  }
}


// Iteration order of the reference's `std::unordered_map<unsigned, LandmarkBeliefState>` (Simulation2D.h:269) after the keys
// 0 .. n-1 were emplaced in order into a fresh map (Simulator2D.cpp:448-463).  The order decides which landmark consumes which
// noise draw in Simulator2D::measure (the kd-tree is filled by iterating the map, Simulator2D.cpp:335-340), so it is part of the
// arithmetic.  It depends on the libstdc++ that built the reference: the host's std::unordered_map (GCC 13: 7,6,..,0 for n = 8)
// does NOT reproduce the reference's result files; the hashtable of GCC 5 .. 7 (the compilers of the reference's era: Ubuntu
// 16.04 / 18.04) does -- every ordered pair the golden CSVs can confirm agrees with it (tests/test_oracle_cpu.py), e.g.
// 7,6,5,4,0,1,2,3 for n = 8.  That hashtable is restated here: singly linked node list + bucket table, insertion at the
// bucket's begin (`_M_insert_bucket_begin`), rehash by re-linking in list order (`_M_rehash_aux`, unique keys), and the prime
// rehash policy of those releases (`__n_elt + __n_ins >= _M_next_resize`, 12-entry fast table, growth factor 2).
std::vector<uint32_t> reference_unordered_map_order(int n) {
  static const unsigned long primes[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 103, 109, 113, 127,
                                         137, 139, 149, 157, 167, 179, 193, 199, 211, 227, 241, 257, 277, 293, 313, 337, 359, 383, 409, 439, 467, 503, 541,
                                         577, 619, 661, 709, 761, 823, 887, 953, 1031, 1109, 1193, 1289, 1381, 1493, 1613, 1741, 1879, 2029, 2179, 2357};
  static const unsigned char fast_bkt[12] = {2, 2, 2, 3, 5, 5, 7, 7, 11, 11, 11, 11};
  size_t next_resize = 0, n_bkt = 1;
  auto next_bkt = [&](size_t want) -> size_t {
    if (want <= 11) { next_resize = (size_t)std::ceil(fast_bkt[want] * 1.0L); return fast_bkt[want]; }
    const unsigned long *p = std::lower_bound(primes, primes + sizeof(primes) / sizeof(primes[0]), (unsigned long)want);
    next_resize = (size_t)std::ceil(*p * 1.0L);
    return *p;
  };
  const int HEAD = -1;                       // the list's before-begin node
  std::vector<int> nxt(n, -2);               // node = key
  int head_next = -2;                        // -2 = null
  std::vector<int> bucket(1, -3);            // bucket -> node BEFORE its first node (HEAD or a key), -3 = empty
  auto next_of = [&](int node) -> int & { return node == HEAD ? head_next : nxt[node]; };
  for (int key = 0; key < n; ++key) {
    const size_t n_elt = (size_t)key;
    if (n_elt + 1 >= next_resize) {          // _Prime_rehash_policy::_M_need_rehash
      const long double min_bkts = (long double)(n_elt + 1);
      if (min_bkts >= (long double)n_bkt) {
        const size_t nb = next_bkt(std::max<size_t>((size_t)std::floor(min_bkts) + 1, n_bkt * 2));
        std::vector<int> nbk(nb, -3);        // _M_rehash_aux(n, true_type)
        int p = head_next;
        head_next = -2;
        size_t bbegin_bkt = 0;
        while (p != -2) {
          const int pn = nxt[p];
          const size_t b = (size_t)p % nb;
          if (nbk[b] == -3) {
            nxt[p] = head_next; head_next = p; nbk[b] = HEAD;
            if (nxt[p] != -2) nbk[bbegin_bkt] = p;
            bbegin_bkt = b;
          } else {
            nxt[p] = next_of(nbk[b]); next_of(nbk[b]) = p;
          }
          p = pn;
        }
        bucket.swap(nbk);
        n_bkt = nb;
      } else {
        next_resize = (size_t)std::floor(n_bkt * 1.0L);
      }
    }
    const size_t b = (size_t)key % n_bkt;    // _M_insert_bucket_begin
    if (bucket[b] != -3) {
      nxt[key] = next_of(bucket[b]); next_of(bucket[b]) = key;
    } else {
      nxt[key] = head_next; head_next = key;
      if (nxt[key] != -2) bucket[(size_t)nxt[key] % n_bkt] = key;
      bucket[b] = HEAD;
    }
  }
  std::vector<uint32_t> order;
  for (int p = head_next; p != -2; p = nxt[p]) order.push_back((uint32_t)p);
  return order;
}

// ==================================================================== Env ===
Env::Env(const Config &c) : cfg(c) { setup_grid(); }

void Env::setup_grid() {  // VirtualMap.cpp:318-340
  cols = static_cast<int>(std::floor((cfg.map_max_x - cfg.map_min_x) / cfg.resolution));
  rows = static_cast<int>(std::floor((cfg.map_max_y - cfg.map_min_y) / cfg.resolution));
  prob.assign(rows * cols, 0.5);
  vinfo.assign(rows * cols, M2{{1.0 / (cfg.sigma0 * cfg.sigma0), 0, 0, 1.0 / (cfg.sigma0 * cfg.sigma0)}});
  seen_count.assign(rows * cols, 0);
}

void Env::add_true_landmarks(const std::vector<double> &xy) {
  Lt = static_cast<int>(xy.size() / 2);
  lm_x.resize(Lt); lm_y.resize(Lt);
  for (int i = 0; i < Lt; ++i) { lm_x[i] = xy[2 * i]; lm_y[i] = xy[2 * i + 1]; }
  scan_id = reference_unordered_map_order(Lt);  // Simulation2D.h:269, Simulator2D.cpp:335-340 ; q6
  observed.assign(Lt, 0);
  lin_l.assign(2 * Lt, 0); est_l.assign(2 * Lt, 0); delta_l.assign(2 * Lt, 0);
  land_cov.assign(Lt, M2{{0, 0, 0, 0}}); land_info.assign(Lt, M2{{0, 0, 0, 0}});
}

void Env::init(uint32_t seed, Pose start, StepNoise *rec) {
  // Simulator2D ctor: three generators, same seed  Simulator2D.cpp:436-443
  rng_sim = Rng(seed);
  // Simulator2D::addLandmarks  Simulator2D.cpp:445-465
  std::vector<double> xy;
  for (int i = 0; i < cfg.num_landmarks;) {
    const double x = rng_sim.uniformReal(cfg.env_min_x, cfg.env_max_x);
    const double y = rng_sim.uniformReal(cfg.env_min_y, cfg.env_max_y);
    const double dx = x - start.x, dy = y - start.y;
    if (std::sqrt(dx * dx + dy * dy) < 2.0) continue;
    xy.push_back(x); xy.push_back(y);
    ++i;
  }
  init_with_landmarks(seed, start, xy, rec);
}

void Env::init_with_landmarks(uint32_t seed, Pose start, const std::vector<double> &xy, StepNoise *rec) {
  rng_sensor = Rng(seed); rng_control = Rng(seed);
  if (xy.size() / 2 != static_cast<size_t>(cfg.num_landmarks)) cfg.num_landmarks = static_cast<int>(xy.size() / 2);
  add_true_landmarks(xy);
  setup_grid();
  start.th = wrap_pi(start.th);  // Pose2(x0, y0, theta0) keeps (cos, sin); theta() is normalised
  true_pose = start;
  // SLAM2D::addPrior  SLAM2D.cpp:44-57 ; pyss2d.py:127-133
  T = 1;
  prior_pose = start;
  lin_pose.assign(1, start); est_pose.assign(1, start);
  delta_pose.assign(3, 0.0);
  odom.clear(); meas.clear(); meas_ptr.assign(1, 0);
  pose_cov.clear(); pose_info.clear();
  update_count = 0; sim_step = 0; dist = 0;
  if (rec) rec->v.assign(3 + 4 * Lt, 0.0);
  std::vector<Meas> ms;
  measure(ms, rec, 1);          // pyss2d.py:134 self.measure()
  add_measurements(ms);
  meas_ptr.push_back(static_cast<int32_t>(meas.size()));
  slam_optimize();              // pyss2d.py:135
  sim_step = 1;                 // pyss2d.py:136
}

// Simulator2D::move  Simulator2D.cpp:491-503 -> SimpleControlModel::evolve :161-182 ;
// SLAM2D::addOdometry  SLAM2D.cpp:70-89
void Env::move(const double o[3], StepNoise *rec) {
  // forced_noise (tests only): explicit noise in the StepNoise layout instead of the mt19937 streams
  const double nx = forced_noise ? forced_noise[0] : rng_control.normal(0.0, cfg.trans_noise);
  const double ny = forced_noise ? forced_noise[1] : rng_control.normal(0.0, cfg.trans_noise);
  const double nt = forced_noise ? forced_noise[2] : rng_control.normal(0.0, cfg.rot_noise);
  if (rec) { rec->v[0] = nx; rec->v[1] = ny; rec->v[2] = nt; }
  const Pose od{o[0], o[1], o[2]};
  true_pose = compose(compose(true_pose, od), Pose{nx, ny, nt});
  odom.push_back(o[0]); odom.push_back(o[1]); odom.push_back(o[2]);
  const Pose p2 = compose(est_pose[T - 1], od);  // SLAM2D.cpp:80-88 (result_ holds the latest estimate)
  lin_pose.push_back(p2); est_pose.push_back(p2);
  delta_pose.insert(delta_pose.end(), 3, 0.0);
  ++T;
}

// Simulator2D::measure  Simulator2D.cpp:505-527
void Env::measure(std::vector<Meas> &out, StepNoise *rec, int call) {
  out.clear();
  for (int s = 0; s < Lt; ++s) {
    const uint32_t id = scan_id[s];
    const double dx = true_pose.x - lm_x[id], dy = true_pose.y - lm_y[id];
    if (!(std::sqrt(dx * dx + dy * dy) < cfg.max_range)) continue;   // Distance.cpp:86-88
    const double nb = forced_noise ? forced_noise[3 + call * 2 * Lt + 2 * s] : rng_sensor.normal(0.0, cfg.bearing_noise);     // Simulator2D.cpp:116-117
    const double nr = forced_noise ? forced_noise[3 + call * 2 * Lt + 2 * s + 1] : rng_sensor.normal(0.0, cfg.range_noise);
    if (rec) { rec->v[3 + call * 2 * Lt + 2 * s] = nb; rec->v[3 + call * 2 * Lt + 2 * s + 1] = nr; }
    double b, r;
    predict_br(true_pose, lm_x[id], lm_y[id], b, r, nullptr, nullptr);
    b += nb; r += nr;
    if (b < cfg.max_bearing && b > cfg.min_bearing && r < cfg.max_range && r > cfg.min_range)  // :100-105
      out.push_back(Meas{static_cast<int32_t>(id), b, r});
  }
}

// SLAM2D::addMeasurement  SLAM2D.cpp:103-124
void Env::add_measurements(const std::vector<Meas> &ms) {
  for (const Meas &m : ms) {
    meas.push_back(m);
    if (!observed[m.id]) {
      const Pose &o = est_pose[T - 1];  // initial estimate of the newest pose
      const double qx = m.range * std::cos(m.bearing), qy = m.range * std::sin(m.bearing);  // Simulator2D.cpp:95-98
      const double c = std::cos(o.th), s = std::sin(o.th);
      const double gx = o.x + c * qx - s * qy, gy = o.y + s * qx + c * qy;
      observed[m.id] = 1;
      lin_l[2 * m.id] = est_l[2 * m.id] = gx;
      lin_l[2 * m.id + 1] = est_l[2 * m.id + 1] = gy;
      delta_l[2 * m.id] = delta_l[2 * m.id + 1] = 0.0;
    }
  }
}

bool Env::simulate(const double o[3], StepNoise *rec) {
  // pyss2d.py:173-176 (q3: the bounds test is on the odom vector itself)
  if (!(cfg.map_min_x < o[0] && o[0] < cfg.map_max_x) || !(cfg.map_min_y < o[1] && o[1] < cfg.map_max_y)) return false;
  if (rec) rec->v.assign(3 + 4 * Lt, 0.0);
  move(o, rec);
  std::vector<Meas> ms;
  measure(ms, rec, 0);           // pyss2d.py:182 obstacle probe (q4) -- burns sensor draws
  ++sim_step;
  measure(ms, rec, 1);           // pyss2d.py:203
  add_measurements(ms);
  meas_ptr.push_back(static_cast<int32_t>(meas.size()));
  slam_optimize();               // pyss2d.py:204
  update_virtual_map();          // pyss2d.py:205
  return true;
}

void Env::step(const double o[3], StepNoise *rec) {  // exploration_env.py:98-105 (q17)
  simulate(o, rec);
  dist += std::sqrt(o[0] * o[0] + o[1] * o[1]);
}

int Env::n_observed() const { int n = 0; for (uint8_t v : observed) n += v; return n; }

// ------------------------------------------------------------ SLAM solve ---
// SLAM2D::optimize (SLAM2D.cpp:374-430) with gtsam::ISAM2 (default ISAM2Params,
// SLAM2D.cpp:10-12) replaced by its documented schedule: fixed linearisation point
// theta, relinearise every `relin_skip`-th update the variables whose |delta|_inf >=
// relin_thresh, exact solve of the normal equations, estimate = theta (+) delta,
// marginals = diagonal blocks of (J^T J)^-1 at theta (FastMarginals.cpp:171-186).
void Env::slam_optimize() {
  ++update_count;
  std::vector<uint8_t> relin_p(T, 0), relin_l(Lt, 0);   // (read by the default-off experiment below only)
  if (cfg.relin_skip > 0 && update_count % cfg.relin_skip == 0) {
    for (int k = 0; k < T; ++k) {
      const double *d = &delta_pose[3 * k];
      if (std::max(std::fabs(d[0]), std::max(std::fabs(d[1]), std::fabs(d[2]))) >= cfg.relin_thresh) {
        lin_pose[k] = retract(lin_pose[k], d);
        delta_pose[3 * k] = delta_pose[3 * k + 1] = delta_pose[3 * k + 2] = 0;
        relin_p[k] = 1;
      }
    }
    for (int j = 0; j < Lt; ++j) {
      if (!observed[j]) continue;
      if (std::max(std::fabs(delta_l[2 * j]), std::fabs(delta_l[2 * j + 1])) >= cfg.relin_thresh) {
        lin_l[2 * j] += delta_l[2 * j]; lin_l[2 * j + 1] += delta_l[2 * j + 1];
        delta_l[2 * j] = delta_l[2 * j + 1] = 0;
        relin_l[j] = 1;
      }
    }
  }
  std::vector<int> lidx(Lt, -1);
  int nl = 0;
  for (int j = 0; j < Lt; ++j) if (observed[j]) lidx[j] = nl++;
  pose_cov.resize(T); pose_info.resize(T);
  // EXPERIMENT (off unless DGE_ORACLE_WILDFIRE is set; not part of the pinned oracle, never used by the tests): ISAM2 does not
  // hand every variable its exact delta -- cliques that were not re-eliminated keep their OLD delta unless it moved by at least
  // wildfireThreshold (gtsam ISAM2-impl optimizeWildfireNode).  Without the Bayes tree the re-eliminated set is approximated by the
  // pose range [k0, t] -- k0 = the LAST earlier observation of any landmark seen at x_t / x_{t-1} (the path from that landmark's
  // clique to the root), or the pose of the oldest relinearised variable -- plus the landmarks seen from that range; if it holds
  // >= 65 % of all variables ISAM2 re-eliminates everything (its batchThreshold).  Outside it a variable keeps its old delta unless
  // it moved by >= the threshold.  Effect on the policy-free yardstick (tests/golden/scan_guided.py, 1000 episodes): 24 363 rows
  // against 24 146 of the exact solve -- a small gain; cruder variants of the set LOSE 2 700 - 5 300 rows.
  static const char *wf_env = std::getenv("DGE_ORACLE_WILDFIRE");
  const double wf = wf_env ? std::atof(wf_env) : 0.0;
  std::vector<double> old_dp, old_dl;
  if (wf > 0) { old_dp = delta_pose; old_dl = delta_l; }
  if (use_dense_solver) solve_dense(lidx, nl); else solve_structured(lidx, nl);
  if (wf > 0 && T >= 2) {
    std::vector<uint8_t> lmA(Lt, 0);
    for (int k = std::max(0, T - 2); k < T; ++k)
      for (int p = meas_ptr[k]; p < meas_ptr[k + 1]; ++p) lmA[meas[p].id] = 1;
    for (int j = 0; j < Lt; ++j) if (relin_l[j]) lmA[j] = 1;
    int oldest_relin = T;
    for (int k = T - 1; k >= 0; --k) if (relin_p[k]) oldest_relin = k;
    int k0 = T - 2;
    for (int k = T - 3; k >= 0; --k) {
      bool hit = relin_p[k];
      for (int p = meas_ptr[k]; p < meas_ptr[k + 1] && !hit; ++p) hit = lmA[meas[p].id];
      if (hit) { k0 = k; if (oldest_relin >= k) break; }
    }
    std::vector<uint8_t> inA(Lt, 0);
    for (int k = std::max(0, k0); k < T; ++k)
      for (int p = meas_ptr[k]; p < meas_ptr[k + 1]; ++p) inA[meas[p].id] = 1;
    for (int j = 0; j < Lt; ++j) if (relin_l[j]) inA[j] = 1;
    int nA = T - std::max(0, k0), nAll = T;
    for (int j = 0; j < Lt; ++j) if (observed[j]) { ++nAll; nA += inA[j]; }
    if (nA < 0.65 * nAll) {
      for (int k = 0; k < k0 && 3 * k + 2 < (int)old_dp.size(); ++k) {
        double ch = 0;
        for (int i = 0; i < 3; ++i) ch = std::max(ch, std::fabs(delta_pose[3 * k + i] - old_dp[3 * k + i]));
        if (ch < wf) for (int i = 0; i < 3; ++i) delta_pose[3 * k + i] = old_dp[3 * k + i];
      }
      for (int j = 0; j < Lt; ++j) {
        if (!observed[j] || inA[j]) continue;
        const double ch = std::max(std::fabs(delta_l[2 * j] - old_dl[2 * j]), std::fabs(delta_l[2 * j + 1] - old_dl[2 * j + 1]));
        if (ch < wf) { delta_l[2 * j] = old_dl[2 * j]; delta_l[2 * j + 1] = old_dl[2 * j + 1]; }
      }
    }
  }
  for (int k = 0; k < T; ++k) {
    est_pose[k] = retract(lin_pose[k], &delta_pose[3 * k]);
    pose_info[k] = inv3(pose_cov[k]);   // SLAM2D.cpp:400  inverse(covariance)
  }
  for (int j = 0; j < Lt; ++j) {
    if (!observed[j]) continue;
    est_l[2 * j] = lin_l[2 * j] + delta_l[2 * j];
    est_l[2 * j + 1] = lin_l[2 * j + 1] + delta_l[2 * j + 1];
    const M2 &c = land_cov[j];
    const double dd = det2(c);           // SLAM2D.cpp:420  marginalCovariance(l).inverse()  (general inverse)
    land_info[j] = M2{{c.a[3] / dd, -c.a[1] / dd, -c.a[2] / dd, c.a[0] / dd}};
  }
}

namespace {
// whitened linearisation of every factor attached to pose k, evaluated at theta.
struct PoseLin {
  double D[9];    // own diagonal block contribution
  double U[9];    // block (k, k+1)
  double g[3];
  double Dn[9];   // contribution of the odometry factor k->k+1 to D_{k+1}
  double gn[3];   // ... and to g_{k+1}
};
}

// Block-tridiagonal (poses) + arrow (landmarks) elimination; SURVEY section 10 "Structure
// for the fast path".  Poses first (chain order), Schur complement on the landmarks,
// Takahashi / Kaess-Dellaert recursion (what FastMarginals::recover computes) backwards.
void Env::solve_structured(const std::vector<int> &lidx, int nl) {
  const int n2 = 2 * nl;
  const double wo[3] = {1.0 / (cfg.trans_noise * cfg.trans_noise), 1.0 / (cfg.trans_noise * cfg.trans_noise), 1.0 / (cfg.rot_noise * cfg.rot_noise)};
  const double wm[2] = {1.0 / (cfg.bearing_noise * cfg.bearing_noise), 1.0 / (cfg.range_noise * cfg.range_noise)};
  const double wp[3] = {1.0 / (cfg.sigma_x0 * cfg.sigma_x0), 1.0 / (cfg.sigma_y0 * cfg.sigma_y0), 1.0 / (cfg.sigma_theta0 * cfg.sigma_theta0)};

  std::vector<double> S(n2 * n2, 0.0), gl(n2, 0.0);
  std::vector<double> FU(T * 9), FB(static_cast<size_t>(T) * 3 * n2), ff(T * 3), Dinv(T * 9);
  std::vector<double> Bt(3 * n2), carryB(3 * n2, 0.0);
  double carryD[9] = {0}, carryg[3] = {0};
  double Dnext[9] = {0}, gnext[3] = {0};

  for (int k = 0; k < T; ++k) {
    double D[9], g[3], U[9] = {0};
    for (int i = 0; i < 9; ++i) D[i] = Dnext[i] + carryD[i];
    for (int i = 0; i < 3; ++i) g[i] = gnext[i] + carryg[i];
    for (int i = 0; i < 9; ++i) Dnext[i] = 0;
    for (int i = 0; i < 3; ++i) gnext[i] = 0;
    std::copy(carryB.begin(), carryB.end(), Bt.begin());
    if (k == 0) {  // PriorFactor<Pose2>: error = -Local(x, prior), H = I  (gtsam 4.0 PriorFactor.h)
      const Pose e = between(lin_pose[0], prior_pose, nullptr);
      const double r[3] = {-e.x, -e.y, -e.th};
      for (int i = 0; i < 3; ++i) { D[i * 3 + i] += wp[i]; g[i] -= wp[i] * r[i]; }
    }
    if (k + 1 < T) {  // BetweenFactor<Pose2>(x_k, x_k+1, odom): error = Local(odom, between), J = [H1 | I]
      double H1[9];
      const Pose h = between(lin_pose[k], lin_pose[k + 1], H1);
      const Pose od{odom[3 * k], odom[3 * k + 1], odom[3 * k + 2]};
      const Pose e = between(od, h, nullptr);
      const double r[3] = {e.x, e.y, e.th};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double s = 0;
          for (int m = 0; m < 3; ++m) s += H1[m * 3 + i] * wo[m] * H1[m * 3 + j];
          D[i * 3 + j] += s;
          U[i * 3 + j] = H1[j * 3 + i] * wo[j];
        }
      for (int i = 0; i < 3; ++i) {
        double s = 0;
        for (int m = 0; m < 3; ++m) s += H1[m * 3 + i] * wo[m] * r[m];
        g[i] -= s;
        Dnext[i * 3 + i] = wo[i];
        gnext[i] = -wo[i] * r[i];
      }
    }
    for (int p = meas_ptr[k]; p < meas_ptr[k + 1]; ++p) {  // BearingRangeFactor<Pose2,Point2>
      const Meas &m = meas[p];
      const int c0 = 2 * lidx[m.id];
      double b, rg, Hx[6], Hl[4];
      predict_br(lin_pose[k], lin_l[2 * m.id], lin_l[2 * m.id + 1], b, rg, Hx, Hl);
      const double r[2] = {wrap_pi(b - m.bearing), rg - m.range};
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) D[i * 3 + j] += Hx[i] * wm[0] * Hx[j] + Hx[3 + i] * wm[1] * Hx[3 + j];
        for (int j = 0; j < 2; ++j) Bt[i * n2 + c0 + j] += Hx[i] * wm[0] * Hl[j] + Hx[3 + i] * wm[1] * Hl[2 + j];
        g[i] -= Hx[i] * wm[0] * r[0] + Hx[3 + i] * wm[1] * r[1];
      }
      for (int i = 0; i < 2; ++i) {
        for (int j = 0; j < 2; ++j) S[(c0 + i) * n2 + c0 + j] += Hl[i] * wm[0] * Hl[j] + Hl[2 + i] * wm[1] * Hl[2 + j];
        gl[c0 + i] -= Hl[i] * wm[0] * r[0] + Hl[2 + i] * wm[1] * r[1];
      }
    }
    double Di[9];
    if (!chol_inv(3, D, Di)) throw std::runtime_error("oracle: pose block not SPD");
    std::copy(Di, Di + 9, &Dinv[9 * k]);
    double *fu = &FU[9 * k], *fb = &FB[static_cast<size_t>(k) * 3 * n2], *f = &ff[3 * k];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) fu[i * 3 + j] = Di[i * 3] * U[j] + Di[i * 3 + 1] * U[3 + j] + Di[i * 3 + 2] * U[6 + j];
      for (int c = 0; c < n2; ++c) fb[i * n2 + c] = Di[i * 3] * Bt[c] + Di[i * 3 + 1] * Bt[n2 + c] + Di[i * 3 + 2] * Bt[2 * n2 + c];
      f[i] = Di[i * 3] * g[0] + Di[i * 3 + 1] * g[1] + Di[i * 3 + 2] * g[2];
    }
    for (int r = 0; r < n2; ++r) {
      for (int c = 0; c < n2; ++c) S[r * n2 + c] -= Bt[r] * fb[c] + Bt[n2 + r] * fb[n2 + c] + Bt[2 * n2 + r] * fb[2 * n2 + c];
      gl[r] -= Bt[r] * f[0] + Bt[n2 + r] * f[1] + Bt[2 * n2 + r] * f[2];
    }
    for (int i = 0; i < 3; ++i) {  // carry = -U^T F
      for (int j = 0; j < 3; ++j) carryD[i * 3 + j] = -(U[i] * fu[j] + U[3 + i] * fu[3 + j] + U[6 + i] * fu[6 + j]);
      for (int c = 0; c < n2; ++c) carryB[i * n2 + c] = -(U[i] * fb[c] + U[3 + i] * fb[n2 + c] + U[6 + i] * fb[2 * n2 + c]);
      carryg[i] = -(U[i] * f[0] + U[3 + i] * f[1] + U[6 + i] * f[2]);
    }
  }
  std::vector<double> Sll(n2 * n2, 0.0), dl(n2, 0.0);
  if (n2 > 0) {
    if (!chol_inv(n2, S.data(), Sll.data())) throw std::runtime_error("oracle: landmark Schur complement not SPD");
    for (int r = 0; r < n2; ++r) { double s = 0; for (int c = 0; c < n2; ++c) s += Sll[r * n2 + c] * gl[c]; dl[r] = s; }
  }
  for (int j = 0; j < Lt; ++j) {
    if (lidx[j] < 0) continue;
    const int c0 = 2 * lidx[j];
    delta_l[2 * j] = dl[c0]; delta_l[2 * j + 1] = dl[c0 + 1];
    land_cov[j] = M2{{Sll[c0 * n2 + c0], Sll[c0 * n2 + c0 + 1], Sll[(c0 + 1) * n2 + c0], Sll[(c0 + 1) * n2 + c0 + 1]}};
  }
  // backward: delta_k = f_k - FU_k delta_{k+1} - FB_k delta_l ; marginals by the recursion
  std::vector<double> Skl(3 * n2, 0.0), Skl_next(3 * n2, 0.0);
  double Skk_next[9] = {0}, dnext[3] = {0};
  for (int k = T - 1; k >= 0; --k) {
    const double *fu = &FU[9 * k], *fb = &FB[static_cast<size_t>(k) * 3 * n2], *f = &ff[3 * k], *Di = &Dinv[9 * k];
    double d[3];
    for (int i = 0; i < 3; ++i) {
      double s = f[i] - (fu[i * 3] * dnext[0] + fu[i * 3 + 1] * dnext[1] + fu[i * 3 + 2] * dnext[2]);
      for (int c = 0; c < n2; ++c) s -= fb[i * n2 + c] * dl[c];
      d[i] = s;
    }
    // Sigma_{k,l} = -FU Sigma_{k+1,l} - FB Sigma_ll
    for (int i = 0; i < 3; ++i)
      for (int c = 0; c < n2; ++c) {
        double s = -(fu[i * 3] * Skl_next[c] + fu[i * 3 + 1] * Skl_next[n2 + c] + fu[i * 3 + 2] * Skl_next[2 * n2 + c]);
        for (int r = 0; r < n2; ++r) s -= fb[i * n2 + r] * Sll[r * n2 + c];
        Skl[i * n2 + c] = s;
      }
    // Sigma_{k,k+1} = -FU Sigma_{k+1,k+1} - FB Sigma_{l,k+1}
    double Skn[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = -(fu[i * 3] * Skk_next[j] + fu[i * 3 + 1] * Skk_next[3 + j] + fu[i * 3 + 2] * Skk_next[6 + j]);
        for (int r = 0; r < n2; ++r) s -= fb[i * n2 + r] * Skl_next[j * n2 + r];
        Skn[i * 3 + j] = s;
      }
    // Sigma_kk = Dinv - Sigma_{k,k+1} FU^T - Sigma_{k,l} FB^T
    double Skk[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = Di[i * 3 + j] - (Skn[i * 3] * fu[j * 3] + Skn[i * 3 + 1] * fu[j * 3 + 1] + Skn[i * 3 + 2] * fu[j * 3 + 2]);
        for (int c = 0; c < n2; ++c) s -= Skl[i * n2 + c] * fb[j * n2 + c];
        Skk[i * 3 + j] = s;
      }
    for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 3; ++j) { const double m = 0.5 * (Skk[i * 3 + j] + Skk[j * 3 + i]); Skk[i * 3 + j] = Skk[j * 3 + i] = m; }
    std::copy(Skk, Skk + 9, pose_cov[k].a);
    std::copy(d, d + 3, &delta_pose[3 * k]);
    std::copy(Skk, Skk + 9, Skk_next);
    std::copy(d, d + 3, dnext);
    std::swap(Skl, Skl_next);
  }
}

// Independent check of solve_structured: assemble the full normal equations densely
// and invert with one Cholesky.  O(n^3); tests only.
void Env::solve_dense(const std::vector<int> &lidx, int nl) {
  const int n = 3 * T + 2 * nl;
  std::vector<double> A(static_cast<size_t>(n) * n, 0.0), b(n, 0.0);
  const double so[3] = {cfg.trans_noise, cfg.trans_noise, cfg.rot_noise};
  const double sm[2] = {cfg.bearing_noise, cfg.range_noise};
  const double sp[3] = {cfg.sigma_x0, cfg.sigma_y0, cfg.sigma_theta0};
  auto add_row = [&](const std::vector<std::pair<int, double>> &row, double r) {  // one whitened Jacobian row
    for (auto &a : row) {
      for (auto &c : row) A[static_cast<size_t>(a.first) * n + c.first] += a.second * c.second;
      b[a.first] -= a.second * r;
    }
  };
  {
    const Pose e = between(lin_pose[0], prior_pose, nullptr);
    const double r[3] = {-e.x, -e.y, -e.th};
    for (int i = 0; i < 3; ++i) add_row({{i, 1.0 / sp[i]}}, r[i] / sp[i]);
  }
  for (int k = 0; k + 1 < T; ++k) {
    double H1[9];
    const Pose h = between(lin_pose[k], lin_pose[k + 1], H1);
    const Pose e = between(Pose{odom[3 * k], odom[3 * k + 1], odom[3 * k + 2]}, h, nullptr);
    const double r[3] = {e.x, e.y, e.th};
    for (int i = 0; i < 3; ++i) {
      std::vector<std::pair<int, double>> row;
      for (int j = 0; j < 3; ++j) row.push_back({3 * k + j, H1[i * 3 + j] / so[i]});
      row.push_back({3 * (k + 1) + i, 1.0 / so[i]});
      add_row(row, r[i] / so[i]);
    }
  }
  for (int k = 0; k < T; ++k)
    for (int p = meas_ptr[k]; p < meas_ptr[k + 1]; ++p) {
      const Meas &m = meas[p];
      const int c0 = 3 * T + 2 * lidx[m.id];
      double bb, rg, Hx[6], Hl[4];
      predict_br(lin_pose[k], lin_l[2 * m.id], lin_l[2 * m.id + 1], bb, rg, Hx, Hl);
      const double r[2] = {wrap_pi(bb - m.bearing), rg - m.range};
      for (int i = 0; i < 2; ++i) {
        std::vector<std::pair<int, double>> row;
        for (int j = 0; j < 3; ++j) row.push_back({3 * k + j, Hx[i * 3 + j] / sm[i]});
        for (int j = 0; j < 2; ++j) row.push_back({c0 + j, Hl[i * 2 + j] / sm[i]});
        add_row(row, r[i] / sm[i]);
      }
    }
  std::vector<double> Ai(static_cast<size_t>(n) * n);
  if (!chol_inv(n, A.data(), Ai.data())) throw std::runtime_error("oracle: dense information matrix not SPD");
  std::vector<double> d(n, 0.0);
  for (int i = 0; i < n; ++i) { double s = 0; for (int j = 0; j < n; ++j) s += Ai[static_cast<size_t>(i) * n + j] * b[j]; d[i] = s; }
  for (int k = 0; k < T; ++k) {
    for (int i = 0; i < 3; ++i) {
      delta_pose[3 * k + i] = d[3 * k + i];
      for (int j = 0; j < 3; ++j) pose_cov[k].a[i * 3 + j] = Ai[static_cast<size_t>(3 * k + i) * n + 3 * k + j];
    }
  }
  for (int j = 0; j < Lt; ++j) {
    if (lidx[j] < 0) continue;
    const int c0 = 3 * T + 2 * lidx[j];
    delta_l[2 * j] = d[c0]; delta_l[2 * j + 1] = d[c0 + 1];
    land_cov[j] = M2{{Ai[static_cast<size_t>(c0) * n + c0], Ai[static_cast<size_t>(c0) * n + c0 + 1],
                      Ai[static_cast<size_t>(c0 + 1) * n + c0], Ai[static_cast<size_t>(c0 + 1) * n + c0 + 1]}};
  }
}

// ------------------------------------------------------------ virtual map ---
void Env::update_virtual_map() {
  std::vector<double> pose(3 * T), info(9 * T), lm;
  for (int k = 0; k < T; ++k) {
    pose[3 * k] = est_pose[k].x; pose[3 * k + 1] = est_pose[k].y; pose[3 * k + 2] = est_pose[k].th;
    std::memcpy(&info[9 * k], pose_info[k].a, sizeof(double) * 9);
  }
  // map_.landmarks_ : unordered_map, order irrelevant for the occupancy result (q8)
  for (int j = 0; j < Lt; ++j) if (observed[j]) { lm.push_back(est_l[2 * j]); lm.push_back(est_l[2 * j + 1]); }
  virtual_map_rebuild(cfg, T, pose.data(), info.data(), static_cast<int>(lm.size() / 2), lm.data(), rows, cols,
                      prob.data(), reinterpret_cast<double *>(vinfo.data()), seen_count.data());
}

void Env::cov_trace(std::vector<double> &out) const {  // VirtualMap.cpp:153-159
  out.resize(rows * cols);
  for (int i = 0; i < rows * cols; ++i) { const M2 c = inv2(vinfo[i]); out[i] = c.a[0] + c.a[3]; }
}

double Env::explored() const {  // VirtualMap.cpp:47-59 ; count_explored_ :341
  int count = 0;
  const int extg = 20;
  for (int i = 0; i < rows * cols; ++i) {
    const double x = (i % cols + 0.5) * cfg.resolution + cfg.map_min_x, y = (i / cols + 0.5) * cfg.resolution + cfg.map_min_y;
    if ((prob[i] < 0.49 || prob[i] > 0.6) && cfg.map_min_x + extg <= x && x <= cfg.map_max_x - extg &&
        cfg.map_min_y + extg <= y && y <= cfg.map_max_y - extg)
      ++count;
  }
  const int ce = (rows - extg * 2 / static_cast<int>(cfg.resolution)) * (cols - extg * 2 / static_cast<int>(cfg.resolution));
  return static_cast<double>(count) / ce;
}

double Env::utility(double distance) const {  // Planner2D.cpp:343-366
  int known = 0;
  double unc = 0.0;
  for (int i = 0; i < rows * cols; ++i) {
    if (prob[i] < cfg.occupancy_threshold) ++known;
    const M2 c = inv2(vinfo[i]);
    unc += 1.0 * (c.a[0] + c.a[3]);
  }
  const double pk = static_cast<double>(known) / (rows * cols);
  return unc + distance * (cfg.dist_w0 - (cfg.dist_w0 - cfg.dist_w1) * pk);
}

bool Env::done() const { return sim_step > cfg.max_steps || explored() > 0.85; }

double Env::landmark_error(double sigma0) const {  // exploration_env.py:170-176
  double err = 0;
  int nobs = 0;
  for (int j = 0; j < Lt; ++j) {
    if (!observed[j]) continue;
    ++nobs;
    const double dx = lm_x[j] - est_l[2 * j], dy = lm_y[j] - est_l[2 * j + 1];
    err += std::sqrt(dx * dx + dy * dy);
  }
  err += sigma0 * (Lt - nobs);
  return err / Lt;
}

double Env::max_traj_uncertainty() const {  // exploration_env.py:190-194 ; SLAM2D.cpp:235-266 (q13)
  double m = -std::numeric_limits<double>::infinity();
  for (int k = 0; k < T; ++k) m = std::max(m, pose_cov[k].a[0] + pose_cov[k].a[4] + pose_cov[k].a[8]);
  return m;
}

// ------------------------------------------------------- frontier + graph ---
static inline double points2dist(double ax, double ay, double bx, double by) {  // exploration_env.py:374-376
  return std::sqrt((ax - bx) * (ax - bx) + (ay - by) * (ay - by));
}
static inline double py_round(double v) { return std::nearbyint(v); }  // Python round(): half to even
static inline double diff_theta(double p1x, double p1y, double p2x, double p2y, double root) {  // :378-387
  double goal = std::atan2(p1y - p2y, p1x - p2x);
  if (goal < 0) goal = kPi * 2 + goal;
  if (root < 0) root = kPi * 2 + root;
  double diff = goal - root;
  if (diff < 0) diff = kPi * 2 + diff;
  return diff;
}

void Env::graph(GraphOut &g) const {
  const double res = cfg.resolution, ext = 20.0;
  const Pose &rob = est_pose[T - 1];
  // key order: 'l' ids ascending, then 'x' (SURVEY 3.3 ordering contract)
  std::vector<int> land_ids;
  for (int j = 0; j < Lt; ++j) if (observed[j]) land_ids.push_back(j);
  const int L = static_cast<int>(land_ids.size()), K = L + T;
  auto key_xy = [&](int i, double &x, double &y) {  // SLAM2D::get_key_points  SLAM2D.cpp:152-166
    if (i < L) { x = est_l[2 * land_ids[i]]; y = est_l[2 * land_ids[i] + 1]; }
    else { x = est_pose[i - L].x; y = est_pose[i - L].y; }
  };
  // ---- frontier()  exploration_env.py:289-348 ----
  std::vector<double> fx, fy;
  g.all_frontier_cells.clear();
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < cols; ++j) {
      if (!(prob[i * cols + j] < 0.45)) continue;
      const int i0 = std::max(i - 1, 0), i1 = std::min(i + 1, rows - 1), j0 = std::max(j - 1, 0), j1 = std::min(j + 1, cols - 1);
      int count = 0;
      for (int a = i0; a <= i1; ++a)
        for (int b = j0; b <= j1; ++b)
          if (0.49 < prob[a * cols + b] && prob[a * cols + b] < 0.51) ++count;
      if (count < 2) continue;
      const double x = (j + 0.5) * res + cfg.map_min_x, y = (i + 0.5) * res + cfg.map_min_y;  // index2coor :364-367
      if (cfg.map_min_x + ext <= x && x <= cfg.map_max_x - ext && cfg.map_min_y + ext <= y && y <= cfg.map_max_y - ext) {
        fx.push_back(x); fy.push_back(y); g.all_frontier_cells.push_back(i * cols + j);
      }
    }
  auto nearest = [&](double px, double py) {  // nearest_frontier :350-358 (strict <, first minimum)
    double md = std::numeric_limits<double>::infinity();
    int mi = -1;
    for (size_t q = 0; q < fx.size(); ++q) {
      const double d = points2dist(px, py, fx[q], fy[q]);
      if (d < md) { md = d; mi = static_cast<int>(q); }
    }
    return mi;
  };
  std::vector<int> fro;                    // index into all_frontiers
  std::vector<std::vector<int>> fro_index; // [0] = robot, ip+1 = landmark ip
  g.frontier_xy.clear();
  if (!fx.empty()) {  // q15: the reference crashes when no frontier exists; here F = 0
    fro.push_back(nearest(rob.x + g_knife_dx, rob.y + g_knife_dy));   // (analysis knob, 0 in the pinned restatement)
    fro_index.push_back({0});
    for (int ip = 0; ip < L; ++ip) {
      double lx, ly; key_xy(ip, lx, ly);
      const int c = nearest(lx, ly);
      auto it = std::find(fro.begin(), fro.end(), c);  // coordinate equality == cell equality
      if (it != fro.end()) fro_index[it - fro.begin()].push_back(ip + 1);
      else { fro.push_back(c); fro_index.push_back({ip + 1}); }
    }
  }
  const int F = static_cast<int>(fro.size()), N = K + F;
  g.n_nodes = N; g.key_size = K; g.land_size = L; g.fro_size = F; g.nearest_frontier_node = K;
  for (int f = 0; f < F; ++f) { g.frontier_xy.push_back(fx[fro[f]]); g.frontier_xy.push_back(fy[fro[f]]); }

  // ---- adjacency: SLAM2D::adjacency_degree_get (SLAM2D.cpp:198-273) + frontier edges
  // (exploration_env.py:212-224), emitted in data_process order (policy.py:216-227):
  // row-major over the dense matrix, first encounter (i<j), both directions.
  std::vector<std::vector<std::pair<int, double>>> nbr(N);  // neighbours j > i in ascending j
  for (int k = 0; k < T; ++k)
    for (int p = meas_ptr[k]; p < meas_ptr[k + 1]; ++p) {
      const int li = static_cast<int>(std::lower_bound(land_ids.begin(), land_ids.end(), meas[p].id) - land_ids.begin());
      nbr[li].push_back({L + k, meas[p].range});                       // SLAM2D.cpp:256
    }
  for (int k = 0; k + 1 < T; ++k)
    nbr[L + k].push_back({L + k + 1, std::sqrt(odom[3 * k] * odom[3 * k] + odom[3 * k + 1] * odom[3 * k + 1]) + 0.001});  // :239
  for (int f = 0; f < F; ++f)
    for (int idx : fro_index[f]) {
      const double px = g.frontier_xy[2 * f], py = g.frontier_xy[2 * f + 1];
      if (idx == 0) nbr[K - 1].push_back({K + f, points2dist(px, py, rob.x, rob.y)});
      else { double kx, ky; key_xy(idx - 1, kx, ky); nbr[idx - 1].push_back({K + f, points2dist(px, py, kx, ky)}); }
    }
  g.edge_src.clear(); g.edge_dst.clear(); g.edge_w.clear();
  for (int i = 0; i < N; ++i) {
    std::stable_sort(nbr[i].begin(), nbr[i].end(), [](const std::pair<int, double> &a, const std::pair<int, double> &b) { return a.first < b.first; });
    for (auto &e : nbr[i]) {
      if (e.second == 0) continue;  // policy.py:219 (zero weight == no edge)
      g.edge_src.push_back(i); g.edge_dst.push_back(e.first); g.edge_w.push_back(e.second);
      g.edge_src.push_back(e.first); g.edge_dst.push_back(i); g.edge_w.push_back(e.second);
    }
  }
  // ---- features  exploration_env.py:226-274 ----
  std::vector<double> tmap; cov_trace(tmap);
  auto coor2index = [&](double x, double y, int &mi, int &mj) {  // :369-372
    mj = static_cast<int>(py_round((x - cfg.map_min_x) / res - 0.5));
    mi = static_cast<int>(py_round((y - cfg.map_min_y) / res - 0.5));
  };
  g.features.assign(static_cast<size_t>(N) * 5, 0.0);
  for (int i = 0; i < N; ++i) {
    double x, y, f0;
    if (i < L) { key_xy(i, x, y); const M2 &c = land_cov[land_ids[i]]; f0 = c.a[0] + c.a[3]; }
    else if (i < K) { key_xy(i, x, y); const M3 &c = pose_cov[i - L]; f0 = c.a[0] + c.a[4] + c.a[8]; }
    else { x = g.frontier_xy[2 * (i - K)]; y = g.frontier_xy[2 * (i - K) + 1]; int mi, mj; coor2index(x, y, mi, mj); f0 = tmap[mi * cols + mj]; }
    int mi, mj; coor2index(x, y, mi, mj);
    g.features[5 * i + 0] = f0;
    g.features[5 * i + 1] = points2dist(x, y, rob.x, rob.y);
    g.features[5 * i + 2] = diff_theta(x, y, rob.x, rob.y, rob.th);
    g.features[5 * i + 3] = prob[mi * cols + mj];
    g.features[5 * i + 4] = (i < K - 1) ? -1.0 : (i == K - 1 ? 0.0 : 1.0);
  }
}

// ----------------------------------------------------------- line planner ---
std::vector<Pose> Env::line_plan(double gx, double gy) const {  // Planner2D.cpp:937-1041
  std::vector<Pose> actions;
  const Pose &root = est_pose[T - 1];
  double root_theta = root.th;
  double goal_theta = std::atan2(gy - root.y, gx - root.x);
  if (root_theta < 0) root_theta = kPi * 2 + root_theta;
  if (goal_theta < 0) goal_theta = kPi * 2 + goal_theta;
  const double dr = 180 * kPi / 180;
  double diff = goal_theta - root_theta, sign;
  if (diff > kPi) { diff = 2 * kPi - diff; sign = -1; }
  else if (diff > -kPi && diff < 0) { diff = std::fabs(diff); sign = -1; }
  else if (diff <= -kPi) { diff = 2 * kPi - std::fabs(diff); sign = 1; }
  else { sign = 1; }
  const int quotient = static_cast<int>(diff / dr);
  const double remainder = diff - dr * quotient;
  for (int i = 0; i < quotient; ++i) actions.push_back(Pose{0, 0, sign * dr});
  actions.push_back(Pose{0, 0, sign * remainder});
  const double path = std::sqrt(std::pow(root.x - gx, 2) + std::pow(root.y - gy, 2));
  const int dq = static_cast<int>(path / cfg.max_edge_length);
  const double drem = path - dq * cfg.max_edge_length;
  for (int i = 0; i < dq; ++i) actions.push_back(Pose{cfg.max_edge_length, 0, 0});
  actions.push_back(Pose{drem, 0, 0});
  return actions;
}

// ------------------------------------------------------- roll-out reward ---
double Env::simulations_reward(const std::vector<Pose> &actions, const double *noise) const {  // Planner2D.cpp:1416-1468
  Env tmp(*this);  // SLAM2D / VirtualMap / Simulator2D copies, RNG state included (:1417-1420)
  // SLAM2D::set_copy_isam (SLAM2D.cpp:490-497): fresh ISAM2 linearised at calculateBestEstimate()
  tmp.lin_pose = tmp.est_pose;
  std::fill(tmp.delta_pose.begin(), tmp.delta_pose.end(), 0.0);
  tmp.lin_l = tmp.est_l;
  std::fill(tmp.delta_l.begin(), tmp.delta_l.end(), 0.0);
  tmp.update_count = 1;
  const double initial_u = tmp.utility(0.0);
  double dist_ = 0;
  for (const Pose &a : actions) {
    dist_ += std::sqrt(a.x * a.x + a.y * a.y + cfg.angle_weight * a.th * a.th);
    const double o[3] = {a.x, a.y, a.th};
    tmp.forced_noise = noise ? noise + static_cast<size_t>(&a - actions.data()) * (3 + 4 * Lt) : nullptr;
    tmp.move(o, nullptr);
    std::vector<Meas> ms;
    tmp.measure(ms, nullptr, 1);   // a single measure() here (no obstacle probe)
    tmp.add_measurements(ms);
    tmp.meas_ptr.push_back(static_cast<int32_t>(tmp.meas.size()));
    tmp.slam_optimize();           // copy_optimize(true)
    tmp.update_virtual_map();
  }
  return initial_u - tmp.utility(dist_);
}

}  // namespace orc
