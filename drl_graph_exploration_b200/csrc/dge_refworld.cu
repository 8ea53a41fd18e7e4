// The reference's random streams on the host, for the drop-in's `test=True` worlds (host code only; nothing here runs on the GPU).
//
// `ExplorationEnv(map_size, env_index, True)` of the reference (scripts/envs/exploration_env.py:389-407) builds its world and draws its
// noise from libstdc++ generators seeded with env_index (scripts/envs/pyss2d.py:89-119, src/em_exploration/RNG.h:47-126):
//   * Simulator2D keeps three std::mt19937 with the SAME seed (Simulator2D.cpp:436-443): one draws the landmarks -- uniform in the
//     env bounds, rejected within 2 m of the start pose (:452-463) -- one the control noise (three normals per move, x y theta,
//     :161-173), one the sensor noise (bearing then range per landmark within max_range of the TRUE pose, drawn before the
//     field-of-view gate, :113-117, :505-527) -- and pyss2d.simulate calls measure() twice per step (obstacle probe, then the real
//     one, pyss2d.py:182-203), so the sensor stream advances twice;
//   * the landmarks are visited in the iteration order of an std::unordered_map<unsigned, ...> (Simulation2D.h:269,
//     Simulator2D.cpp:335-340), which is a property of the libstdc++ that built the reference: the released result files agree
//     with the hashtable of GCC 5-7 (insertion at the bucket's begin, rehash by re-linking in list order, prime policy with the
//     12-entry fast table), not with today's.  That order is re-derived below by running that hashtable's algorithm on the keys 0..n-1.
// The engine takes worlds and noise as explicit inputs (dge_reset / dge_step: start, landmarks, scan order, noise rows of 3 + 4 Lt
// doubles); this file produces exactly those rows, tracking the true pose with the same composition the simulator applies.
// std::mt19937, std::normal_distribution (Marsaglia polar, one spare) and std::uniform_real_distribution ARE the reference's
// generators (RNG.h wraps them), so they are used as they are.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <random>
#include <vector>

#include "../../include/dge.h"

namespace {

struct Stream {
  std::mt19937 gen;
  std::uniform_real_distribution<> uni{0.0, 1.0};
  std::normal_distribution<> nrm{0.0, 1.0};
  explicit Stream(uint32_t seed = 0) : gen(seed) {}
  double uniform(double lo, double hi) { return (hi - lo) * uni(gen) + lo; }   // RNG.h:68-71
  double normal(double mean, double sigma) { return nrm(gen) * sigma + mean; }  // RNG.h:87-96
};

// iteration order of std::unordered_map<unsigned, T> after inserting the keys 0 .. n-1 one by one, libstdc++ of GCC 5-7
std::vector<int32_t> old_libstdcxx_hash_order(int n) {
  static const unsigned long primes[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 103, 109, 113, 127,
                                         137, 139, 149, 157, 167, 179, 193, 199, 211, 227, 241, 257, 277, 293, 313, 337, 359, 383, 409, 439, 467, 503, 541,
                                         577, 619, 661, 709, 761, 823, 887, 953, 1031, 1109, 1193, 1289, 1381, 1493, 1613, 1741, 1879, 2029, 2179, 2357};
  static const unsigned char fast[12] = {2, 2, 2, 3, 5, 5, 7, 7, 11, 11, 11, 11};
  const int NIL = -2, BEFORE = -1, EMPTY = -3;     // list end / the before-begin node / bucket without nodes
  std::vector<int> next(n, NIL);                   // node k = key k; singly linked list of all nodes
  int first = NIL;
  std::vector<int> bucket(1, EMPTY);               // bucket -> the node in FRONT of the bucket's first node
  size_t n_bkt = 1, resize_at = 0;
  auto link = [&](int node) -> int & { return node == BEFORE ? first : next[node]; };
  auto grow_to = [&](size_t want) {                // _Prime_rehash_policy::_M_next_bkt
    size_t nb;
    if (want <= 11) nb = fast[want];
    else nb = *std::lower_bound(primes, primes + sizeof(primes) / sizeof(primes[0]), (unsigned long)want);
    resize_at = (size_t)std::ceil((long double)nb * 1.0L);
    return nb;
  };
  for (int key = 0; key < n; ++key) {
    const size_t count = (size_t)key;
    if (count + 1 >= resize_at) {                  // _M_need_rehash(n_bkt, n_elt, 1)
      const long double need = (long double)(count + 1);
      if (need >= (long double)n_bkt) {
        const size_t nb = grow_to(std::max<size_t>((size_t)std::floor(need) + 1, n_bkt * 2));
        std::vector<int> fresh(nb, EMPTY);         // _M_rehash_aux, unique keys: walk the list, re-link every node into its new bucket
        int p = first;
        first = NIL;
        size_t begin_bkt = 0;
        while (p != NIL) {
          const int after = next[p];
          const size_t b = (size_t)p % nb;
          if (fresh[b] == EMPTY) {
            next[p] = first; first = p; fresh[b] = BEFORE;
            if (next[p] != NIL) fresh[begin_bkt] = p;
            begin_bkt = b;
          } else {
            next[p] = link(fresh[b]); link(fresh[b]) = p;
          }
          p = after;
        }
        bucket.swap(fresh);
        n_bkt = nb;
      } else {
        resize_at = (size_t)std::floor((long double)n_bkt * 1.0L);
      }
    }
    const size_t b = (size_t)key % n_bkt;          // _M_insert_bucket_begin
    if (bucket[b] != EMPTY) {
      next[key] = link(bucket[b]); link(bucket[b]) = key;
    } else {
      next[key] = first; first = key;
      if (next[key] != NIL) bucket[(size_t)next[key] % n_bkt] = key;
      bucket[b] = BEFORE;
    }
  }
  std::vector<int32_t> order;
  for (int p = first; p != NIL; p = next[p]) order.push_back(p);
  return order;
}

inline double wrap_pi(double a) {                   // Rot2::theta(): atan2 of the kept (cos, sin)
  return std::atan2(std::sin(a), std::cos(a));
}

}  // namespace

struct dge_refworld {
  dge_config cfg;
  int Lt;
  Stream sim, sensor, control;
  double x, y, th;                                  // the simulator's true pose
  std::vector<double> lm;                           // [Lt,2] by id
  std::vector<int32_t> scan;                        // visiting order of the landmarks
  std::vector<double> init_noise;                   // noise row of the first measure() (pyss2d.py:134)

  void measure(double *row, int call) {             // Simulator2D::measure: the draws only (the engine applies them)
    for (int s = 0; s < Lt; ++s) {
      const int id = scan[s];
      const double dx = x - lm[2 * id], dy = y - lm[2 * id + 1];
      if (!(std::sqrt(dx * dx + dy * dy) < cfg.max_range)) continue;           // Distance.cpp:86-88
      row[3 + call * 2 * Lt + 2 * s] = sensor.normal(0.0, cfg.bearing_noise);
      row[3 + call * 2 * Lt + 2 * s + 1] = sensor.normal(0.0, cfg.range_noise);
    }
  }
};

extern "C" dge_refworld *dge_refworld_create(const dge_config *cfg, uint32_t seed, const double *start) {
  if (!cfg || !start || cfg->num_landmarks < 1) return nullptr;
  dge_refworld *w = new dge_refworld{*cfg, cfg->num_landmarks, Stream(seed), Stream(seed), Stream(seed), start[0], start[1], wrap_pi(start[2]), {}, {}, {}};
  for (int i = 0; i < w->Lt;) {                                                 // Simulator2D::addLandmarks (random_landmarks)
    const double lx = w->sim.uniform(cfg->env_min_x, cfg->env_max_x);
    const double ly = w->sim.uniform(cfg->env_min_y, cfg->env_max_y);
    const double dx = lx - start[0], dy = ly - start[1];
    if (std::sqrt(dx * dx + dy * dy) < 2.0) continue;
    w->lm.push_back(lx); w->lm.push_back(ly);
    ++i;
  }
  w->scan = old_libstdcxx_hash_order(w->Lt);
  w->init_noise.assign(3 + 4 * (size_t)w->Lt, 0.0);
  w->measure(w->init_noise.data(), 1);
  return w;
}

extern "C" void dge_refworld_destroy(dge_refworld *w) { delete w; }

extern "C" int dge_refworld_world(const dge_refworld *w, double *landmarks, int32_t *scan, double *init_noise) {
  if (!w || !landmarks || !scan || !init_noise) return DGE_EINVAL;
  std::copy(w->lm.begin(), w->lm.end(), landmarks);
  std::copy(w->scan.begin(), w->scan.end(), scan);
  std::copy(w->init_noise.begin(), w->init_noise.end(), init_noise);
  return DGE_OK;
}

// noise row of one pyss2d.simulate(odom): control noise, then the sensor draws of the obstacle probe and of the real measure()
extern "C" int dge_refworld_step(dge_refworld *w, const double *odom, double *noise) {
  if (!w || !odom || !noise) return DGE_EINVAL;
  std::fill(noise, noise + 3 + 4 * (size_t)w->Lt, 0.0);
  const double nx = w->control.normal(0.0, w->cfg.trans_noise), ny = w->control.normal(0.0, w->cfg.trans_noise), nt = w->control.normal(0.0, w->cfg.rot_noise);
  noise[0] = nx; noise[1] = ny; noise[2] = nt;
  auto compose = [](double &px, double &py, double &pt, double ox, double oy, double ot) {
    const double c = std::cos(pt), s = std::sin(pt);
    const double nx_ = px + c * ox - s * oy, ny_ = py + s * ox + c * oy;
    px = nx_; py = ny_; pt = wrap_pi(pt + ot);
  };
  compose(w->x, w->y, w->th, odom[0], odom[1], odom[2]);                        // Simulator2D::move: (true o odom) o noise
  compose(w->x, w->y, w->th, nx, ny, nt);
  w->measure(noise, 0);
  w->measure(noise, 1);
  return DGE_OK;
}
