// vsr_abi.cu -- the C ABI of include/vsr_b200.h: validation + host-side compilation of a
// batch descriptor (what robot.reset()/assemble()/placement do in the reference,
// Locomotion.java:172-191), upload, kernel launch, read-back.
//
// No CPU fallback: upload/run/results fail with VSR_ERR_NO_DEVICE / VSR_ERR_CUDA when
// there is no usable GPU.  vsr_batch_create itself is host-only (so the argument
// validation -- the reference's IllegalArgumentException cases -- is testable anywhere).
#include "../../include/vsr_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <algorithm>
#include <vector>

#include "vsr_device.cuh"

using namespace vsr;

static thread_local std::string g_err;
static int fail(int code, const std::string& m) {
  g_err = m;
  return code;
}
#define CUDA_TRY(x)                                                                              \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess)                                                                       \
      return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? VSR_ERR_NO_DEVICE \
                                                                               : VSR_ERR_CUDA,   \
                  std::string(#x) + ": " + cudaGetErrorString(e_));                              \
  } while (0)

extern "C" const char* vsr_last_error(void) { return g_err.c_str(); }
extern "C" int vsr_abi_version(void) { return VSR_ABI_VERSION; }
extern "C" int vsr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" void vsr_default_settings(vsr_settings* s) {
  memset(s, 0, sizeof *s);
  s->step_frequency = 1.0 / 60.0;
  s->velocity_iterations = 10;
  s->position_iterations = 10;
  s->linear_tolerance = 0.005;
  s->angular_tolerance = 2.0 * M_PI / 180.0;
  s->max_linear_correction = 0.2;
  s->baumgarte = 0.2;
  s->restitution_velocity = 1.0;
  s->max_translation = 2.0;
  s->max_rotation = 0.5 * M_PI;
  s->gravity_x = 0.0;
  s->gravity_y = -9.8;
  s->ground_friction = 0.2;
  s->ground_restitution = 0.0;
  s->spring_mass_mode = VSR_SPRING_MASS_REDUCED;
  s->self_collision = 1;
  s->continuous_detection = 1;  // ContinuousDetectionMode.ALL, the default of dyn4j's `new Settings()` (Locomotion.java:50)
  s->reserved = 0;
}

extern "C" void vsr_default_material(vsr_material* m) {
  memset(m, 0, sizeof *m);
  m->side_length = 3.0;
  m->mass_side_length_ratio = 0.35;
  m->spring_f = 8.0;
  m->spring_d = 0.3;
  m->mass_linear_damping = 0.1;
  m->mass_angular_damping = 0.1;
  m->friction = 10.0;
  m->restitution = 0.1;
  m->mass = 1.0;
  m->area_ratio_passive_min = -std::numeric_limits<double>::infinity();
  m->area_ratio_passive_max = std::numeric_limits<double>::infinity();
  m->area_ratio_active_min = 0.8;
  m->area_ratio_active_max = 1.2;
  m->spring_scaffoldings = VSR_SCAFFOLD_ALL;
}

// Locomotion.java:195-203 + AbstractTask.java:48: t = 0; while (t < finalT) t = t + dT;
static std::vector<double> tick_times(double final_t, double dt, double t_start = 0.0) {
  std::vector<double> t;
  double tt = t_start;
  while (tt < final_t) {
    tt = tt + dt;
    t.push_back(tt);
  }
  return t;
}
extern "C" int vsr_tick_count(double final_t, double dt) {
  if (!(dt > 0)) return 0;
  return (int)tick_times(final_t, dt).size();
}
extern "C" int vsr_tick_count_from(double t_start, double final_t, double dt) {
  if (!(dt > 0)) return 0;
  return (int)tick_times(final_t, dt, t_start).size();
}

namespace {

struct Xfer {
  void* h = nullptr;  // pinned staging copy, device precision
  void* d = nullptr;
  size_t bytes = 0;
};

// The constants Voxel.assemble derives from the 13 constructor fields (Voxel.java:241-301), for one material.  A batch
// has a default material (vsr_batch_desc.material) and a shape class may bring its own (vsr_shape.material): all voxels
// of a ROBOT share one (the weld and contact rows between voxels are solved in closed form for equal masses).
struct MatDev {
  vsr_material m;
  double L = 0, ms = 0, inv_m = 0, inv_i = 0, half = 0, red_mass = 0;
  double rng_min[3], rng_rest[3], rng_max[3];
  bool lim_lo_on = false, lim_hi_on = false;  // passive-range limits of the springs (Voxel.java:262-285,331-338)
  double lim_lo[3] = {0, 0, 0}, lim_hi[3] = {0, 0, 0};
  unsigned spring_mask = 0;
};

static int validate_material(const vsr_material& m) {
  double lim = (2 * m.mass_side_length_ratio) * (2 * m.mass_side_length_ratio);
  if (m.area_ratio_passive_min > -INFINITY && m.area_ratio_passive_min < lim)  // Voxel.java:132-140
    return fail(VSR_ERR_INVALID_ARGUMENT, "Min of areaRatioPassiveRange cannot be lower than the value imposed by massSideLengthRatio");
  if (m.area_ratio_passive_max < m.area_ratio_passive_min || m.area_ratio_active_max < m.area_ratio_active_min)
    return fail(VSR_ERR_INVALID_ARGUMENT, "Max has to be lower or equal than min");  // DoubleRange.java:29-35
  if (!(m.side_length > 0) || !(m.mass > 0) || !(m.mass_side_length_ratio > 0) || !(m.spring_f > 0))
    return fail(VSR_ERR_INVALID_ARGUMENT, "non-positive voxel material parameter");
  return VSR_OK;
}

// Capacity check instead of a run-time flag: a mass (a square of side mass_side_length, any rotation: an
// interval of at most its diagonal on the x axis) can overlap at most `most` ground polygons at once
// (Ground.java:61-79: one polygon per segment).  The device keeps MAXC constraints per mass, which covers
// every terrain Locomotion.createTerrain can generate (segments >= 1 wide; steppy risers 0.5 wide between
// treads >= 1 wide: at most 3); a profile that could need more is refused here, never truncated at run time.
static int check_terrain_capacity(const vsr_batch_desc* d, const vsr_material& m) {
  const double diag = m.side_length * m.mass_side_length_ratio * std::sqrt(2.0) + 1e-9;
  std::vector<double> lo, hi;  // non-degenerate segments
  for (int i = 0; i + 1 < d->terrain_n; i++)
    if (d->terrain_xs[i + 1] > d->terrain_xs[i]) { lo.push_back(d->terrain_xs[i]); hi.push_back(d->terrain_xs[i + 1]); }
  int most = 0;
  for (size_t i = 0; i < lo.size(); i++) {
    size_t k = i;
    while (k + 1 < lo.size() && lo[k + 1] <= hi[i] + diag) k++;
    most = std::max(most, (int)(k - i + 1));
  }
  if (most > vsr::MAXC)
    return fail(VSR_ERR_UNSUPPORTED, "ground profile too fine: a mass could touch " + std::to_string(most) +
                                         " ground polygons at once, the device keeps " + std::to_string(vsr::MAXC));
  return VSR_OK;
}

static int derive_material(const vsr_material& m, MatDev& o) {
  // ---- material: Voxel.assemble constants, Voxel.java:241-301
  const double L = m.side_length, ms = L * m.mass_side_length_ratio;
  o.m = m;
  o.L = L;
  o.ms = ms;
  const double density = (m.mass / 4) / (ms * ms);
  const double bmass = density * ms * ms, binertia = bmass * (ms * ms + ms * ms) / 12.0;
  o.inv_m = 1.0 / bmass;
  o.inv_i = 1.0 / binertia;
  o.half = ms / 2.0;
  o.red_mass = bmass * bmass / (bmass + bmass);
  {
    double amin = sqrt(L * L * m.area_ratio_active_min), amax = sqrt(L * L * m.area_ratio_active_max);
    double sp_min = amin - 2.0 * ms, sp_rest = L - 2.0 * ms, sp_max = amax - 2.0 * ms;
    o.rng_min[0] = sp_min; o.rng_rest[0] = sp_rest; o.rng_max[0] = sp_max;
    o.rng_min[1] = sqrt(ms * ms + sp_min * sp_min);
    o.rng_rest[1] = sqrt(ms * ms + sp_rest * sp_rest);
    o.rng_max[1] = sqrt(ms * ms + sp_max * sp_max);
    o.rng_min[2] = (amin - ms) * sqrt(2.0);
    o.rng_rest[2] = (L - ms) * sqrt(2.0);
    o.rng_max[2] = (amax - ms) * sqrt(2.0);
    for (int c = 0; c < 3; c++)  // SpringRange, Voxel.java:189-193
      if (o.rng_min[c] > o.rng_rest[c] || o.rng_max[c] < o.rng_rest[c] || o.rng_min[c] < 0) {
        return fail(VSR_ERR_INVALID_ARGUMENT, "Wrong spring range");
      }
    // the passive SpringRanges: finite bounds become the DistanceJoints' lower / upper limits
    o.lim_lo_on = std::isfinite(m.area_ratio_passive_min);
    o.lim_hi_on = std::isfinite(m.area_ratio_passive_max);
    {
      const double pmin = o.lim_lo_on ? sqrt(L * L * m.area_ratio_passive_min) : 0.0;
      const double pmax = o.lim_hi_on ? sqrt(L * L * m.area_ratio_passive_max) : 0.0;
      o.lim_lo[0] = pmin - 2.0 * ms; o.lim_hi[0] = pmax - 2.0 * ms;
      o.lim_lo[1] = sqrt(ms * ms + o.lim_lo[0] * o.lim_lo[0]); o.lim_hi[1] = sqrt(ms * ms + o.lim_hi[0] * o.lim_hi[0]);
      o.lim_lo[2] = (pmin - ms) * sqrt(2.0); o.lim_hi[2] = (pmax - ms) * sqrt(2.0);
      for (int c = 0; c < 3; c++)  // SpringRange, Voxel.java:189-193 (a NaN bound of an infinite range passes there too)
        if ((o.lim_lo_on && (o.lim_lo[c] > o.rng_rest[c] || o.lim_lo[c] < 0)) || (o.lim_hi_on && o.lim_hi[c] < o.rng_rest[c])) {
          return fail(VSR_ERR_INVALID_ARGUMENT, "Wrong spring range");
        }
    }
    unsigned mask = 0;
    if (m.spring_scaffoldings & VSR_SCAFFOLD_SIDE_INTERNAL) mask |= 0x000Fu;
    if (m.spring_scaffoldings & VSR_SCAFFOLD_SIDE_EXTERNAL) mask |= 0x00F0u;
    if (m.spring_scaffoldings & VSR_SCAFFOLD_SIDE_CROSS) mask |= 0xFF00u;
    if (m.spring_scaffoldings & VSR_SCAFFOLD_CENTRAL_CROSS) mask |= 0x30000u;
    o.spring_mask = mask;
  }

  return VSR_OK;
}

struct Bucket {
  MatDev mat;  // the material of this shape class
  ShapeDev shape;
  std::vector<int> robot_ids;  // batch indices, ascending
  std::vector<double> params;  // device layout, double on the host
  long long traj_robots = 0;
  // device
  ShapeDev* d_shape = nullptr;
  int* d_ids = nullptr;
  void* d_params = nullptr;
  void* d_ring = nullptr;
  void* d_traj = nullptr;
  // per-robot initial placement (vsr_batch_desc.robot_placement): offsets from the shape's default one
  std::vector<double> place_off;  // [robots][2]
  void* d_place = nullptr;
  std::vector<double> unplaced_x, unplaced_y;  // the shape's masses with bbox.min.x = 0, y as assembled
  double default_dx = 0, default_dy = 0;
  void* d_vtrace = nullptr;
  void* d_rtrace = nullptr;  // [traj_robots][n_ticks][n_readings]
  void* d_ccd = nullptr;     // continuous detection: poses at the start of the step, poses now, velocities: [36][grid threads]
  size_t ccd_bytes = 0;
  void* d_cbuf = nullptr;
  size_t cbuf_bytes = 0;
  void* d_sbuf = nullptr;  // intra-robot contact arena
  size_t sbuf_bytes = 0;
  void* d_lbuf = nullptr;  // per-block contact lists
  size_t lbuf_bytes = 0;
  // launch plan, fixed at upload (nothing in it depends on the stream): vsr_batch_run is one <<<>>> per bucket
  std::vector<unsigned char> pblob;  // Params<real> by value
  const void* kern = nullptr;
  unsigned grid = 0, threads = 0;
  size_t smem = 0;
};

}  // namespace

struct vsr_batch {
  vsr_batch_desc desc;  // scalar fields only are valid after create
  int precision = 0;
  size_t n_robots = 0;
  std::vector<Bucket> buckets;
  std::vector<int> robot_bucket, robot_pos, robot_nvox;
  std::vector<double> ticks;
  std::vector<double> noise;  // Noisy sensors: the first 2 * n_ticks gaussians of java.util.Random(seed)
  bool noisy = false;
  long long noise_seed = 0;
  double* d_noise = nullptr;
  std::vector<int16_t> win_first;  // [MAX_WINDOWS][n_ticks]
  int ring_depth = 1;
  std::vector<double> terr_x, terr_y;  // local frame
  double base_y = 0, origin_x = 0;
  size_t voxel_steps = 0;
  size_t upload_bytes = 0;
  long long plan_warps = 0;  // warps of all buckets together (the launch plan fills the chip once, not once per bucket)
  // device
  int device = -1;
  bool uploaded = false, ran = false;
  cudaStream_t run_stream = nullptr;
  // shape buckets are independent launches: with several of them they run on side streams, forked
  // from and joined back into the caller's stream
  std::vector<cudaStream_t> side_streams;
  std::vector<cudaEvent_t> side_done;
  cudaEvent_t fork_event = nullptr;
  double* d_ticks = nullptr;
  int16_t* d_win = nullptr;
  void* d_terr_x = nullptr;
  std::vector<double> hgrid;  // coarse height map of the terrain (see Params::hgrid)
  void* d_hgrid = nullptr;
  double hgrid_x0 = 0, hgrid_cw = 2.0;
  void* d_terr_y = nullptr;
  double* d_results = nullptr;
  double* d_bbox = nullptr;  // Robot.boundingBox at the last tick, [n_robots][4]
  int last_launches = 0;
  std::vector<Xfer> xfers;  // not shared: one handle per host thread
};

static int sensor_dim(const vsr_sensor& s) {
  if (s.base == VSR_SENSOR_VELOCITY) return ((s.axes & VSR_AXIS_X) ? 1 : 0) + ((s.axes & VSR_AXIS_Y) ? 1 : 0);
  if (s.base == VSR_SENSOR_LIDAR) return s.axes;
  return 1;
}

// Ground.yAt, Ground.java:104-111
static double ground_y_at(const vsr_batch_desc* d, double x) {
  for (int i = 1; i < d->terrain_n; i++)
    if (d->terrain_xs[i - 1] <= x && x <= d->terrain_xs[i])
      return (x - d->terrain_xs[i - 1]) * (d->terrain_ys[i] - d->terrain_ys[i - 1]) /
                 (d->terrain_xs[i] - d->terrain_xs[i - 1]) +
             d->terrain_ys[i - 1];
  return std::numeric_limits<double>::quiet_NaN();
}

static void free_xfers(vsr_batch* b);
static void free_device(vsr_batch* b) {
  if (b->device >= 0) cudaSetDevice(b->device);
  for (auto st : b->side_streams) cudaStreamDestroy(st);
  for (auto ev : b->side_done) cudaEventDestroy(ev);
  if (b->fork_event) cudaEventDestroy(b->fork_event);
  b->side_streams.clear(); b->side_done.clear(); b->fork_event = nullptr;
  free_xfers(b);  // owns d_ticks, d_win, d_terr_*, d_shape, d_ids, d_params and their pinned copies
  for (auto& k : b->buckets) {
    cudaFree(k.d_ring); cudaFree(k.d_traj); cudaFree(k.d_vtrace); cudaFree(k.d_rtrace); cudaFree(k.d_ccd); cudaFree(k.d_cbuf); cudaFree(k.d_sbuf); cudaFree(k.d_lbuf);
    k.d_shape = nullptr; k.d_ids = nullptr; k.d_params = nullptr; k.d_place = nullptr; k.d_ring = nullptr; k.d_traj = nullptr; k.d_vtrace = nullptr; k.d_rtrace = nullptr; k.d_ccd = nullptr; k.ccd_bytes = 0;
    k.d_cbuf = nullptr; k.cbuf_bytes = 0;
    k.d_sbuf = nullptr; k.sbuf_bytes = 0;
    k.d_lbuf = nullptr; k.lbuf_bytes = 0;
  }
  cudaFree(b->d_results);
  cudaFree(b->d_bbox);
  b->d_bbox = nullptr;
  b->d_ticks = nullptr; b->d_win = nullptr; b->d_terr_x = b->d_terr_y = nullptr; b->d_results = nullptr;
  b->uploaded = b->ran = false;
}

extern "C" void vsr_batch_destroy(vsr_batch* b) {
  if (!b) return;
  if (b->uploaded) free_device(b);
  delete b;
}

extern "C" int vsr_batch_create(const vsr_batch_desc* d, vsr_batch** out) {
  if (!d || !out) return fail(VSR_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (d->abi_version != VSR_ABI_VERSION) return fail(VSR_ERR_INVALID_ARGUMENT, "ABI version mismatch");
  if (d->precision != VSR_PRECISION_F32 && d->precision != VSR_PRECISION_F64)
    return fail(VSR_ERR_INVALID_ARGUMENT, "unknown precision");
  if (d->n_robots == 0) return fail(VSR_ERR_INVALID_ARGUMENT, "empty batch");
  if (d->n_shapes <= 0 || !d->shapes) return fail(VSR_ERR_INVALID_ARGUMENT, "no shape classes");
  if (!d->terrain_xs || !d->terrain_ys) return fail(VSR_ERR_INVALID_ARGUMENT, "no ground profile");
  if (d->terrain_n < 2) return fail(VSR_ERR_INVALID_ARGUMENT, "There must be at least 2 points");  // Ground.java:52-54
  for (int i = 1; i < d->terrain_n; i++)
    if (d->terrain_xs[i] < d->terrain_xs[i - 1])
      return fail(VSR_ERR_INVALID_ARGUMENT, "x coordinates must be sorted");  // Ground.java:55-58
  // every material of the batch: the default one and those the shape classes bring
  for (int si = -1; si < d->n_shapes; si++) {
    const vsr_material* mm = si < 0 ? &d->material : d->shapes[si].material;
    if (!mm) continue;
    int rc = validate_material(*mm);
    if (rc) return rc;
    if ((rc = check_terrain_capacity(d, *mm))) return rc;
  }
  if (!d->params || !d->param_offsets) return fail(VSR_ERR_INVALID_ARGUMENT, "no controller parameters");
  if (d->controller_kind >= VSR_CTRL_DISTRIBUTED && (d->distributed_signals < 0 || d->distributed_signals > 4))
    return fail(VSR_ERR_UNSUPPORTED, "DistributedSensing: 0..4 signals per side are supported");
  if (d->controller_kind < VSR_CTRL_PHASE_SIN || d->controller_kind > VSR_CTRL_DISTRIBUTED_ND)
    return fail(VSR_ERR_INVALID_ARGUMENT, "unknown controller kind");
  const vsr_settings& st = d->settings;
  if (!(st.step_frequency > 0) || st.velocity_iterations < 0 || st.position_iterations < 0)
    return fail(VSR_ERR_INVALID_ARGUMENT, "invalid Settings");
  if (st.spring_mass_mode != VSR_SPRING_MASS_REDUCED && st.spring_mass_mode != VSR_SPRING_MASS_EFFECTIVE)
    return fail(VSR_ERR_INVALID_ARGUMENT, "unknown spring mass mode");
  // (the time-of-impact pass packs a ground polygon's index into 17 bits of its candidate list)
  if (st.continuous_detection && d->terrain_n > (1 << 17))
    return fail(VSR_ERR_UNSUPPORTED, "continuous detection: more than 131072 ground points");

  vsr_batch* b = new vsr_batch();
  b->desc = *d;
  b->precision = (int)d->precision;
  b->n_robots = d->n_robots;
  b->ticks = tick_times(d->final_t, st.step_frequency, d->t_start);
  const int nt = (int)b->ticks.size();
  if (nt == 0) {
    delete b;
    return fail(VSR_ERR_INVALID_ARGUMENT, "finalT gives no tick");
  }
  if (nt > 32000) {
    delete b;
    return fail(VSR_ERR_UNSUPPORTED, "more than 32000 ticks");
  }

  // ---- terrain, local frame
  b->origin_x = d->initial_placement;
  double miny = INFINITY;
  for (int i = 0; i < d->terrain_n; i++) {
    b->terr_x.push_back(d->terrain_xs[i] - b->origin_x);
    b->terr_y.push_back(d->terrain_ys[i]);
    miny = std::min(miny, d->terrain_ys[i]);
  }
  b->base_y = miny - 50.0;  // Ground.MIN_Y_THICKNESS, Ground.java:37,61

  // ---- aggregator windows (AggregatorSensor.java:56-66), one table per distinct interval
  std::vector<double> intervals;
  b->win_first.assign((size_t)MAX_WINDOWS * nt, 0);
  auto window_id = [&](double iv) -> int {
    for (size_t i = 0; i < intervals.size(); i++)
      if (intervals[i] == iv) return (int)i;
    if ((int)intervals.size() == MAX_WINDOWS) return -1;
    intervals.push_back(iv);
    int id = (int)intervals.size() - 1, f = 0;
    for (int k = 0; k < nt; k++) {
      while (b->ticks[f] < b->ticks[k] - iv) f++;
      b->win_first[(size_t)id * nt + k] = (int16_t)f;
      b->ring_depth = std::max(b->ring_depth, k - f + 1);
    }
    return id;
  };

  // ---- shape classes
  std::vector<Bucket> buckets((size_t)d->n_shapes);
  for (int si = 0; si < d->n_shapes; si++) {
    const vsr_shape& sh = d->shapes[si];
    ShapeDev& S = buckets[si].shape;
    memset(&S, 0, sizeof S);
    {
      int rc = derive_material(sh.material ? *sh.material : d->material, buckets[si].mat);
      if (rc) {
        delete b;
        return rc;
      }
    }
    const double L = buckets[si].mat.L, ms = buckets[si].mat.ms;
    if (sh.w <= 0 || sh.h <= 0 || !sh.mask) {
      delete b;
      return fail(VSR_ERR_INVALID_ARGUMENT, "invalid shape grid");
    }
    std::vector<int> grid((size_t)sh.w * sh.h, -1), gx, gy;
    for (int i = 0; i < sh.w * sh.h; i++)
      if (sh.mask[i]) {
        grid[i] = (int)gx.size();
        gx.push_back(i % sh.w);
        gy.push_back(i / sh.w);
      }
    const int nv = (int)gx.size();
    if (nv == 0) {
      delete b;
      return fail(VSR_ERR_INVALID_ARGUMENT, "robot without voxels");
    }
    if (nv > MAX_VOX) {
      delete b;
      return fail(VSR_ERR_UNSUPPORTED, "robots with more than 128 voxels are not on the accelerated path yet");
    }
    S.nvox = nv;
    auto at = [&](int x, int y) { return (x < 0 || y < 0 || x >= sh.w || y >= sh.h) ? -1 : grid[(size_t)y * sh.w + x]; };
    for (int v = 0; v < nv; v++) {
      S.nbr[0][v] = (int8_t)at(gx[v] - 1, gy[v]);
      S.nbr[1][v] = (int8_t)at(gx[v] + 1, gy[v]);
      S.nbr[2][v] = (int8_t)at(gx[v], gy[v] - 1);
      S.nbr[3][v] = (int8_t)at(gx[v], gy[v] + 1);
    }
    // Voxel.java:253-256 + Robot.java:99 + placement Locomotion.java:180-187
    static const double sx[4] = {-1, +1, +1, -1}, sy[4] = {+1, +1, -1, -1};
    const double off = L / 2.0 - ms / 2.0, h = ms / 2.0;
    std::vector<double> X((size_t)nv * 4), Y((size_t)nv * 4);
    double minx = INFINITY;
    for (int v = 0; v < nv; v++)
      for (int k = 0; k < 4; k++) {
        X[v * 4 + k] = sx[k] * off + (double)gx[v] * L;
        Y[v * 4 + k] = sy[k] * off + (double)gy[v] * L;
      }
    for (int v = 0; v < nv; v++) {
      minx = std::min(minx, X[v * 4 + 0] - h);
      minx = std::min(minx, X[v * 4 + 3] - h);
    }
    const double dx = d->initial_placement - minx;
    for (auto& xx : X) xx += dx;
    double min_gap = INFINITY;
    for (int v = 0; v < nv; v++) {
      double cx = 0, mny = INFINITY;
      for (int k = 0; k < 4; k++) {
        cx += X[v * 4 + k];
        mny = std::min(mny, Y[v * 4 + k] - h);
      }
      cx /= 4.0;
      min_gap = std::min(min_gap, mny - ground_y_at(d, cx));
    }
    if (!std::isfinite(min_gap)) {
      delete b;
      return fail(VSR_ERR_INVALID_ARGUMENT, "initial placement outside the ground profile");
    }
    const double dy = 1.0 - min_gap;  // INITIAL_PLACEMENT_Y_GAP
    for (int i = 0; i < nv * 4; i++) {
      S.init_x[i] = X[i] - b->origin_x;
      S.init_y[i] = Y[i] + dy;
    }
    {
      Bucket& kb = buckets[si];
      kb.unplaced_x.resize((size_t)nv * 4);
      kb.unplaced_y = Y;
      for (int i = 0; i < nv * 4; i++) kb.unplaced_x[i] = X[i] - d->initial_placement;  // bbox.min.x at 0
      kb.default_dx = d->initial_placement;
      kb.default_dy = dy;
    }
    // sensors
    int ro = 0, maxch = 0;
    for (int v = 0; v < nv; v++) {
      int s0 = sh.sensor_offsets ? sh.sensor_offsets[v] : 0, s1 = sh.sensor_offsets ? sh.sensor_offsets[v + 1] : 0;
      if (s1 - s0 > MAX_SENS) {
        delete b;
        return fail(VSR_ERR_UNSUPPORTED, "more than 16 sensors per voxel");
      }
      S.n_sens[v] = (int8_t)(s1 - s0);
      S.read_off[v] = (int16_t)ro;
      int ch = 0, nr = 0;
      for (int k = s0; k < s1; k++) {
        const vsr_sensor& q = sh.sensors[k];
        SensorDev& D = S.sens[v][k - s0];
        if (q.base < VSR_SENSOR_TOUCH || q.base > VSR_SENSOR_LIDAR ||
            (q.base == VSR_SENSOR_LIDAR && (q.axes < 1 || q.axes > VSR_LIDAR_MAX_RAYS || q.aggregator != VSR_AGG_NONE)) || q.soft_norm < 0 || q.soft_norm > VSR_NORM_LINEAR ||
            q.aggregator < VSR_AGG_NONE ||
            q.aggregator > VSR_AGG_DYNAMIC_NORM) {
          delete b;
          return fail(VSR_ERR_INVALID_ARGUMENT, "unknown sensor");
        }
        int dim = sensor_dim(q);
        D.base = (int8_t)q.base; D.axes = (int8_t)q.axes; D.rotated = (int8_t)(q.rotated != 0);
        D.agg = (int8_t)q.aggregator; D.soft = (int8_t)q.soft_norm; D.dim = (int8_t)dim;
        D.max_norm = q.max_velocity_norm; D.interval = (float)q.interval;
        D.noise = q.noise_sigma > 0 ? q.noise_sigma : 0.0;
        for (int z = 0; z < VSR_LIDAR_MAX_RAYS; z++) D.dirs[z] = q.base == VSR_SENSOR_LIDAR && z < q.axes ? q.lidar_dirs[z] : 0.0;
        if (q.noise_sigma > 0) {
          if (b->noisy && b->noise_seed != (long long)q.noise_seed) {
            delete b;
            return fail(VSR_ERR_UNSUPPORTED, "Noisy sensors: one seed per batch");
          }
          b->noisy = true;
          b->noise_seed = (long long)q.noise_seed;
        }
        D.win = 0; D.ch = 0;
        if (q.aggregator != VSR_AGG_NONE) {
          if (!(q.interval > 0)) {
            delete b;
            return fail(VSR_ERR_INVALID_ARGUMENT, "aggregator interval must be positive");
          }
          int wid = window_id(q.interval);
          if (wid < 0) {
            delete b;
            return fail(VSR_ERR_UNSUPPORTED, "more than 2 distinct aggregator intervals");
          }
          D.win = (int8_t)wid;
          D.ch = (int8_t)ch;
          ch += dim;
        }
        nr += dim;
      }
      if (nr > MAX_READ) {
        delete b;
        return fail(VSR_ERR_UNSUPPORTED, "more than 24 readings per voxel");
      }
      ro += nr;
      maxch = std::max(maxch, ch);
    }
    S.n_readings = ro;
    S.n_channels = maxch;
    // BreakableVoxel parameters (vsr_breakable), per voxel
    S.has_breakable = 0;
    if (sh.breakable)
      for (int v = 0; v < nv; v++) {
        const vsr_breakable& q = sh.breakable[v];
        if (!q.enabled) continue;
        if (q.malfunctions[VSR_COMPONENT_STRUCTURE] & ~((1u << VSR_MALFUNCTION_NONE) | (1u << VSR_MALFUNCTION_FROZEN))) {
          delete b;
          return fail(VSR_ERR_INVALID_ARGUMENT, "Unsupported structure malfunction type.");  // BreakableVoxel.java:258
        }
        if ((q.malfunctions[0] | q.malfunctions[1] | q.malfunctions[2]) & ~15u) {
          delete b;
          return fail(VSR_ERR_INVALID_ARGUMENT, "unknown malfunction type");
        }
        S.has_breakable = 1;
        S.brk_enabled[v] = 1;
        for (int z = 0; z < 3; z++) {
          S.brk_mal[v][z] = (uint8_t)q.malfunctions[z];
          S.brk_thr[v][z] = q.thresholds[z];
        }
        S.brk_restore[v] = q.restore_time;
        S.brk_seed[v] = (long long)q.random_seed;
      }
    // PhaseSin / TimeFunctions never read a sensor: unless the per-tick dumps (vsr_batch_sensor_readings) were asked
    // for, no reading is observable through the boundary and the sensor programs of such a batch are validated
    // above and then dropped
    if ((d->controller_kind == VSR_CTRL_PHASE_SIN || d->controller_kind == VSR_CTRL_TABULATED) && d->trajectory_robots == 0) {
      for (int v = 0; v < nv; v++) S.n_sens[v] = 0;
      S.n_channels = 0;
    }
    // controller layout
    if (d->controller_kind == VSR_CTRL_PHASE_SIN) S.n_params = 2 + nv;
    else if (d->controller_kind == VSR_CTRL_TABULATED) S.n_params = nt * nv;
    else if (d->controller_kind >= VSR_CTRL_DISTRIBUTED) {
      // one MultiLayerPerceptron per voxel: nOfInputs = readings + 4 * signals, nOfOutputs = 1 + 4 * signals
      // (DistributedSensing.java:141-147); hidden sizes shared; weights stay in flat order per voxel
      if (sh.n_hidden < 0 || sh.n_hidden + 2 > MAX_LAYERS) {
        delete b;
        return fail(VSR_ERR_UNSUPPORTED, "too many MLP layers");
      }
      const int sg = d->distributed_signals;
      S.n_layers = sh.n_hidden + 2;
      for (int i = 0; i < sh.n_hidden; i++) {
        S.neurons[1 + i] = sh.hidden[i];
        if (sh.hidden[i] <= 0 || sh.hidden[i] > VSR_DIST_WIDTH) {
          delete b;
          return fail(VSR_ERR_UNSUPPORTED, "DistributedSensing: hidden layers of 1..40 neurons are supported");
        }
      }
      S.neurons[0] = 0;  // per voxel
      S.neurons[S.n_layers - 1] = d->controller_kind == VSR_CTRL_DISTRIBUTED_ND ? 1 + sg : 1 + 4 * sg;
      int off = 0;
      for (int v = 0; v < nv; v++) {
        const int nin = (v + 1 < nv ? S.read_off[v + 1] : S.n_readings) - S.read_off[v] + 4 * sg;
        S.dist_off[v] = off;
        if (nin > VSR_DIST_WIDTH || S.neurons[S.n_layers - 1] > VSR_DIST_WIDTH) {
          delete b;
          return fail(VSR_ERR_UNSUPPORTED, "DistributedSensing: at most 40 inputs / outputs per voxel");
        }
        int prev = nin;
        for (int i = 1; i < S.n_layers; i++) {
          off += S.neurons[i] * (prev + 1);
          prev = S.neurons[i];
        }
      }
      S.n_params = off;
    } else {
      if (sh.n_hidden < 0 || sh.n_hidden + 2 > MAX_LAYERS) {
        delete b;
        return fail(VSR_ERR_UNSUPPORTED, "too many MLP layers");
      }
      S.n_layers = sh.n_hidden + 2;
      S.neurons[0] = ro;  // CentralizedSensing.nOfInputs, :68-75
      for (int i = 0; i < sh.n_hidden; i++) S.neurons[1 + i] = sh.hidden[i];
      S.neurons[S.n_layers - 1] = nv;  // nOfOutputs, :77-83
      int off = 0;
      for (int i = 1; i < S.n_layers; i++) {
        if (S.neurons[i] <= 0) {
          delete b;
          return fail(VSR_ERR_INVALID_ARGUMENT, "empty MLP layer");
        }
        S.layer_off[i] = off;
        off += S.neurons[i] * (S.neurons[i - 1] + 1);
      }
      S.n_params = off;  // MultiLayerPerceptron.countWeights, :106-112
    }
  }
  for (auto& k : buckets) k.shape.ring_depth = b->ring_depth;

  // ---- robots -> buckets, parameters -> device layout
  b->robot_bucket.resize(b->n_robots);
  b->robot_pos.resize(b->n_robots);
  b->robot_nvox.resize(b->n_robots);
  for (size_t r = 0; r < b->n_robots; r++) {
    int si = d->robot_shape ? d->robot_shape[r] : 0;
    if (si < 0 || si >= d->n_shapes) {
      delete b;
      return fail(VSR_ERR_INVALID_ARGUMENT, "robot_shape out of range");
    }
    Bucket& k = buckets[si];
    const ShapeDev& S = k.shape;
    size_t np = d->param_offsets[r + 1] - d->param_offsets[r];
    if (np != (size_t)S.n_params) {
      delete b;
      char msg[256];
      if (d->controller_kind >= VSR_CTRL_MLP)  // MultiLayerPerceptron.java:56-62
        snprintf(msg, sizeof msg, "Wrong number of weights: %d expected, %zu found", S.n_params, np);
      else
        snprintf(msg, sizeof msg, "robot %zu: %d controller parameters expected, %zu found", r, S.n_params, np);
      return fail(VSR_ERR_INVALID_ARGUMENT, msg);
    }
    const double* P = d->params + d->param_offsets[r];
    b->robot_bucket[r] = si;
    b->robot_pos[r] = (int)k.robot_ids.size();
    b->robot_nvox[r] = S.nvox;
    k.robot_ids.push_back((int)r);
    if ((size_t)r < d->trajectory_robots) k.traj_robots++;
    if (d->robot_placement) {
      // Locomotion.java:180-187 for this robot's own placement: x to the placement, y one unit above the ground
      const double px = d->robot_placement[r];
      double min_gap = INFINITY;
      const double hh = k.mat.half;
      for (int v = 0; v < S.nvox; v++) {
        double cx = 0, mny = INFINITY;
        for (int q = 0; q < 4; q++) {
          cx += k.unplaced_x[v * 4 + q] + px;
          mny = std::min(mny, k.unplaced_y[v * 4 + q] - hh);
        }
        cx /= 4.0;
        min_gap = std::min(min_gap, mny - ground_y_at(d, cx));
      }
      if (!std::isfinite(px) || !std::isfinite(min_gap)) {
        delete b;
        return fail(VSR_ERR_INVALID_ARGUMENT, "initial placement outside the ground profile");
      }
      k.place_off.push_back(px - k.default_dx);
      k.place_off.push_back((1.0 - min_gap) - k.default_dy);
    }
    size_t base = k.params.size();
    k.params.resize(base + np);
    if (d->controller_kind == VSR_CTRL_MLP) {
      // Java flat order [layer][neuron][bias, w_1..w_np] -> [layer][k][neuron] so that
      // consecutive lanes (neurons) read consecutive addresses
      size_t c = 0;
      for (int i = 1; i < S.n_layers; i++) {
        int nn = S.neurons[i], npv = S.neurons[i - 1] + 1;
        for (int j = 0; j < nn; j++)
          for (int q = 0; q < npv; q++) k.params[base + S.layer_off[i] + (size_t)q * nn + j] = P[c++];
      }
    } else {
      for (size_t q = 0; q < np; q++) k.params[base + q] = P[q];
    }
    b->voxel_steps += (size_t)S.nvox * (size_t)nt;
  }
  for (auto& k : buckets)
    if (!k.robot_ids.empty()) b->buckets.push_back(std::move(k));
  // invalidate the caller's pointers in our copy of the descriptor
  b->desc.shapes = nullptr; b->desc.robot_shape = nullptr; b->desc.params = nullptr;
  b->desc.param_offsets = nullptr; b->desc.terrain_xs = nullptr; b->desc.terrain_ys = nullptr;
  *out = b;
  return VSR_OK;
}

// One host->device transfer: a pinned staging copy (already in device precision) and its
// device destination.  Built at the first upload; later uploads of the same batch only
// re-issue the copies (no allocation, pinned source).

static int add_xfer(std::vector<Xfer>& xs, const void* src, size_t bytes, void** dptr) {
  Xfer x;
  x.bytes = bytes;
  CUDA_TRY(cudaMalloc(&x.d, std::max<size_t>(bytes, 16)));
  CUDA_TRY(cudaHostAlloc(&x.h, std::max<size_t>(bytes, 16), cudaHostAllocDefault));
  if (bytes) memcpy(x.h, src, bytes);
  *dptr = x.d;
  xs.push_back(x);
  return VSR_OK;
}
template <typename real>
static int add_xfer_real(std::vector<Xfer>& xs, const std::vector<double>& src, void** dptr) {
  std::vector<real> tmp(src.begin(), src.end());
  return add_xfer(xs, tmp.data(), tmp.size() * sizeof(real), dptr);
}

static void free_xfers(vsr_batch* b) {
  for (auto& x : b->xfers) {
    cudaFree(x.d);
    cudaFreeHost(x.h);
  }
  b->xfers.clear();
}

template <typename real>
static int prepare_bucket(vsr_batch* b, Bucket& k);
static int upload_impl(vsr_batch* b, int device, void* stream);

extern "C" int vsr_batch_upload(vsr_batch* b, int device, void* stream) {
  if (!b) return fail(VSR_ERR_INVALID_ARGUMENT, "null batch");
  const bool first = !b->uploaded;
  const int rc = upload_impl(b, device, stream);
  // a first upload that failed part-way leaves nothing behind: no leak, and a retry starts from scratch
  if (rc != VSR_OK && first) free_device(b);
  return rc;
}

static int upload_impl(vsr_batch* b, int device, void* stream) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(VSR_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return fail(VSR_ERR_INVALID_ARGUMENT, "device out of range");
  if (b->uploaded && b->device != device) free_device(b);
  CUDA_TRY(cudaSetDevice(device));
  b->device = device;
  cudaStream_t s = (cudaStream_t)stream;
  const bool f64 = b->precision == VSR_PRECISION_F64;
  const size_t esz = f64 ? 8 : 4;
  const int nt = (int)b->ticks.size();
  if (!b->uploaded) {
    std::vector<Xfer>& xs = b->xfers;
    int rc;
    if ((rc = add_xfer(xs, b->ticks.data(), nt * sizeof(double), (void**)&b->d_ticks))) return rc;
    if ((rc = add_xfer(xs, b->win_first.data(), b->win_first.size() * sizeof(int16_t), (void**)&b->d_win))) return rc;
    if (b->noisy) {
      // java.util.Random: 48-bit LCG, nextGaussian by the polar method (StrictMath.log/sqrt; libm agrees to 1 ulp)
      unsigned long long seed = ((unsigned long long)b->noise_seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
      auto next = [&](int bits) {
        seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return (long long)(seed >> (48 - bits));
      };
      auto next_double = [&]() { return (double)((next(26) << 27) + next(27)) * (1.0 / (double)(1LL << 53)); };
      b->noise.resize((size_t)VSR_LIDAR_MAX_RAYS * nt);  // a sensor of dimension dim reads entries tick * dim + d
      bool have = false;
      double spare = 0;
      for (auto& g : b->noise) {
        if (have) {
          g = spare;
          have = false;
          continue;
        }
        double v1, v2, sq;
        do {
          v1 = 2 * next_double() - 1;
          v2 = 2 * next_double() - 1;
          sq = v1 * v1 + v2 * v2;
        } while (sq >= 1 || sq == 0);
        double mul = std::sqrt(-2 * std::log(sq) / sq);
        spare = v2 * mul;
        have = true;
        g = v1 * mul;
      }
      if ((rc = add_xfer(xs, b->noise.data(), b->noise.size() * sizeof(double), (void**)&b->d_noise))) return rc;
    }
    {
      // the highest point of the profile over each cell: the polyline's own maximum over the cell's x range (closed,
      // and widened by 1e-3 on both sides: the device assigns cells in float), NOT the higher end of every segment that
      // touches the cell -- the robots start one unit away from the border slope, whose higher end is 100 up
      const size_t tn = b->terr_x.size();
      b->hgrid_x0 = b->terr_x.front();
      const int nc = std::max(1, (int)std::ceil((b->terr_x.back() - b->hgrid_x0) / b->hgrid_cw));
      b->hgrid.assign((size_t)nc, -INFINITY);
      const double pad = 1e-3;
      for (size_t i = 0; i + 1 < tn; i++) {
        const double x0 = b->terr_x[i], x1 = b->terr_x[i + 1], y0 = b->terr_y[i], y1 = b->terr_y[i + 1];
        const int c0 = std::max(0, std::min(nc - 1, (int)std::floor((x0 - pad - b->hgrid_x0) / b->hgrid_cw)));
        const int c1 = std::max(0, std::min(nc - 1, (int)std::floor((x1 + pad - b->hgrid_x0) / b->hgrid_cw)));
        for (int c = c0; c <= c1; c++) {
          const double lo = b->hgrid_x0 + c * b->hgrid_cw - pad, hi = b->hgrid_x0 + (c + 1) * b->hgrid_cw + pad;
          double h = std::max(y0, y1);
          if (x1 > x0) {
            const double xa = std::min(std::max(lo, x0), x1), xb = std::max(std::min(hi, x1), x0);
            h = std::max(y0 + (y1 - y0) * (xa - x0) / (x1 - x0), y0 + (y1 - y0) * (xb - x0) / (x1 - x0));
          }
          b->hgrid[(size_t)c] = std::max(b->hgrid[(size_t)c], h);
        }
      }
      rc = f64 ? add_xfer_real<double>(xs, b->hgrid, &b->d_hgrid) : add_xfer_real<float>(xs, b->hgrid, &b->d_hgrid);
      if (rc) return rc;
    }
    if (f64) {
      if ((rc = add_xfer_real<double>(xs, b->terr_x, &b->d_terr_x))) return rc;
      if ((rc = add_xfer_real<double>(xs, b->terr_y, &b->d_terr_y))) return rc;
    } else {
      if ((rc = add_xfer_real<float>(xs, b->terr_x, &b->d_terr_x))) return rc;
      if ((rc = add_xfer_real<float>(xs, b->terr_y, &b->d_terr_y))) return rc;
    }
    CUDA_TRY(cudaMalloc((void**)&b->d_results, b->n_robots * sizeof(vsr_result)));
    CUDA_TRY(cudaMalloc((void**)&b->d_bbox, b->n_robots * 4 * sizeof(double)));
    for (auto& k : b->buckets) {
      const size_t nr = k.robot_ids.size();
      if ((rc = add_xfer(xs, &k.shape, sizeof(ShapeDev), (void**)&k.d_shape))) return rc;
      if ((rc = add_xfer(xs, k.robot_ids.data(), nr * sizeof(int), (void**)&k.d_ids))) return rc;
      rc = f64 ? add_xfer_real<double>(xs, k.params, &k.d_params) : add_xfer_real<float>(xs, k.params, &k.d_params);
      if (rc) return rc;
      if (!k.place_off.empty()) {
        rc = f64 ? add_xfer_real<double>(xs, k.place_off, &k.d_place) : add_xfer_real<float>(xs, k.place_off, &k.d_place);
        if (rc) return rc;
      }
      size_t ring = (size_t)k.shape.ring_depth * (size_t)std::max(1, k.shape.n_channels) * nr * (size_t)k.shape.nvox;
      CUDA_TRY(cudaMalloc(&k.d_ring, ring * esz));
      if (k.traj_robots > 0) {
        size_t tb = (size_t)k.traj_robots * nt * (size_t)k.shape.nvox * 4 * 6 * esz;
        CUDA_TRY(cudaMalloc(&k.d_traj, tb));
        CUDA_TRY(cudaMalloc(&k.d_vtrace, (size_t)k.traj_robots * nt * (size_t)k.shape.nvox * 2 * esz));
        CUDA_TRY(cudaMalloc(&k.d_rtrace, std::max<size_t>(1, (size_t)k.traj_robots * nt * (size_t)k.shape.n_readings) * esz));
      }
    }
    // kernel arguments, launch geometry, shared-memory carve-out and the contact arenas of every bucket.  The
    // buckets run concurrently (side streams) and every block takes a whole SM's shared memory, so the block size
    // follows the load of the WHOLE batch: small buckets of a big batch get full blocks, not one warp per SM.
    b->plan_warps = 0;
    for (auto& k : b->buckets) {
      const int nv = k.shape.nvox;
      b->plan_warps += nv > 32 ? (long long)k.robot_ids.size() * ((nv + 31) / 32)
                               : ((long long)k.robot_ids.size() + 32 / nv - 1) / (32 / nv);
    }
    for (auto& k : b->buckets)
      if ((rc = f64 ? prepare_bucket<double>(b, k) : prepare_bucket<float>(b, k))) return rc;
  }
  // the copies proper: every input of the episode, from pinned memory
  b->upload_bytes = 0;
  for (auto& x : b->xfers) {
    if (x.bytes) CUDA_TRY(cudaMemcpyAsync(x.d, x.h, x.bytes, cudaMemcpyHostToDevice, s));
    b->upload_bytes += x.bytes;
  }
  CUDA_TRY(cudaMemsetAsync(b->d_results, 0, b->n_robots * sizeof(vsr_result), s));
  for (auto& k : b->buckets)
    if (k.traj_robots > 0)
    {
      CUDA_TRY(cudaMemsetAsync(k.d_traj, 0, (size_t)k.traj_robots * nt * (size_t)k.shape.nvox * 4 * 6 * esz, s));
      CUDA_TRY(cudaMemsetAsync(k.d_vtrace, 0, (size_t)k.traj_robots * nt * (size_t)k.shape.nvox * 2 * esz, s));
      CUDA_TRY(cudaMemsetAsync(k.d_rtrace, 0, std::max<size_t>(1, (size_t)k.traj_robots * nt * (size_t)k.shape.n_readings) * esz, s));
    }
  CUDA_TRY(cudaStreamSynchronize(s));
  b->uploaded = true;
  b->ran = false;
  return VSR_OK;
}

// Development knobs (block size, pool size, optional barriers): compiled in with -DVSR_TUNING only; the release
// library reads no environment variable.
static const char* tuning_env(const char* name) {
#ifdef VSR_TUNING
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

// Everything of a launch that does not depend on the stream, done once at upload.
template <typename real>
static int prepare_bucket(vsr_batch* b, Bucket& k) {
  const vsr_batch_desc& d = b->desc;
  const vsr_settings& st = d.settings;
  const MatDev& M = k.mat;
  const vsr_material& m = M.m;
  Params<real> P;
  memset(&P, 0, sizeof P);
  P.shape = k.d_shape;
  P.n_robots = (long long)k.robot_ids.size();
  P.robot_ids = k.d_ids;
  P.params = (const real*)k.d_params;
  P.final_bbox = b->d_bbox;
  P.place_off = (const real*)k.d_place;
  P.results = b->d_results;
  P.ring = (real*)k.d_ring;
  P.trajectory = (real*)k.d_traj;
  P.traj_robots = k.traj_robots;
  P.vtrace = (real*)k.d_vtrace;
  P.rtrace = (real*)k.d_rtrace;
  P.n_ticks = (int)b->ticks.size();
  P.tick_t = b->d_ticks;
  P.noise_tab = b->d_noise;
  P.win_first = b->d_win;
  P.terr_x = (const real*)b->d_terr_x;
  P.terr_y = (const real*)b->d_terr_y;
  P.terr_n = (int)b->terr_x.size();
  P.hgrid = (const real*)b->d_hgrid;
  P.hgrid_n = (int)b->hgrid.size();
  P.hgrid_x0 = (real)b->hgrid_x0;
  P.hgrid_inv = (real)(1.0 / b->hgrid_cw);
  P.base_y = (real)b->base_y;
  P.origin_x = b->origin_x;
  P.dt = (real)st.step_frequency;
  P.grav_x = (real)st.gravity_x;
  P.grav_y = (real)st.gravity_y;
  P.lin_damp = (real)m.mass_linear_damping;
  P.ang_damp = (real)m.mass_angular_damping;
  P.inv_m = (real)M.inv_m;
  P.inv_i = (real)M.inv_i;
  P.half = (real)M.half;
  P.side = (real)m.side_length;
  P.lin_tol = (real)st.linear_tolerance;
  P.ang_tol = (real)st.angular_tolerance;
  P.max_corr = (real)st.max_linear_correction;
  P.baumgarte = (real)st.baumgarte;
  P.rest_vel = (real)st.restitution_velocity;
  P.max_trans = (real)st.max_translation;
  P.max_rot = (real)st.max_rotation;
  P.mu = (real)sqrt(m.friction * st.ground_friction);               // CoefficientMixer: sqrt(f1 f2)
  P.rest_e = (real)std::max(m.restitution, st.ground_restitution);  // max(e1, e2)
  P.spring_w = (real)(2.0 * M_PI * m.spring_f);
  P.spring_zeta = (real)m.spring_d;
  P.red_mass = (real)M.red_mass;
  for (int c = 0; c < 3; c++) {
    P.rng_min[c] = (real)M.rng_min[c];
    P.rng_rest[c] = (real)M.rng_rest[c];
    P.rng_max[c] = (real)M.rng_max[c];
    P.lim_lo[c] = (real)M.lim_lo[c];
    P.lim_hi[c] = (real)M.lim_hi[c];
  }
  P.lim_lo_on = M.lim_lo_on ? 1 : 0;
  P.lim_hi_on = M.lim_hi_on ? 1 : 0;
  // the MODE 2 instantiation: rope limits and / or BreakableVoxels (its rigid-row tables serve both)
  const bool lim = M.lim_lo_on || M.lim_hi_on || k.shape.has_breakable;
  // MODE: 1 = the default voxel, 3 = the default voxel + continuous detection, 2 = limits / BreakableVoxels (and
  // continuous detection for everything that is not the default voxel), 0 = the other generic cases
  const bool fast = !lim && (M.spring_mask == 0x3FFFFu) && st.spring_mass_mode == VSR_SPRING_MASS_REDUCED;
  const int mode = fast ? (st.continuous_detection ? 3 : 1) : ((lim || st.continuous_detection) ? 2 : 0);
  const bool lim_layout = mode == 2;  // the MODE 2 tables: 256 / 128 thread columns, (rigid mass, length) per spring
  P.weld_len = (real)(2.0 * M.half);
  P.vel_iters = st.velocity_iterations;
  P.pos_iters = st.position_iterations;
  P.spring_mode = st.spring_mass_mode;
  P.self_coll = st.self_collision ? 1 : 0;
  P.ccd = st.continuous_detection ? 1 : 0;
  P.mu_self = (real)std::sqrt(m.friction * m.friction);          // CoefficientMixer: sqrt(f1 f2)
  P.rest_e_self = (real)std::max(m.restitution, m.restitution);  // max(e1, e2)
  P.spring_mask = M.spring_mask;
  P.mlp_act = d.mlp_activation;
  P.dist_signals = d.controller_kind >= VSR_CTRL_DISTRIBUTED ? d.distributed_signals : -1;
  P.dist_nd = d.controller_kind == VSR_CTRL_DISTRIBUTED_ND ? 1 : 0;
  P.debug_robot = -1;
  P.debug_tick = -1;
  if (const char* dbg = tuning_env("VSR_DEBUG")) {  // "robot:tick" (bucket index)
    long dr = -1, dk = -1, de = -1;
    int got = sscanf(dbg, "%ld:%ld:%ld", &dr, &dk, &de);
    if (got >= 2) { P.debug_robot = dr; P.debug_tick = (int)dk; P.debug_tick_end = (int)(got == 3 ? de : dk); }
  }
  const int nvox = k.shape.nvox;
  const bool blk = nvox > 32;  // one block per robot (33..128 voxels)
  const int G = blk ? 1 : 32 / nvox;
  int maxn = 1;
  for (int i = 0; i < k.shape.n_layers; i++) maxn = std::max(maxn, k.shape.neurons[i]);
  int per_robot = 2 * ((maxn + 3) & ~3);
  int wpb_env = tuning_env("VSR_WPB") ? atoi(tuning_env("VSR_WPB")) : 0;
  const long long groups_all = ((long long)k.robot_ids.size() + G - 1) / G;
  // Warps of a block walk the instruction stream together (see the barriers in the kernel), so
  // the more warps per block the better, up to the register-file limit of 12.  A batch that does
  // not fill the chip is spread as one block per SM with just enough warps each.
  const int WPB_MAX = VSR_MAX_THREADS / 32;
  int wpb = WPB_MAX;
  {
    int sms_ = 148;
    cudaDeviceGetAttribute(&sms_, cudaDevAttrMultiProcessorCount, b->device);
    const long long slots = (long long)sms_ * VSR_MIN_BLOCKS;  // resident blocks on the chip
    const long long load = std::max(groups_all, b->plan_warps);
    if (load <= slots * WPB_MAX) wpb = (int)std::max<long long>(1, (load + slots - 1) / slots);
    wpb = (int)std::max<long long>(1, std::min<long long>(wpb, groups_all));
  }
  const int wpb_fill = wpb;  // warps per block that fill the chip once with this batch
  if (wpb_env > 0) wpb = std::min(WPB_MAX, wpb_env);
  const bool eff_mode = st.spring_mass_mode == VSR_SPRING_MASS_EFFECTIVE;
  if (sizeof(real) == 8) wpb = std::min(wpb, VSR_MAX_THREADS / 64);  // fp64: 32-byte vectors, half the stride
  // Big robots: a robot takes wpr whole warps and a block holds rpb robots that walk the instruction stream
  // together (one 2..4-warp block per robot ran at its own PC: instruction fetch was 30 % of the stall samples);
  // the block's tables have `stride` thread columns (fp32 256, fp64 128).
  const int stride = (blk || lim_layout) ? (sizeof(real) == 4 ? 256 : 128) : (sizeof(real) == 4 ? VSR_MAX_THREADS : VSR_MAX_THREADS / 2);
  if (!blk) wpb = std::min(wpb, stride / 32);
  const int wpr = blk ? (nvox + 31) / 32 : 1;
  int rpb = blk ? std::max(1, std::min(stride / 32, std::max(wpb_fill, wpr)) / wpr) : 1;
  if (blk) {
    if (const char* e = tuning_env("VSR_RPB")) rpb = std::max(1, std::min(rpb, atoi(e)));
    rpb = (int)std::max<long long>(1, std::min<long long>(rpb, (long long)k.robot_ids.size()));
    wpb = wpr * rpb;
  }
  if (blk) P.smem_per_warp = std::max(128, (per_robot + wpr - 1) / wpr);  // a robot's staging = wpr * this
  else P.smem_per_warp = std::max(128, per_robot * G);
  const size_t limit = 227 * 1024 - 16;  // (16: the kernel's one static word, the candidate count of continuous detection)
  // [18][stride] spring vectors (6 reals, + gamma in effective-mass mode) + per-warp staging and
  // contact exchange (+ the neighbour exchange of the big-robot variant)
  auto fixed_smem = [&](int wpb_, int rpb_) {
    return (size_t)18 * stride * (6 + (eff_mode ? 1 : 0) + (lim_layout ? 2 : 0)) * sizeof(real) +
           ((size_t)wpb_ * P.smem_per_warp + (size_t)16 * stride + (size_t)(blk ? rpb_ : wpb_) * VSR_LHD_REALS) * sizeof(real) +
           (blk ? (size_t)2 * 20 * stride * sizeof(real) : 0);
  };
  // (a big robot's pool must at least hold the pooled detection scratch of its warps: 576 words each)
  const size_t min_pool = blk ? (size_t)((576 * wpr + VSR_POOL_FIELDS - 1) / VSR_POOL_FIELDS + 3) / 4 * 4 * VSR_POOL_FIELDS * sizeof(real) : 0;
  while (blk && rpb > 1 && fixed_smem(wpr * rpb, rpb) + rpb * min_pool > limit) rpb--;
  if (blk) wpb = wpr * rpb;
  P.wpr = wpr;
  P.rpb = rpb;
  const int threads = 32 * wpb;
  size_t smem = fixed_smem(wpb, rpb);
  {
    // what is left of the SM's shared memory caches the block's contact constraints
    long long cap = smem < limit ? (long long)((limit - smem) / (VSR_POOL_FIELDS * sizeof(real))) : 0;
    // one pool per scope: the warp (its robots' constraints), or the block of a big robot
    const int scopes = blk ? rpb : wpb;
    cap = std::min<long long>(cap / scopes, blk ? (long long)VSR_ITEMS_PER_THREAD * 32 * wpr : VSR_ITEMS_PER_THREAD * 32) / 4 * 4;
    // a big robot: enough for the pooled detection scratch of its warps (576 words each), no more than 160 constraints
    if (blk) cap = std::min<long long>(cap, tuning_env("VSR_BLK_POOL") ? atoll(tuning_env("VSR_BLK_POOL")) : 160);
    if (const char* e = tuning_env("VSR_POOL")) cap = std::min<long long>(cap, atoll(e)) / 4 * 4;
    P.pool_cap = (int)cap;
    smem += (size_t)cap * scopes * VSR_POOL_FIELDS * sizeof(real);
  }
  P.sync_mode = tuning_env("VSR_SYNC") ? atoi(tuning_env("VSR_SYNC")) : (2 | 16 | 32 | 64);
  void (*kern)(const Params<real>) = nullptr;
#define VSR_PICK(C)                                                                                  \
  kern = blk ? (mode == 1 ? vsr_episode_kernel<real, C, 1, true> : (mode == 3 ? vsr_episode_kernel<real, C, 3, true> : (mode == 2 ? vsr_episode_kernel<real, C, 2, true> : vsr_episode_kernel<real, C, 0, true>))) \
             : (mode == 1 ? vsr_episode_kernel<real, C, 1, false> : (mode == 3 ? vsr_episode_kernel<real, C, 3, false> : (mode == 2 ? vsr_episode_kernel<real, C, 2, false> : vsr_episode_kernel<real, C, 0, false>)))
#ifdef VSR_DEV_SUBSET  // development builds only (scripts/dev_build.sh): the two instantiations of BASELINE config 2
  if constexpr (std::is_same<real, float>::value) {
    if (!blk && d.controller_kind == VSR_CTRL_PHASE_SIN)
      kern = mode == 1 ? vsr_episode_kernel<float, CTRL_PHASE_SIN, 1, false> : (mode == 3 ? vsr_episode_kernel<float, CTRL_PHASE_SIN, 3, false> : nullptr);
  } else {
#ifdef VSR_DEV_F64
    if (!blk && d.controller_kind == VSR_CTRL_PHASE_SIN && mode == 3) kern = vsr_episode_kernel<double, CTRL_PHASE_SIN, 3, false>;
#endif
  }
  if (!kern) return fail(VSR_ERR_CUDA, "this development build holds the config-2 kernels only");
#else
  switch (d.controller_kind) {
    case VSR_CTRL_PHASE_SIN: VSR_PICK(CTRL_PHASE_SIN); break;
    case VSR_CTRL_TABULATED: VSR_PICK(CTRL_TABULATED); break;
    default: VSR_PICK(CTRL_MLP); break;
  }
#endif
#undef VSR_PICK
  P.ctrl = d.controller_kind;
  // the attribute belongs to the kernel, not to the launch: buckets that share an instantiation ask for different
  // sizes, so it is raised to the SM's limit once and for all (the carve-out follows each launch's actual size)
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 16));  // (see `limit`)
  int occ = 0, sms = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device));
  if (occ < 1) return fail(VSR_ERR_CUDA, "episode kernel does not fit on an SM");
  const long long groups = (P.n_robots + G - 1) / G;
  const long long want = blk ? (groups + rpb - 1) / rpb : (groups + wpb - 1) / wpb;
  // persistent grid: one wave of resident blocks (a multiple of the SM count), each warp
  // (or block, for big robots) strides over robot groups
  const long long grid = std::max<long long>(1, std::min<long long>(want, (long long)occ * sms));
  {
    size_t need = (size_t)grid * threads * 8 * sizeof(CBody<real>);
    if (need > k.cbuf_bytes) {
      cudaFree(k.d_cbuf);
      k.d_cbuf = nullptr;
      CUDA_TRY(cudaMalloc(&k.d_cbuf, need));
      k.cbuf_bytes = need;
    }
    P.cbuf = k.d_cbuf;
    size_t need_s = P.self_coll ? (size_t)grid * threads * 2 * sizeof(SList<real>) : 0;
    if (need_s > k.sbuf_bytes) {
      cudaFree(k.d_sbuf);
      k.d_sbuf = nullptr;
      CUDA_TRY(cudaMalloc(&k.d_sbuf, need_s));
      k.sbuf_bytes = need_s;
    }
    P.sbuf = k.d_sbuf;
    size_t need_c = P.ccd ? (size_t)grid * threads * 40 * sizeof(real) : 0;
    if (need_c > k.ccd_bytes) {
      cudaFree(k.d_ccd);
      k.d_ccd = nullptr;
      CUDA_TRY(cudaMalloc(&k.d_ccd, need_c));
      k.ccd_bytes = need_c;
    }
    P.ccd_buf = (real*)k.d_ccd;
    size_t need_l = (size_t)grid * VSR_ITEMS_PER_THREAD * stride * sizeof(int);
    if (need_l > k.lbuf_bytes) {
      cudaFree(k.d_lbuf);
      k.d_lbuf = nullptr;
      CUDA_TRY(cudaMalloc(&k.d_lbuf, need_l));
      k.lbuf_bytes = need_l;
    }
    P.lbuf = (int*)k.d_lbuf;
  }
  k.pblob.assign(reinterpret_cast<const unsigned char*>(&P), reinterpret_cast<const unsigned char*>(&P) + sizeof P);
  k.kern = reinterpret_cast<const void*>(kern);
  k.grid = (unsigned)grid;
  k.threads = (unsigned)threads;
  k.smem = smem;
  return VSR_OK;
}

static int launch_bucket(vsr_batch* b, Bucket& k, cudaStream_t s) {
  void* args[1] = {k.pblob.data()};
  CUDA_TRY(cudaLaunchKernel(k.kern, dim3(k.grid), dim3(k.threads), args, k.smem, s));
  b->last_launches++;
  return VSR_OK;
}

extern "C" int vsr_batch_run(vsr_batch* b, void* stream) {
  if (!b) return fail(VSR_ERR_INVALID_ARGUMENT, "null batch");
  if (!b->uploaded) return fail(VSR_ERR_STATE, "vsr_batch_run before vsr_batch_upload");
  CUDA_TRY(cudaSetDevice(b->device));
  cudaStream_t s = (cudaStream_t)stream;
  b->run_stream = s;
  b->last_launches = 0;
  const size_t nb = b->buckets.size();
  if (nb <= 1) {
    for (auto& k : b->buckets) {
      int rc = launch_bucket(b, k, s);
      if (rc) return rc;
    }
  } else {
    const size_t ns = std::min<size_t>(nb, 16);
    if (b->side_streams.empty()) {
      b->side_streams.resize(ns);
      b->side_done.resize(ns);
      for (size_t i = 0; i < ns; i++) {
        CUDA_TRY(cudaStreamCreateWithFlags(&b->side_streams[i], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&b->side_done[i], cudaEventDisableTiming));
      }
      CUDA_TRY(cudaEventCreateWithFlags(&b->fork_event, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventRecord(b->fork_event, s));
    for (size_t i = 0; i < ns; i++) CUDA_TRY(cudaStreamWaitEvent(b->side_streams[i], b->fork_event, 0));
    // the heaviest buckets first
    std::vector<size_t> order(nb);
    for (size_t i = 0; i < nb; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t x, size_t y) {
      return b->buckets[x].robot_ids.size() * (size_t)b->buckets[x].shape.nvox >
             b->buckets[y].robot_ids.size() * (size_t)b->buckets[y].shape.nvox;
    });
    for (size_t q = 0; q < nb; q++) {
      Bucket& k = b->buckets[order[q]];
      cudaStream_t ss = b->side_streams[q % ns];
      int rc = launch_bucket(b, k, ss);
      if (rc) return rc;
    }
    for (size_t i = 0; i < ns; i++) {
      CUDA_TRY(cudaEventRecord(b->side_done[i], b->side_streams[i]));
      CUDA_TRY(cudaStreamWaitEvent(s, b->side_done[i], 0));
    }
  }
  b->ran = true;
  return VSR_OK;
}

extern "C" int vsr_batch_results(vsr_batch* b, vsr_result* out, size_t n) {
  if (!b || !out) return fail(VSR_ERR_INVALID_ARGUMENT, "null argument");
  if (!b->ran) return fail(VSR_ERR_STATE, "vsr_batch_results before vsr_batch_run");
  if (n != b->n_robots) return fail(VSR_ERR_INVALID_ARGUMENT, "result count differs from n_robots");
  static_assert(sizeof(vsr_result) == 11 * sizeof(double), "vsr_result layout");
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaMemcpyAsync(out, b->d_results, n * sizeof(vsr_result), cudaMemcpyDeviceToHost, b->run_stream));
  CUDA_TRY(cudaStreamSynchronize(b->run_stream));
  return VSR_OK;
}

extern "C" int vsr_batch_results_device(vsr_batch* b, void** dptr, size_t* bytes) {
  if (!b || !dptr) return fail(VSR_ERR_INVALID_ARGUMENT, "null argument");
  if (!b->uploaded) return fail(VSR_ERR_STATE, "not uploaded");
  *dptr = b->d_results;
  if (bytes) *bytes = b->n_robots * sizeof(vsr_result);
  return VSR_OK;
}

extern "C" long long vsr_batch_trajectory(vsr_batch* b, size_t robot, double* out, size_t capacity) {
  if (!b || !out) return fail(VSR_ERR_INVALID_ARGUMENT, "null argument");
  if (!b->ran) return fail(VSR_ERR_STATE, "vsr_batch_trajectory before vsr_batch_run");
  if (robot >= b->n_robots || robot >= b->desc.trajectory_robots)
    return fail(VSR_ERR_INVALID_ARGUMENT, "robot has no recorded trajectory");
  // buckets were compacted after create: find the one holding this robot
  Bucket* kb = nullptr;
  int pos = -1;
  for (auto& k : b->buckets)
    for (size_t i = 0; i < k.robot_ids.size() && (long long)i < k.traj_robots; i++)
      if ((size_t)k.robot_ids[i] == robot) {
        kb = &k;
        pos = (int)i;
      }
  if (!kb) return fail(VSR_ERR_STATE, "trajectory bucket not found");
  const int nt = (int)b->ticks.size();
  const size_t n = (size_t)nt * kb->shape.nvox * 4 * 6;
  if (capacity < n) return fail(VSR_ERR_INVALID_ARGUMENT, "trajectory buffer too small");
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaStreamSynchronize(b->run_stream));
  if (b->precision == VSR_PRECISION_F64) {
    CUDA_TRY(cudaMemcpy(out, (double*)kb->d_traj + (size_t)pos * n, n * sizeof(double), cudaMemcpyDeviceToHost));
  } else {
    std::vector<float> tmp(n);
    CUDA_TRY(cudaMemcpy(tmp.data(), (float*)kb->d_traj + (size_t)pos * n, n * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) out[i] = tmp[i];
  }
  // world frame
  for (size_t i = 0; i < n; i += 6) out[i] += b->origin_x;
  return (long long)n;
}

extern "C" int vsr_batch_final_bbox(vsr_batch* b, double* out, size_t n) {
  if (!b || !out) return fail(VSR_ERR_INVALID_ARGUMENT, "null argument");
  if (!b->ran) return fail(VSR_ERR_STATE, "vsr_batch_final_bbox before vsr_batch_run");
  if (n != b->n_robots) return fail(VSR_ERR_INVALID_ARGUMENT, "bounding box count differs from n_robots");
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaMemcpyAsync(out, b->d_bbox, n * 4 * sizeof(double), cudaMemcpyDeviceToHost, b->run_stream));
  CUDA_TRY(cudaStreamSynchronize(b->run_stream));
  return VSR_OK;
}

extern "C" long long vsr_batch_voxel_trace(vsr_batch* b, size_t robot, double* out, size_t capacity) {
  if (!b || !out) return fail(VSR_ERR_INVALID_ARGUMENT, "null argument");
  if (!b->ran) return fail(VSR_ERR_STATE, "vsr_batch_voxel_trace before vsr_batch_run");
  if (robot >= b->n_robots || robot >= b->desc.trajectory_robots)
    return fail(VSR_ERR_INVALID_ARGUMENT, "robot has no recorded trajectory");
  Bucket* kb = nullptr;
  int pos = -1;
  for (auto& k : b->buckets)
    for (size_t i = 0; i < k.robot_ids.size() && (long long)i < k.traj_robots; i++)
      if ((size_t)k.robot_ids[i] == robot) {
        kb = &k;
        pos = (int)i;
      }
  if (!kb) return fail(VSR_ERR_STATE, "trajectory bucket not found");
  const int nt = (int)b->ticks.size();
  const size_t n = (size_t)nt * kb->shape.nvox * 2;
  if (capacity < n) return fail(VSR_ERR_INVALID_ARGUMENT, "voxel trace buffer too small");
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaStreamSynchronize(b->run_stream));
  if (b->precision == VSR_PRECISION_F64) {
    CUDA_TRY(cudaMemcpy(out, (double*)kb->d_vtrace + (size_t)pos * n, n * sizeof(double), cudaMemcpyDeviceToHost));
  } else {
    std::vector<float> tmp(n);
    CUDA_TRY(cudaMemcpy(tmp.data(), (float*)kb->d_vtrace + (size_t)pos * n, n * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) out[i] = tmp[i];
  }
  return (long long)n;
}

extern "C" long long vsr_batch_sensor_readings(vsr_batch* b, size_t robot, double* out, size_t capacity) {
  if (!b || !out) return fail(VSR_ERR_INVALID_ARGUMENT, "null argument");
  if (!b->ran) return fail(VSR_ERR_STATE, "vsr_batch_sensor_readings before vsr_batch_run");
  if (robot >= b->n_robots || robot >= b->desc.trajectory_robots)
    return fail(VSR_ERR_INVALID_ARGUMENT, "robot has no recorded trajectory");
  Bucket* kb = nullptr;
  int pos = -1;
  for (auto& k : b->buckets)
    for (size_t i = 0; i < k.robot_ids.size() && (long long)i < k.traj_robots; i++)
      if ((size_t)k.robot_ids[i] == robot) {
        kb = &k;
        pos = (int)i;
      }
  if (!kb) return fail(VSR_ERR_STATE, "trajectory bucket not found");
  const int nt = (int)b->ticks.size();
  const size_t n = (size_t)nt * (size_t)kb->shape.n_readings;
  if (capacity < n) return fail(VSR_ERR_INVALID_ARGUMENT, "sensor readings buffer too small");
  if (n == 0) return 0;
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaStreamSynchronize(b->run_stream));
  if (b->precision == VSR_PRECISION_F64) {
    CUDA_TRY(cudaMemcpy(out, (double*)kb->d_rtrace + (size_t)pos * n, n * sizeof(double), cudaMemcpyDeviceToHost));
  } else {
    std::vector<float> tmp(n);
    CUDA_TRY(cudaMemcpy(tmp.data(), (float*)kb->d_rtrace + (size_t)pos * n, n * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) out[i] = tmp[i];
  }
  return (long long)n;
}

extern "C" size_t vsr_batch_voxel_steps(const vsr_batch* b) { return b ? b->voxel_steps : 0; }
extern "C" int vsr_batch_n_ticks(const vsr_batch* b) { return b ? (int)b->ticks.size() : 0; }
extern "C" int vsr_batch_n_bodies(const vsr_batch* b, size_t robot) {
  if (!b || robot >= b->n_robots) return 0;
  return 4 * b->robot_nvox[robot];
}
extern "C" int vsr_batch_last_launches(const vsr_batch* b) { return b ? b->last_launches : 0; }
extern "C" size_t vsr_batch_upload_bytes(const vsr_batch* b) { return b ? b->upload_bytes : 0; }

#ifdef VSR_CCD_COUNT  // development builds only
extern "C" void vsr_dev_counters(unsigned long long* out) {
  cudaMemcpyFromSymbol(&out[0], g_ccd_count, 8);
  cudaMemcpyFromSymbol(&out[1], g_ccd_count2, 8);
  cudaMemcpyFromSymbol(&out[2], g_ccd_count3, 8);
  cudaMemcpyFromSymbol(&out[3], g_ccd_hist, 32);
}
#endif
