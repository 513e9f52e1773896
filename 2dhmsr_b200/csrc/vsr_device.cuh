// vsr_device.cuh -- the episode kernel of the B200-native batched VSR stepper.
//
// Replaces, for N independent robots at once, the per-tick path
//   AbstractTask.updateWorld (tasks/AbstractTask.java:41-64):
//     World.step(1)  [org.dyn4j 4.2.1]  +  Robot.act(t) (core/objects/Robot.java:73-77)
// inside Locomotion.apply's loop (tasks/locomotion/Locomotion.java:196-203).
//
// Mapping:
//   * one LANE per voxel, floor(32 / n_voxels) robots per warp (one block per robot above 32
//     voxels), a warp runs the whole episode (all ticks) of its robots; the warps of a block walk
//     the (large, unrolled) instruction stream together behind data-free barriers;
//   * the 4 masses of a voxel (pose + velocity) and its 18 spring-dampers
//     (Voxel.java:239-479) live in that lane's registers: the 18 soft DistanceJoint
//     solves of a velocity iteration are pure register arithmetic, fully unrolled;
//   * the rigid WeldJoints between neighbouring voxels (Robot.java:91-114) are solved once, by
//     the lane of body 1, after gathering the partner mass' state with warp shuffles; the impulse
//     is handed to the lane of body 2: first all left/right welds, then all up/down welds (each
//     mass has at most one of each => a 2-colouring);
//   * contacts -- a mass against a ground polygon (Ground.java:64-70) or against another mass of
//     the robot (Voxel.java:178-185,474) -- are one record type: SAT + clipping manifold,
//     sequential impulses with the 2-point block solver, Baumgarte position pass; the constraints
//     of a robot scope are listed per tick by dependency level and solved one per lane;
//   * sensors (Touch/AreaRatio/Velocity/Angle/... with Average/Trend windows, the normalisations
//     and Noisy) and controllers (PhaseSin, tabulated TimeFunctions, centralized or distributed
//     per-robot MLPs) run in the same kernel right after the step.
// The resulting Gauss-Seidel order (springs -> horizontal welds -> vertical welds ->
// contacts) is the oracle's VSR_ORACLE_ORDER_CANONICAL.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

// margin of the geometric tie rules in the fp32 build (1e-12 in fp64), see collide_mass_segment
#ifndef VSR_TIE_EPS_F32
#define VSR_TIE_EPS_F32 1e-6f
#endif

// candidate mass pairs per thread and tick that pass the broadphase of the intra-robot contacts (8 overflowed for
// 7 of 131 072 robots of BASELINE config 4: tall bipeds that fold up; 16 has not been seen to)
#define VSR_MAX_CAND 16

namespace vsr {

constexpr int MAX_VOX = 128;   // voxels per robot (<= 32: robots share a warp; else one block per robot)
constexpr int MAX_SENS = 16;   // sensors per voxel (uniformAll-*: all 15 predefined ones)
constexpr int MAX_READ = 24;   // readings per voxel (uniformAll-*: 21)
constexpr int MAXC = 3;        // simultaneous ground constraints per mass (vsr_batch_create rejects terrains that could need more)
constexpr int MAX_LAYERS = 6;  // MLP layers (input + hidden + output)
#define VSR_DIST_WIDTH 40      // DistributedSensing: widest layer of a voxel's MLP (24 readings + 4 x 4 signals = 40 inputs)
constexpr int MAX_WINDOWS = 2; // distinct aggregator intervals per shape

enum : int { CTRL_PHASE_SIN = 0, CTRL_TABULATED = 1, CTRL_MLP = 2 };
enum : int { SENSOR_TOUCH = 0, SENSOR_AREA = 1, SENSOR_VELOCITY = 2, SENSOR_ANGLE = 3, SENSOR_CONSTANT = 4,
             SENSOR_APPLIED_FORCE = 5, SENSOR_MALFUNCTION = 6, SENSOR_TIME_SIN = 7, SENSOR_CRUMPLING = 8,
             SENSOR_CONTROL_POWER = 9, SENSOR_LIDAR = 10 };
enum : int { AGG_NONE = 0, AGG_AVERAGE = 1, AGG_TREND = 2, AGG_DYNAMIC_NORM = 3 };
// vsr_result.status bits (include/vsr_b200.h): the three capacity bits mean that constraints were DROPPED for
// this robot (its result is not the reference's physics)
enum : int { STATUS_NONFINITE = 1, STATUS_CONTACT_OVERFLOW = 2, STATUS_PAIR_OVERFLOW = 4, STATUS_LEVEL_OVERFLOW = 8 };

struct SensorDev {
  int8_t base, axes, rotated, agg, soft, win, ch, dim;
  float interval;
  double max_norm;  // Velocity: domain bound; Constant / TimeFunction: the value / the frequency
  double noise;     // Noisy: sigma (0 = not noisy)
  double dirs[8];   // Lidar: ray directions (axes = their number, max_norm = the ray length)
};

static_assert(offsetof(SensorDev, interval) == 8 && offsetof(SensorDev, max_norm) == 16 && sizeof(SensorDev) % 8 == 0,
              "the kernel reads the head of a SensorDev as two 8-byte words");
// the scalars of a SensorDev, unpacked in registers
struct SensorHdr {
  int base, axes, rotated, agg, soft, win, ch, dim;
  float interval;
  double max_norm, noise;
};

// One shape class (Grid<Boolean> + sensor layout), compiled on the host.
struct ShapeDev {
  int nvox;
  int n_readings;  // CentralizedSensing.nOfInputs
  int n_channels;  // aggregator ring channels per voxel
  int ring_depth;
  int8_t nbr[4][MAX_VOX];  // neighbour voxel index: left, right, down, up (-1 = none)
  double init_x[MAX_VOX * 4], init_y[MAX_VOX * 4];  // placed masses, local frame
  int8_t n_sens[MAX_VOX];
  int16_t read_off[MAX_VOX];
  SensorDev sens[MAX_VOX][MAX_SENS];
  int n_layers;
  int neurons[MAX_LAYERS];
  int layer_off[MAX_LAYERS];  // offset of layer i's (transposed) weights
  int n_params;               // controller parameters per robot (device layout)
  int dist_off[MAX_VOX];      // DistributedSensing: offset of each voxel's weights in the robot's parameters
  // BreakableVoxel parameters per voxel (vsr_breakable); read by the MODE 2 instantiation only
  int has_breakable;
  int8_t brk_enabled[MAX_VOX];
  uint8_t brk_mal[MAX_VOX][4];   // malfunction type masks of ACTUATOR, SENSORS, STRUCTURE
  double brk_thr[MAX_VOX][3];    // thresholds of CONTROL, AREA, TIME (NaN = absent)
  double brk_restore[MAX_VOX];
  long long brk_seed[MAX_VOX];
};

template <typename real>
struct Params {
  const ShapeDev* shape;
  // robots of this bucket
  long long n_robots;
  const int* robot_ids;       // bucket index -> batch index
  const real* params;         // [n_robots][n_params] (bucket order)
  const real* place_off;      // [n_robots][2] offset of the robot's placement from the shape's default, or null
  double* results;            // vsr_result records, batch order (11 doubles each)
  double* final_bbox;         // [n_robots][4] Robot.boundingBox at the last tick (min x, min y, max x, max y), batch order
  real* ring;                 // [ring_depth][n_channels][n_robots * nvox]
  real* trajectory;           // [traj_robots][n_ticks][nvox*4][6], bucket order, or null
  long long traj_robots;
  real* vtrace;               // [traj_robots][n_ticks][nvox][2]: touching the ground, applied force
  real* rtrace;               // [traj_robots][n_ticks][n_readings]: Voxel.getSensorReadings, voxels in order
  // time
  int n_ticks;
  const double* tick_t;       // [n_ticks] accumulated t
  const int16_t* win_first;   // [MAX_WINDOWS][n_ticks] oldest kept tick
  const double* noise_tab;    // [2 * n_ticks] java.util.Random(seed).nextGaussian() stream (Noisy sensors)
  // terrain (local frame: x - origin_x)
  const real* terr_x;
  const real* terr_y;
  int terr_n;
  // coarse height map of the terrain: the highest point of the profile over each cell of width 1 / hgrid_inv
  // (local frame, from hgrid_x0): a mass wholly above it cannot touch the ground
  const real* hgrid;
  int hgrid_n;
  real hgrid_x0, hgrid_inv;
  real base_y;
  double origin_x;
  // settings / material
  real dt, grav_x, grav_y, lin_damp, ang_damp;
  real inv_m, inv_i, half, side;
  real lin_tol, ang_tol, max_corr, baumgarte, rest_vel, max_trans, max_rot;
  real mu, rest_e;
  real spring_w, spring_zeta, red_mass;
  real rng_min[3], rng_rest[3], rng_max[3];  // side-parallel, side-cross, central-cross
  real weld_len;                             // |localAnchor2| = mass side
  int vel_iters, pos_iters;
  int spring_mode;       // 0 reduced mass, 1 effective mass
  unsigned spring_mask;  // bit s = spring s present (scaffoldings)
  // passive-range ("rope") limits of the springs, per range class; only the LIM instantiation reads them
  int lim_lo_on, lim_hi_on;
  real lim_lo[3], lim_hi[3];
  int mlp_act;
  int dist_signals;      // DistributedSensing: signals per side; < 0 for the other controllers
  int dist_nd;           // DistributedSensingNonDirectional: one set of signals for all neighbours
  int smem_per_warp;     // reals of staging per warp
  int wpr, rpb;          // one-block-per-robot variant: warps per robot, robots per block (they walk the code together)
  void* cbuf;            // contact arena: [grid threads][4 masses][2 buffers] CBody<real>
  void* sbuf;            // intra-robot contact arena: [grid threads][2 buffers] SList<real>
  int* lbuf;             // contact lists: [grid blocks][VSR_ITEMS_PER_THREAD * max threads per block]
  int pool_cap;          // constraints of a block cached in shared memory (the rest stay in global)
  int ccd;               // vsr_settings.continuous_detection: the time-of-impact pass after the position solve
  real* ccd_buf;         // [36][grid threads]: poses at the start of the step (Body.transform0), poses now, velocities
  int self_coll;         // masses of one robot collide with each other (vsr_settings.self_collision)
  real mu_self, rest_e_self;
  int ctrl;              // CTRL_*
  int sync_mode;         // optional lockstep barriers (code locality only): 2 per velocity iteration, 8 inside it,
                         // 16 after the spring init, 32 after the velocity loop, 64 after the position loop.
                         // The barriers at the top of a tick and before Robot.act are unconditional: they also
                         // separate the phases that share the block-wide scratch `aux`.
  long long debug_robot; // bucket index of a robot whose manifolds are printed (-1 = off)
  int debug_tick, debug_tick_end;
};

// ---------------------------------------------------------------- math wrappers
__device__ __forceinline__ float r_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double r_sqrt(double x) { return sqrt(x); }
// fp32 sincos: Cody-Waite reduction by pi/2 (3 constants) + the classic degree-7/8 minimax
// polynomials on [-pi/4, pi/4]; ~1 ulp for |a| < 1e4 rad (a mass never spins that far) in ~25
// instructions -- sincosf() inlines ~85 with a slow path, and there are many call sites.
__device__ __forceinline__ void r_sincos(float a, float* s, float* c) {
  float q = rintf(a * 0.636619772f);
  float r = fmaf(q, -1.57079601e+00f, a);
  r = fmaf(q, -3.13916473e-07f, r);
  r = fmaf(q, -5.39030253e-15f, r);
  int i = (int)q;
  float r2 = r * r;
  float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f), r2 * r, r);
  float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f), r2 * r2,
                  fmaf(-0.5f, r2, 1.0f));
  float ss = (i & 1) ? cp : sp, cc = (i & 1) ? sp : cp;
  *s = (i & 2) ? -ss : ss;
  *c = ((i + 1) & 2) ? -cc : cc;
}
__device__ __forceinline__ void r_sincos(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ float r_atan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double r_atan2(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float r_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ double r_tanh(double x) { return tanh(x); }
__device__ __forceinline__ float r_exp(float x) { return expf(x); }
__device__ __forceinline__ double r_exp(double x) { return exp(x); }
__device__ __forceinline__ float r_sin(float x) { return sinf(x); }
__device__ __forceinline__ double r_sin(double x) { return sin(x); }
__device__ __forceinline__ float r_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double r_abs(double x) { return fabs(x); }
__device__ __forceinline__ float r_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double r_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float r_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double r_max(double a, double b) { return fmax(a, b); }
template <typename real>
__device__ __forceinline__ real r_clamp(real x, real lo, real hi) { return r_min(r_max(x, lo), hi); }

template <typename real>
__device__ __forceinline__ real shfl(real v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Named barriers (ids 1..15; 0 is __syncthreads): the threads of ONE robot of a block that holds several big
// robots.  `n` = the robot's thread count (a multiple of 32); every one of them arrives.
__device__ __forceinline__ void robot_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ bool robot_sync_and(int id, int n, bool p) {
  int r;
  asm volatile(
      "{\n\t.reg .pred q, r;\n\tsetp.ne.b32 q, %3, 0;\n\tbar.red.and.pred r, %1, %2, q;\n\tselp.b32 %0, 1, 0, r;\n\t}"
      : "=r"(r)
      : "r"(id), "r"(n), "r"((int)p)
      : "memory");
  return r != 0;
}

// java.util.Random (48-bit LCG): next(bits), nextDouble, nextInt(bound) -- BreakableVoxel draws from one per voxel
__device__ __forceinline__ long long jr_next(unsigned long long& seed, int bits) {
  seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
  return (long long)(seed >> (48 - bits));
}
__device__ __forceinline__ double jr_next_double(unsigned long long& seed) {
  const long long a = jr_next(seed, 26), b = jr_next(seed, 27);
  return (double)((a << 27) + b) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ int jr_next_int(unsigned long long& seed, int bound) {
  if ((bound & -bound) == bound) return (int)(((long long)bound * jr_next(seed, 31)) >> 31);
  for (;;) {
    const long long bits = jr_next(seed, 31), val = bits % bound;
    if (bits - val + (bound - 1) < (1LL << 31)) return (int)val;
  }
}

// 4-wide value for one 128-bit (fp32) shared-memory access per spring
template <typename real> struct Real4T;
template <> struct Real4T<float> { typedef float4 type; };
template <> struct Real4T<double> { typedef double4 type; };
template <typename real> struct Real2T;
template <> struct Real2T<float> { typedef float2 type; };
template <> struct Real2T<double> { typedef double2 type; };

// ---------------------------------------------------------------- contact constraints
// One record type for both kinds: a mass against a ground segment (tb < 0; seg = the segment) and a
// mass against another mass of the same robot (Voxel.java:178-185,474; owned by the thread of the
// lower body index, ka = its mass; the partner is mass kb of block thread tb).
template <typename real>
struct SPoint {
  real px, py, depth, jn, jt;  // persistent (warm start by id)
  int id;
  real r1x, r1y, r2x, r2y, l1x, l1y, l2x, l2y, massN, massT, vb;  // per step
};
template <typename real>
struct SCon {
  int ka, tb, kb, n, block, seg;
  real nx, ny;  // manifold normal: from body 1 (the owner's mass) to body 2
  SPoint<real> p[2];
  real K00, K01, K11, iK00, iK01, iK11;
};
template <typename real>
struct CBody {  // ground constraints of one mass
  int n;
  SCon<real> c[MAXC];
};
template <typename real>
struct SList {  // intra-robot constraints owned by one thread
  SCon<real> c[VSR_MAX_CAND];  // slot q = the thread's q-th candidate pair of the tick (a bit mask says which hold)
};

template <typename real>
struct Poly4 {
  real vx[4], vy[4], nx[4], ny[4];
};
template <typename real>
struct Edge {
  real x1, y1, x2, y2, mx, my;
  int i1, i2, index;
};

// a[i] for i in 0..3 as register selects: a dynamically indexed local array would live in local memory,
// and with the shared-memory carve-out of this kernel the L1 is too small to hold the stack frames
template <typename real>
__device__ __forceinline__ real sel4(const real* a, int i) {
  const real lo = (i & 1) ? a[1] : a[0], hi = (i & 1) ? a[3] : a[2];
  return (i & 2) ? hi : lo;
}

// Polygon.getFarthestFeature (dyn4j): farthest vertex along n, then the adjacent edge
// that is most perpendicular to n, CCW winding kept.
template <typename real>
__device__ __forceinline__ void farthest_edge(const Poly4<real>& P, real nx, real ny, Edge<real>& e) {
  int idx = 0;
  real best = P.vx[0] * nx + P.vy[0] * ny;
#pragma unroll
  for (int i = 1; i < 4; i++) {
    real v = P.vx[i] * nx + P.vy[i] * ny;
    if (v > best) {
      best = v;
      idx = i;
    }
  }
  const int prev = (idx + 3) & 3, next = (idx + 1) & 3;
  const real ln = sel4(P.nx, prev) * nx + sel4(P.ny, prev) * ny;
  const real rn = sel4(P.nx, idx) * nx + sel4(P.ny, idx) * ny;
  const real ix = sel4(P.vx, idx), iy = sel4(P.vy, idx);
  e.mx = ix;
  e.my = iy;
  if (ln < rn) {
    e.x1 = ix; e.y1 = iy; e.i1 = idx;
    e.x2 = sel4(P.vx, next); e.y2 = sel4(P.vy, next); e.i2 = next;
    e.index = next == 0 ? 4 : next;
  } else {
    e.x1 = sel4(P.vx, prev); e.y1 = sel4(P.vy, prev); e.i1 = prev;
    e.x2 = ix; e.y2 = iy; e.i2 = idx;
    e.index = idx == 0 ? 4 : idx;  // one label per edge (see the oracle's farthest_edge)
  }
}

template <typename real>
struct ClipPt {
  real x, y;
  int idx;
};

// ClippingManifoldSolver.clip (dyn4j): the points of (v1, v2) behind the plane, then the crossing point if
// the two lie on different sides -- at most two; returns how many of (o0, o1) hold.
template <typename real>
__device__ __forceinline__ int clip_edge(const ClipPt<real>& v1, const ClipPt<real>& v2, real nx, real ny, real offset,
                                         ClipPt<real>& o0, ClipPt<real>& o1) {
  const real d1 = nx * v1.x + ny * v1.y - offset, d2 = nx * v2.x + ny * v2.y - offset;
  const bool k1 = d1 <= real(0), k2 = d2 <= real(0), cross = d1 * d2 < real(0);
  const real u = d1 / (d1 - d2);
  ClipPt<real> p;
  p.x = v1.x + (v2.x - v1.x) * u;
  p.y = v1.y + (v2.y - v1.y) * u;
  p.idx = d1 > real(0) ? v1.idx : v2.idx;
  // k1 && k2 excludes a crossing; neither excludes it too (both in front)
  o0 = k1 ? v1 : (k2 ? v2 : p);
  o1 = (k1 && k2) ? v2 : p;
  return (k1 ? 1 : 0) + (k2 ? 1 : 0) + (cross ? 1 : 0);
}

// Narrowphase (SAT == the EPA answer for two convex polygons) + clipping manifold for
// one mass against one ground segment.  Returns the number of manifold points (0 = no collision) and fills c
// (n, normal, points).
template <typename real>
__device__ __noinline__ int collide_mass_segment(real bx, real by, real bc, real bs, real half, real x0, real y0,
                                                 real x1, real y1, real base_y, int mm, SCon<real>& c) {
  // exact geometric ties (a mass lying flat on a flat segment) are decided by rule, with this
  // margin, exactly as in the oracle: the mass' own face wins.
  // mm != 0: the second polygon is another mass, given as (x0, y0, x1, y1) = (centre, cos, sin);
  // neighbouring masses touch corner to corner at rest, so there every comparison carries the
  // margin and a contact needs a penetration above it (see the oracle's sat_penetration).
  const real tie_eps = sizeof(real) == 8 ? real(1e-12) : real(VSR_TIE_EPS_F32);
  const real in_eps = mm ? tie_eps : real(0);
  Poly4<real> A, B;
  const real lx[4] = {-1, 1, 1, -1}, ly[4] = {-1, -1, 1, 1};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    real px = lx[i] * half, py = ly[i] * half;
    A.vx[i] = bx + bc * px - bs * py;
    A.vy[i] = by + bs * px + bc * py;
  }
  A.nx[0] = bs; A.ny[0] = -bc;
  A.nx[1] = bc; A.ny[1] = bs;
  A.nx[2] = -bs; A.ny[2] = bc;
  A.nx[3] = -bc; A.ny[3] = -bs;
  if (mm) {
    const real c2 = x1, s2 = y1;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      real px = lx[i] * half, py = ly[i] * half;
      B.vx[i] = x0 + c2 * px - s2 * py;
      B.vy[i] = y0 + s2 * px + c2 * py;
    }
    B.nx[0] = s2; B.ny[0] = -c2;
    B.nx[1] = c2; B.ny[1] = s2;
    B.nx[2] = -s2; B.ny[2] = c2;
    B.nx[3] = -c2; B.ny[3] = -s2;
  } else {
    B.vx[0] = x0; B.vy[0] = y0;
    B.vx[1] = x0; B.vy[1] = base_y;
    B.vx[2] = x1; B.vy[2] = base_y;
    B.vx[3] = x1; B.vy[3] = y1;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int j = (i + 1) & 3;
      real ex = B.vx[j] - B.vx[i], ey = B.vy[j] - B.vy[i];
      real l = r_sqrt(ex * ex + ey * ey);
      B.nx[i] = ey / l;
      B.ny[i] = -ex / l;
    }
  }
  // SAT: least-negative separation over the face normals of both polygons
  real best = -1e30f, nx = 0, ny = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    real s = 1e30f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      real v = A.nx[i] * (B.vx[j] - A.vx[i]) + A.ny[i] * (B.vy[j] - A.vy[i]);
      s = v < s ? v : s;
    }
    if (s > best + (i > 0 ? in_eps : real(0))) { best = s; nx = A.nx[i]; ny = A.ny[i]; }
  }
  // `best` never decreases: separated along one of the first mass' axes is final (the common case
  // for the corner-to-corner neighbours)
  if (mm && !(best < -in_eps)) return 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    real s = 1e30f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      real v = B.nx[j] * (A.vx[i] - B.vx[j]) + B.ny[j] * (A.vy[i] - B.vy[j]);
      s = v < s ? v : s;
    }
    if (s > best + tie_eps) { best = s; nx = -B.nx[j]; ny = -B.ny[j]; }
  }
  if (!(best < -in_eps)) return 0;
  // ClippingManifoldSolver.getManifold (dyn4j)
  Edge<real> ref, inc;
  farthest_edge(A, nx, ny, ref);
  farthest_edge(B, -nx, -ny, inc);
  int flipped = 0;
  real e1 = r_abs((ref.x2 - ref.x1) * nx + (ref.y2 - ref.y1) * ny);
  real e2 = r_abs((inc.x2 - inc.x1) * nx + (inc.y2 - inc.y1) * ny);
  if (e1 > e2 + tie_eps) {
    Edge<real> t = ref; ref = inc; inc = t;
    flipped = 1;
  }
  real rx = ref.x2 - ref.x1, ry = ref.y2 - ref.y1;
  real rl = r_sqrt(rx * rx + ry * ry);
  rx /= rl;
  ry /= rl;
  ClipPt<real> a = {inc.x1, inc.y1, inc.i1}, b = {inc.x2, inc.y2, inc.i2}, c1[2], c2[2];
  real o1 = -(rx * ref.x1 + ry * ref.y1);
  if (clip_edge(a, b, -rx, -ry, o1, c1[0], c1[1]) < 2) return 0;
  real o2 = rx * ref.x2 + ry * ref.y2;
  if (clip_edge(c1[0], c1[1], rx, ry, o2, c2[0], c2[1]) < 2) return 0;
  real fx = -ry, fy = rx;
  real fo = fx * ref.mx + fy * ref.my;
  c.nx = flipped ? fx : -fx;
  c.ny = flipped ? fy : -fy;
  c.n = 0;
#pragma unroll
  for (int i = 0; i < 2; i++) {
    real depth = fx * c2[i].x + fy * c2[i].y - fo;
    if (depth >= real(0)) {
      SPoint<real>& p = c.p[c.n++];
      p.px = c2[i].x;
      p.py = c2[i].y;
      p.depth = depth;
      p.jn = 0;
      p.jt = 0;
      p.id = (ref.index & 7) | ((inc.index & 7) << 3) | ((c2[i].idx & 7) << 6) | (flipped << 9);
    }
  }
  return c.n;  // (the callers keep the count in a register: c lives in global memory)
}

// terrain description passed by value (registers) to detect_mass
template <typename real>
struct TPar {
  const real* terr_x;
  const real* terr_y;
  int terr_n;
  real half, base_y;
};
// Lidar.sense (Lidar.java:96-103) for one ray: World.raycast against the ground polygons (the filter
// drops every robot part).  A ray that starts above the terrain enters a ground polygon through its
// top edge, and the polygons share their side edges, so the first hit is the nearest crossing of the
// terrain polyline from above; none within `len` reads `len`.
template <typename real>
__device__ __noinline__ real lidar_cast(const TPar<real> P, real ox, real oy, real dx, real dy, real len, int hint) {
  const int n = P.terr_n;
  const real xlo = r_min(ox, ox + dx * len), xhi = r_max(ox, ox + dx * len);
  int s = hint;
  while (s > 0 && P.terr_x[s] > xlo) --s;
  real best = len;
  for (int seg = s; seg + 1 < n && P.terr_x[seg] <= xhi; ++seg) {
    const real x0 = P.terr_x[seg], y0 = P.terr_y[seg], x1 = P.terr_x[seg + 1], y1 = P.terr_y[seg + 1];
    if (!(x1 > x0)) continue;
    const real ex = x1 - x0, ey = y1 - y0;
    const real den = dx * ey - dy * ex;     // cross(D, E); > 0: the ray runs against the upward normal
    if (!(den > real(0))) continue;
    const real ax = x0 - ox, ay = y0 - oy;
    const real t = (ax * ey - ay * ex) / den;   // cross(A - O, E) / cross(D, E)
    const real u = (ax * dy - ay * dx) / den;   // cross(A - O, D) / cross(D, E)
    if (t >= real(0) && u >= real(0) && u <= real(1) && t < best) best = t;
  }
  return best;
}

// World.detect (dyn4j) for one mass: every ground segment its AABB overlaps; impulses of
// points with the same id are carried over (ContactConstraint.update).
template <typename real>
__device__ __noinline__ int detect_mass(const TPar<real> P, real bx, real by, real bc, real bs, int& hint,
                                        const CBody<real>& old, int had_contact, CBody<real>& cb, int& n_out,
                                        real im, real ii, int k_mass) {
  real e = P.half * (r_abs(bc) + r_abs(bs));
  real xlo = bx - e, xhi = bx + e, ylo = by - e;
  int n = P.terr_n;
  int s = hint;
  while (s > 0 && P.terr_x[s] > xlo) --s;
  while (s + 2 < n && P.terr_x[s + 1] < xlo) ++s;
  hint = s;
  const int old_n = had_contact ? old.n : 0;  // only masses already in contact need the warm-start match
  int cn = 0;
  int overflow = 0;
  for (int seg = s; seg + 1 < n && P.terr_x[seg] <= xhi; ++seg) {
    real x0 = P.terr_x[seg], x1 = P.terr_x[seg + 1], y0 = P.terr_y[seg], y1 = P.terr_y[seg + 1];
    if (x1 < xlo) continue;
    if (!(x1 > x0)) continue;
    if (ylo > r_max(y0, y1)) continue;
    // the record is built in place (the arena), not in thread-local memory; a scratch copy only takes the
    // manifold that would not fit
    SCon<real> scratch;
    SCon<real>& c = cn < MAXC ? cb.c[cn] : scratch;
    // (the point count in a register: reading c.n back from the arena is an L2 round trip before every loop)
    const int n_c = collide_mass_segment(bx, by, bc, bs, P.half, x0, y0, x1, y1, P.base_y, 0, c);
    if (!n_c) continue;
    if (cn == MAXC) { overflow = 1; break; }
    c.seg = seg;
    c.ka = k_mass; c.tb = -1; c.kb = 0;
    // ContactConstraint.update: impulses of the points with the same id are carried over.  Everything the match
    // needs is loaded in one batch (independent loads = one L2 round trip; walking seg -> n -> id -> impulse
    // were four) and compared in registers, in the order of the loops it replaces (later matches win).
    if (old_n > 0) {
      const int idn[2] = {c.p[0].id, c.p[1].id};
      real jn[2], jt[2];
      bool hit[2] = {false, false};
#pragma unroll
      for (int i = 0; i < MAXC; i++) {
        const SCon<real>& o = old.c[i];
        const int oseg = o.seg, on = o.n, oid[2] = {o.p[0].id, o.p[1].id};
        const real ojn[2] = {o.p[0].jn, o.p[1].jn}, ojt[2] = {o.p[0].jt, o.p[1].jt};
        if (i < old_n && oseg == seg) {
#pragma unroll
          for (int k = 0; k < 2; k++)
#pragma unroll
            for (int l = 0; l < 2; l++)
              if (k < n_c && l < on && oid[l] == idn[k]) { jn[k] = ojn[l]; jt[k] = ojt[l]; hit[k] = true; }
        }
      }
#pragma unroll
      for (int k = 0; k < 2; k++)
        if (hit[k]) { c.p[k].jn = jn[k]; c.p[k].jt = jt[k]; }
    }
    // SequentialImpulses.initialize (dyn4j), the part that depends on positions only (the ground
    // is static: invM2 = invI2 = 0): the mass does not move between this detection and the next
    // step's initialisation
    {
      const real nx = c.nx, ny = c.ny, tx = ny, ty = -nx;
#pragma unroll
      for (int k = 0; k < 2; k++) {
        if (k >= n_c) break;
        SPoint<real>& p = c.p[k];
        p.r1x = p.px - bx;
        p.r1y = p.py - by;
        p.l1x = bc * p.r1x + bs * p.r1y;
        p.l1y = -bs * p.r1x + bc * p.r1y;
        p.r2x = real(0); p.r2y = real(0);  // static body 2: never multiplied by anything but 0
        p.l2x = p.px; p.l2y = p.py;        // ... at the origin, unrotated
        real rn1 = p.r1x * ny - p.r1y * nx;
        p.massN = real(1) / (im + ii * rn1 * rn1);
        real rt1 = p.r1x * ty - p.r1y * tx;
        p.massT = real(1) / (im + ii * rt1 * rt1);
        p.vb = real(0);
      }
      c.block = 0;
      if (n_c == 2) {
        real rn1A = c.p[0].r1x * ny - c.p[0].r1y * nx;
        real rn2A = c.p[1].r1x * ny - c.p[1].r1y * nx;
        c.K00 = im + ii * rn1A * rn1A;
        c.K01 = im + ii * rn1A * rn2A;
        c.K11 = im + ii * rn2A * rn2A;
        real det = c.K00 * c.K11 - c.K01 * c.K01;
        if (c.K00 * c.K00 < real(1000) * det) {
          c.block = 1;
          real id = real(1) / det;
          c.iK00 = id * c.K11;
          c.iK01 = -id * c.K01;
          c.iK11 = id * c.K00;
        } else {
          c.n = 1;
        }
      }
    }
    cn++;
  }
  if (cn > 0 || had_contact) cb.n = cn;
  n_out = cn;
  return overflow;
}

// ---------------------------------------------------------------- continuous collision detection
// World.solveTOI for the masses against the ground (dyn4j, ContinuousDetectionMode.ALL; the oracle's ccd_pass): a mass
// moved from pose0 by (v dt, w dt); against every ground polygon its swept box overlaps and it is NOT in contact with
// (the constraints of the last detection), conservative advancement finds the first time in [0, 1] at which the two come
// within distanceEpsilon; the mass is put back to that time (transform0.lerp) and TimeOfImpactSolver pushes it
// linearTolerance into the polygon.  Velocities are not touched.
//
// ONE WARP works on ONE mass at a time.  Conservative advancement is a serial chain of up to 30 distance evaluations,
// and few masses need it (0.4 per robot and tick on BASELINE config 2), so spreading the masses over the lanes leaves a
// block waiting for the one lane with the longest chain (measured: 2.6 x the run time of the step).  Instead the 32 lanes
// share each evaluation: a distance between two quads is the minimum over 2 x 4 x 4 vertex-edge pairs -- one pair per
// lane -- and its SAT test 2 x 4 x 4 projections; minima and the first-minimum rule travel by shuffles.  Everything
// scalar (the advancement itself) is computed redundantly by all lanes.  The pass runs in DOUBLE in the fp32 build too:
// its decisions hang on distances of 6e-6 (distanceEpsilon), a dozen ulps of a float coordinate -- in float the
// advancement does not converge and a population slows down by a quarter.
#ifdef VSR_CCD_COUNT
__device__ unsigned long long g_ccd_count, g_ccd_count2, g_ccd_count3, g_ccd_hist[4];
#endif
template <typename real>
struct Pose3 {
  real x, y, th;
};
template <typename cr>
__device__ __forceinline__ void ccd_mass_vertex(int i, cr x, cr y, cr hc, cr hs, cr& vx, cr& vy) {
  // corners (-h,-h) (+h,-h) (+h,+h) (-h,+h) of the rotated square: x + sx hc - sy hs, y + sx hs + sy hc
  const cr sx = (i == 1 || i == 2) ? cr(1) : cr(-1), sy = (i >= 2) ? cr(1) : cr(-1);
  vx = x + sx * hc - sy * hs;
  vy = y + sx * hs + sy * hc;
}
template <typename cr>
__device__ __forceinline__ void ccd_ground_vertex(int j, cr x0, cr y0, cr x1, cr y1, cr base_y, cr& vx, cr& vy) {
  vx = j < 2 ? x0 : x1;  // (x0, y0) (x0, base) (x1, base) (x1, y1): Ground.java:64-70
  vy = j == 0 ? y0 : (j == 3 ? y1 : base_y);
}
template <typename cr>
__device__ __forceinline__ cr shfl_d(cr v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <typename cr>
__device__ __forceinline__ cr shfl_xor_d(cr v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
// the ground polygon of a segment, as the per-lane constants of the evaluation (lane = pass * 16 + i * 4 + j)
template <typename cr>
struct CcdGround {
  cr x0, y0, x1, y1, base_y;
  cr gnx, gny;  // normal of THIS lane's SAT axis when it is one of the polygon's (lanes 16..31: axis (lane >> 2) & 3)
  cr inv_e;         // 1 / squared length of this lane's edge (pass 0: ground edge j; pass 1: the mass' edge)
};
// distance of the mass at (x, y), rotated by the angle whose sine and cosine are (sn, cs), from the polygon G; all lanes get the same answer (false: they overlap)
template <typename cr>
__device__ __forceinline__ bool ccd_distance(int lane, cr x, cr y, cr sn, cr cs, cr half, const CcdGround<cr>& G, bool sat, cr& dist,
                                             cr& p1x, cr& p1y, cr& p2x, cr& p2y) {
  const cr hc = half * cs, hs = half * sn;
  const int pass = lane >> 4, i = (lane >> 2) & 3, j = lane & 3;
  // (margins: the oracle's in double; the float values belong to the experiment described in ccd_mass_warp)
  const cr tie = sizeof(cr) == 8 ? cr(1e-12) : cr(1e-6), touch2 = sizeof(cr) == 8 ? cr(1e-18) : cr(1e-12);
  if (sat) {  // (warp-uniform)
  // ---- SAT (the oracle's sat_penetration): axis i of the mass (lanes 0..15) / of the polygon (16..31), vertex j of the other
  cr proj;
  {
    cr ax, ay, gx, gy;
    if (pass == 0) {
      const cr nx = i == 0 ? sn : (i == 1 ? cs : (i == 2 ? -sn : -cs)), ny = i == 0 ? -cs : (i == 1 ? sn : (i == 2 ? cs : -sn));
      ccd_mass_vertex(i, x, y, hc, hs, ax, ay);
      ccd_ground_vertex(j, G.x0, G.y0, G.x1, G.y1, G.base_y, gx, gy);
      proj = nx * (gx - ax) + ny * (gy - ay);
    } else {
      ccd_ground_vertex(i, G.x0, G.y0, G.x1, G.y1, G.base_y, gx, gy);
      ccd_mass_vertex(j, x, y, hc, hs, ax, ay);
      proj = G.gnx * (ax - gx) + G.gny * (ay - gy);
    }
  }
  proj = r_min(proj, shfl_xor_d(proj, 1));
  proj = r_min(proj, shfl_xor_d(proj, 2));  // every group of 4 lanes: the axis' smallest projection
  cr best = cr(-1e30);
#pragma unroll
  for (int a = 0; a < 4; a++) {            // the mass' axes: the first maximum
    const cr sm = shfl_d(proj, a * 4);
    if (sm > best) best = sm;
  }
#pragma unroll
  for (int a = 0; a < 4; a++) {            // the polygon's: only by more than the tie margin
    const cr sm = shfl_d(proj, 16 + a * 4);
    if (sm > best + tie) best = sm;
  }
  if (best < cr(0)) return false;            // penetrating (uniform)
  }
  // ---- closest points (the oracle's poly_distance): pass 0 = vertex i of the mass against edge j of the polygon,
  //      pass 1 = vertex i of the polygon against edge j of the mass; the first minimum in lane order
  // (branch-free: the two halves of the warp would otherwise run their sides of an if one after the other -- a mass
  //  vertex is four multiply-adds, a ground vertex two selects, so every lane builds both candidates and selects)
  cr px, py, ex0, ey0, ex1, ey1;
  {
    cr mpx, mpy, gpx, gpy, m0x, m0y, m1x_, m1y_, g0x, g0y, g1x, g1y;
    ccd_mass_vertex(i, x, y, hc, hs, mpx, mpy);
    ccd_ground_vertex(i, G.x0, G.y0, G.x1, G.y1, G.base_y, gpx, gpy);
    ccd_mass_vertex(j, x, y, hc, hs, m0x, m0y);
    ccd_mass_vertex((j + 1) & 3, x, y, hc, hs, m1x_, m1y_);
    ccd_ground_vertex(j, G.x0, G.y0, G.x1, G.y1, G.base_y, g0x, g0y);
    ccd_ground_vertex((j + 1) & 3, G.x0, G.y0, G.x1, G.y1, G.base_y, g1x, g1y);
    const bool p0_ = pass == 0;
    px = p0_ ? mpx : gpx; py = p0_ ? mpy : gpy;
    ex0 = p0_ ? g0x : m0x; ey0 = p0_ ? g0y : m0y;
    ex1 = p0_ ? g1x : m1x_; ey1 = p0_ ? g1y : m1y_;
  }
  const cr ex = ex1 - ex0, ey = ey1 - ey0;
  cr t = ((px - ex0) * ex + (py - ey0) * ey) * G.inv_e;
  t = t < cr(0) ? cr(0) : (t > cr(1) ? cr(1) : t);
  const cr qx = ex0 + t * ex, qy = ey0 + t * ey;
  const cr dx = px - qx, dy = py - qy;
  cr dd = dx * dx + dy * dy;
  // the first minimum in lane order: dd >= 0, so its bit pattern orders like an unsigned integer -- two warp-wide integer
  // minima (high word, then low word among the lanes that hold the smallest high word) and a ballot replace five rounds
  // of three shuffles
  int who;
  if (sizeof(cr) == 8) {
    const unsigned long long key = (unsigned long long)__double_as_longlong((double)dd);
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    who = __ffs((int)__ballot_sync(0xffffffffu, hi == mh && lo == ml)) - 1;
    dd = (cr)__longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
  } else {
    const unsigned key = __float_as_uint((float)dd);
    const unsigned mk = __reduce_min_sync(0xffffffffu, key);
    who = __ffs((int)__ballot_sync(0xffffffffu, key == mk)) - 1;
    dd = (cr)__uint_as_float(mk);
  }
  if (dd <= touch2) return false;  // touching = not separated (the oracle's poly_distance says why the margin matters; uniform)
  // (on the mass: p1; on the polygon: p2)
  const cr m1x = pass == 0 ? px : qx, m1y = pass == 0 ? py : qy, m2x = pass == 0 ? qx : px, m2y = pass == 0 ? qy : py;
  p1x = shfl_d(m1x, who); p1y = shfl_d(m1y, who); p2x = shfl_d(m2x, who); p2y = shfl_d(m2y, who);
  dist = r_sqrt(dd);
  return true;
}
// The pass for ONE mass, executed by a whole (converged) warp; every argument is warp-uniform.
template <typename real>
__device__ __noinline__ Pose3<real> ccd_mass_warp(const TPar<real> P, Pose3<real> p0r, Pose3<real> p1r, real vxr, real vyr, real wr,
                                                  real dtr, const CBody<real>* cur, int has_contacts, real lin_tol_r,
                                                  real max_corr_r, real imr, real iir, int hint) {
  // All of it in double, in both builds.  (Tried in float, in a frame local to the mass so that 24 bits resolve 1e-7:
  // 1.7 x faster per candidate, but the touching / overlapping decisions near the surface go the other way often enough
  // to lower the population's mean displacement by 23 % -- conservative advancement has to resolve distanceEpsilon =
  // 6e-6, and its exits at "touching" need the margins of double.)
  typedef double cr;
  const int lane = threadIdx.x & 31;
  const cr eps_e = cr(2.220446049250313e-16);
  const cr dist_eps = cr(6.0554544523933395e-6);           // cbrt(Epsilon.E)
  const cr half = (cr)P.half, base_y = (cr)P.base_y;
  const cr rmax = half * cr(1.4142135623730951);            // Rectangle.getRadius
  const cr dt = (cr)dtr, lin_tol = (cr)lin_tol_r, max_corr = (cr)max_corr_r, im = (cr)imr, ii = (cr)iir;
  const Pose3<cr> p0 = {(cr)p0r.x, (cr)p0r.y, (cr)p0r.th}, p1 = {(cr)p1r.x, (cr)p1r.y, (cr)p1r.th};
  const cr dpx = (cr)vxr * dt, dpy = (cr)vyr * dt, da = (cr)wr * dt;
  cr minx, maxx, miny, s0, c0;
  {
    // swept box: an oriented square spans centre +- half (|cos| + |sin|) at either pose
    cr s1, c1;
    r_sincos(p0.th, &s0, &c0);
    r_sincos(p1.th, &s1, &c1);
    const cr e0 = half * (r_abs(c0) + r_abs(s0)), e1 = half * (r_abs(c1) + r_abs(s1));
    minx = r_min(p0.x - e0, p1.x - e1);
    maxx = r_max(p0.x + e0, p1.x + e1);
    miny = r_min(p0.y - e0, p1.y - e1);
  }
  const int n = P.terr_n;
  int s = hint;
  while (s > 0 && (cr)P.terr_x[s] > minx) --s;
  cr t2 = cr(1), best_d = cr(0), bp1x = cr(0), bp1y = cr(0), bp2x = cr(0), bp2y = cr(0);
  bool found = false;
  const int n_cur = has_contacts ? cur->n : 0;
#pragma unroll 1
  for (int seg = s; seg + 1 < n && (cr)P.terr_x[seg] <= maxx; ++seg) {
    const cr x0 = (cr)P.terr_x[seg], x1 = (cr)P.terr_x[seg + 1], y0 = (cr)P.terr_y[seg], y1 = (cr)P.terr_y[seg + 1];
    if (x1 < minx || !(x1 > x0) || miny > r_max(y0, y1)) continue;
    {
      // The polygon lies under the line of its top edge: the distance of the mass from the polygon at t = 0 is at
      // least that of the square from that half plane (centre's height above the line minus the square's support
      // along the line's normal).  When that gap is wider than anything the mass can travel in the step, the first
      // advance dist / (dp.n + amax) of the conservative advancement below is > 1 >= t2 and the polygon yields no time
      // of impact -- without evaluating a distance.
      const cr ex = x1 - x0, ey = y1 - y0, el = r_sqrt(ex * ex + ey * ey);
      const cr lnx = -ey / el, lny = ex / el;  // the line's normal, upwards
      const cr gap = (p0.x - x0) * lnx + (p0.y - y0) * lny - half * (r_abs(c0 * lnx + s0 * lny) + r_abs(c0 * lny - s0 * lnx));
      if (gap > (r_sqrt(dpx * dpx + dpy * dpy) + rmax * r_abs(da)) * cr(1.000001) + cr(1e-9)) continue;
    }
    bool in_contact = false;  // body1.isInContact(body2)
#pragma unroll
    for (int i = 0; i < MAXC; i++)
      if (i < n_cur && cur->c[i].seg == seg) in_contact = true;
    if (in_contact) continue;
    // this lane's constants of the polygon: its SAT axis (edge a -> a + 1: left, bottom, right, top) and its edge
    CcdGround<cr> G;
    G.x0 = x0; G.y0 = y0; G.x1 = x1; G.y1 = y1; G.base_y = base_y;
    {
      const int a = (lane >> 2) & 3, b = (a + 1) & 3;
      cr ax, ay, bx, by;
      ccd_ground_vertex(a, x0, y0, x1, y1, base_y, ax, ay);
      ccd_ground_vertex(b, x0, y0, x1, y1, base_y, bx, by);
      const cr ex = bx - ax, ey = by - ay, l = r_sqrt(ex * ex + ey * ey);
      G.gnx = ey / l;
      G.gny = -ex / l;
      const int j = lane & 3, k = (j + 1) & 3;
      cr cx, cy, dx, dy;
      ccd_ground_vertex(j, x0, y0, x1, y1, base_y, cx, cy);
      ccd_ground_vertex(k, x0, y0, x1, y1, base_y, dx, dy);
      G.inv_e = (lane >> 4) == 0 ? cr(1) / ((dx - cx) * (dx - cx) + (dy - cy) * (dy - cy)) : cr(1) / (cr(4) * half * half);
    }
    // ConservativeAdvancement.getTimeOfImpact on [0, t2], as ONE loop around ONE evaluation of the distance:
    // stage 0 = at t = 0, 1 = after an advance, 2 = after backing up from an overshoot
    const cr amax = rmax * r_abs(da);
    cr l = cr(0), l0 = cr(0), adv = cr(0), dist = cr(0), q1x = cr(0), q1y = cr(0), q2x = cr(0), q2y = cr(0);
    bool ok = true, skip = false;
    int iter = 0, stage = 0;
#pragma unroll 1
    for (;;) {
      // (After an advance the separating-axis test is not needed: the advance is conservative -- the separation along
      // the current normal shrinks by at most dp.n + amax per unit of time, so the shapes cannot overlap by more than
      // rounding, and that case is the "touching" exit of the distance itself.)
      cr sn_l, cs_l;  // (six series terms for sin / cos of da l and the addition theorem instead of this sincos: no gain)
      r_sincos(p0.th + da * l, &sn_l, &cs_l);
      const bool sep = ccd_distance(lane, p0.x + dpx * l, p0.y + dpy * l, sn_l, cs_l, half, G, stage != 1, dist, q1x, q1y, q2x, q2y);
#ifdef VSR_CCD_COUNT
      if (lane == 0) atomicAdd(&g_ccd_count3, 1ull);
#endif
      if (stage == 0) {
        if (!sep) { skip = true; break; }  // not separated at t = 0
        if (!(dist >= dist_eps)) break;     // already within distanceEpsilon: time of impact 0
        if (r_sqrt(dpx * dpx + dpy * dpy) + amax == cr(0)) { skip = true; break; }
      } else if (stage == 1) {
        if (!sep) {  // overshoot: back up and take the distance there
          l -= adv;
          stage = 2;
          continue;
        }
      } else {
        break;
      }
      if (!(dist > dist_eps && iter < 30)) break;
      // (n = (q2 - q1) / dist, drel = dp . n + amax, advance = dist / drel -- with one division instead of three)
      const cr num = dpx * (q2x - q1x) + dpy * (q2y - q1y) + amax * dist;
      if (num <= eps_e * dist) { ok = false; break; }
      adv = dist * dist / num;
      l += adv;
      if (l < cr(0) || l > t2) { ok = false; break; }
      if (l <= l0) break;
      l0 = l;
      iter++;
      stage = 1;
    }
#ifdef VSR_CCD_COUNT
    if (lane == 0) atomicAdd(&g_ccd_hist[iter >= 20 ? 3 : (iter >= 8 ? 2 : (iter >= 3 ? 1 : 0))], 1ull);  // (polygons by advances)
#endif
    if (skip || !ok) continue;
    if (l < t2) { t2 = l; found = true; best_d = dist; bp1x = q1x; bp1y = q1y; bp2x = q2x; bp2y = q2y; }
  }
  if (!found) return p1r;
  Pose3<cr> r;
  r.x = p0.x + (p1.x - p0.x) * t2;
  r.y = p0.y + (p1.y - p0.y) * t2;
  r.th = p0.th + (p1.th - p0.th) * t2;
  {
    cr nx = bp2x - bp1x, ny = bp2y - bp1y;
    const cr nl = r_sqrt(nx * nx + ny * ny);
    if (nl > cr(0)) { nx /= nl; ny /= nl; }
    cr C = best_d - lin_tol;
    C = C < -max_corr ? -max_corr : (C > cr(0) ? cr(0) : C);
    const cr r1x = bp1x - r.x, r1y = bp1y - r.y;
    const cr rn1 = r1x * ny - r1y * nx;
    const cr K = im + ii * rn1 * rn1;
    const cr imp = K > eps_e ? -C / K : cr(0);
    const cr Jx = nx * imp, Jy = ny * imp;
    r.x += Jx * im;
    r.y += Jy * im;
    r.th += ii * (r1x * Jy - r1y * Jx);
  }
  Pose3<real> out = {(real)r.x, (real)r.y, (real)r.th};
  return out;
}

// the scalars the contact routines need, passed by value (registers): taking the kernel's
// Params by reference would force a local-memory copy of the whole struct
template <typename real>
struct CPar {
  real im, ii, mu, rest_vel, rest_e, lin_tol, max_corr, baumgarte;
  real mu_s, rest_e_s;  // mass against mass (CoefficientMixer: sqrt(f f), max(e, e))
};

// ---------------------------------------------------------------- constraints: the solver's view
// One record type serves both kinds of contact constraint (SCon): a mass against a ground segment
// (tb < 0: body 2 is static, invM2 = invI2 = 0 -- exactly how the oracle solves it, with its
// STATIC_BODY) and a mass against another mass of the robot.  The mass table `mt` is the block's
// shared-memory exchange: mt[(k * 4 + c) * mts + t] = component c of mass k of thread t;
// components are (x, y, cos, sin) during detection, (vx, vy, w) in the velocity phases and
// (x, y, theta) in the position phase.
#define VSR_MTI(k, c, t) (((k) * 4 + (c)) * mts + (t))

// what the velocity phases need of one constraint, in registers
template <typename real>
struct ConVel {
  int n, block, ka, kb, tb;
  real nx, ny;
  real r1x[2], r1y[2], r2x[2], r2y[2], mN[2], mT[2], vb[2], jn[2], jt[2];
  real K00, K01, K11, iK00, iK01, iK11;
};
// ... and the position phase
template <typename real>
struct ConPos {
  int n, ka, kb, tb;
  real nx, ny;
  real l1x[2], l1y[2], l2x[2], l2y[2], depth[2];
};
#define VSR_POOL_FIELDS 27
template <typename real>
__device__ __forceinline__ int con_hdr(const SCon<real>& c) {
  return c.n | (c.block << 2) | (c.ka << 4) | (c.kb << 6) | ((c.tb + 1) << 8);
}
#define VSR_HDR_UNPACK(d, h) \
  d.n = (h) & 3; d.ka = ((h) >> 4) & 3; d.kb = ((h) >> 6) & 3; d.tb = ((h) >> 8) - 1;

template <typename real>
__device__ __noinline__ ConVel<real> con_vel_from_record(const SCon<real>& c) {
  ConVel<real> d;
  d.n = c.n; d.block = c.block; d.ka = c.ka; d.kb = c.kb; d.tb = c.tb;
  d.nx = c.nx; d.ny = c.ny;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    d.r1x[k] = c.p[k].r1x; d.r1y[k] = c.p[k].r1y; d.r2x[k] = c.p[k].r2x; d.r2y[k] = c.p[k].r2y;
    d.mN[k] = c.p[k].massN; d.mT[k] = c.p[k].massT; d.vb[k] = c.p[k].vb;
    d.jn[k] = c.p[k].jn; d.jt[k] = c.p[k].jt;
  }
  d.K00 = c.K00; d.K01 = c.K01; d.K11 = c.K11; d.iK00 = c.iK00; d.iK01 = c.iK01; d.iK11 = c.iK11;
  return d;
}
// pool[field * cap + slot]: the block's constraints of this tick, cached in shared memory
template <typename real>
__device__ __forceinline__ void con_vel_to_pool(const ConVel<real>& d, real* pool, int cap, int s) {
  reinterpret_cast<int*>(pool)[s * (int)(sizeof(real) / 4)] = d.n | (d.block << 2) | (d.ka << 4) | (d.kb << 6) | ((d.tb + 1) << 8);
  pool[1 * cap + s] = d.nx; pool[2 * cap + s] = d.ny;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int f = 3 + 9 * k;
    pool[(f + 0) * cap + s] = d.r1x[k]; pool[(f + 1) * cap + s] = d.r1y[k]; pool[(f + 2) * cap + s] = d.r2x[k];
    pool[(f + 3) * cap + s] = d.r2y[k]; pool[(f + 4) * cap + s] = d.mN[k]; pool[(f + 5) * cap + s] = d.mT[k];
    pool[(f + 6) * cap + s] = d.vb[k]; pool[(f + 7) * cap + s] = d.jn[k]; pool[(f + 8) * cap + s] = d.jt[k];
  }
  pool[21 * cap + s] = d.K00; pool[22 * cap + s] = d.K01; pool[23 * cap + s] = d.K11;
  pool[24 * cap + s] = d.iK00; pool[25 * cap + s] = d.iK01; pool[26 * cap + s] = d.iK11;
}
template <typename real>
__device__ __forceinline__ ConVel<real> con_vel_from_pool(const real* pool, int cap, int s) {
  ConVel<real> d;
  const int h = reinterpret_cast<const int*>(pool)[s * (int)(sizeof(real) / 4)];
  VSR_HDR_UNPACK(d, h)
  d.block = (h >> 2) & 1;
  d.nx = pool[1 * cap + s]; d.ny = pool[2 * cap + s];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int f = 3 + 9 * k;
    d.r1x[k] = pool[(f + 0) * cap + s]; d.r1y[k] = pool[(f + 1) * cap + s]; d.r2x[k] = pool[(f + 2) * cap + s];
    d.r2y[k] = pool[(f + 3) * cap + s]; d.mN[k] = pool[(f + 4) * cap + s]; d.mT[k] = pool[(f + 5) * cap + s];
    d.vb[k] = pool[(f + 6) * cap + s]; d.jn[k] = pool[(f + 7) * cap + s]; d.jt[k] = pool[(f + 8) * cap + s];
  }
  d.K00 = pool[21 * cap + s]; d.K01 = pool[22 * cap + s]; d.K11 = pool[23 * cap + s];
  d.iK00 = pool[24 * cap + s]; d.iK01 = pool[25 * cap + s]; d.iK11 = pool[26 * cap + s];
  return d;
}
template <typename real>
__device__ __noinline__ ConPos<real> con_pos_from_record(const SCon<real>& c) {
  ConPos<real> d;
  d.n = c.n; d.ka = c.ka; d.kb = c.kb; d.tb = c.tb;
  d.nx = c.nx; d.ny = c.ny;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    d.l1x[k] = c.p[k].l1x; d.l1y[k] = c.p[k].l1y; d.l2x[k] = c.p[k].l2x; d.l2y[k] = c.p[k].l2y;
    d.depth[k] = c.p[k].depth;
  }
  return d;
}
template <typename real>
__device__ __forceinline__ void con_pos_to_pool(const ConPos<real>& d, real* pool, int cap, int s) {
  reinterpret_cast<int*>(pool)[s * (int)(sizeof(real) / 4)] = d.n | (d.ka << 4) | (d.kb << 6) | ((d.tb + 1) << 8);
  pool[1 * cap + s] = d.nx; pool[2 * cap + s] = d.ny;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int f = 3 + 5 * k;
    pool[(f + 0) * cap + s] = d.l1x[k]; pool[(f + 1) * cap + s] = d.l1y[k]; pool[(f + 2) * cap + s] = d.l2x[k];
    pool[(f + 3) * cap + s] = d.l2y[k]; pool[(f + 4) * cap + s] = d.depth[k];
  }
}
template <typename real>
__device__ __forceinline__ ConPos<real> con_pos_from_pool(const real* pool, int cap, int s) {
  ConPos<real> d;
  const int h = reinterpret_cast<const int*>(pool)[s * (int)(sizeof(real) / 4)];
  VSR_HDR_UNPACK(d, h)
  d.nx = pool[1 * cap + s]; d.ny = pool[2 * cap + s];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int f = 3 + 5 * k;
    d.l1x[k] = pool[(f + 0) * cap + s]; d.l1y[k] = pool[(f + 1) * cap + s]; d.l2x[k] = pool[(f + 2) * cap + s];
    d.l2y[k] = pool[(f + 3) * cap + s]; d.depth[k] = pool[(f + 4) * cap + s];
  }
  return d;
}

// the two bodies of a constraint (body 2 of a ground constraint: zeros, never written back)
template <typename real>
struct Body2 {
  real v1x, v1y, w1, v2x, v2y, w2;
};
template <typename real>
__device__ __forceinline__ Body2<real> body2_load(const real* mt, int mts, int ta, int ka, int tb, int kb) {
  Body2<real> B;
  B.v1x = mt[VSR_MTI(ka, 0, ta)]; B.v1y = mt[VSR_MTI(ka, 1, ta)]; B.w1 = mt[VSR_MTI(ka, 2, ta)];
  const bool gnd = tb < 0;
  const int t2 = gnd ? ta : tb;
  B.v2x = gnd ? real(0) : mt[VSR_MTI(kb, 0, t2)];
  B.v2y = gnd ? real(0) : mt[VSR_MTI(kb, 1, t2)];
  B.w2 = gnd ? real(0) : mt[VSR_MTI(kb, 2, t2)];
  return B;
}
template <typename real>
__device__ __forceinline__ void body2_store(real* mt, int mts, int ta, int ka, int tb, int kb, const Body2<real>& B) {
  mt[VSR_MTI(ka, 0, ta)] = B.v1x; mt[VSR_MTI(ka, 1, ta)] = B.v1y; mt[VSR_MTI(ka, 2, ta)] = B.w1;
  if (tb >= 0) {
    mt[VSR_MTI(kb, 0, tb)] = B.v2x; mt[VSR_MTI(kb, 1, tb)] = B.v2y; mt[VSR_MTI(kb, 2, tb)] = B.w2;
  }
}
template <typename real>
__device__ __forceinline__ void body2_apply(const real im, const real ii, const real im2, const real ii2, Body2<real>& B,
                                            real r1x, real r1y, real r2x, real r2y, real Px, real Py) {
  B.v1x -= im * Px; B.v1y -= im * Py; B.w1 -= ii * (r1x * Py - r1y * Px);
  B.v2x += im2 * Px; B.v2y += im2 * Py; B.w2 += ii2 * (r2x * Py - r2y * Px);
}

// SequentialImpulses.initialize (dyn4j), the velocity-dependent half: the restitution bias, from
// the velocities before any warm start (the position-dependent half is done at detection)
template <typename real>
__device__ __forceinline__ void con_bias(const CPar<real>& P, ConVel<real>& d, const Body2<real>& B) {
  const real e = d.tb < 0 ? P.rest_e : P.rest_e_s;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    real rvx = (B.v2x - B.w2 * d.r2y[k]) - (B.v1x - B.w1 * d.r1y[k]);
    real rvy = (B.v2y + B.w2 * d.r2x[k]) - (B.v1y + B.w1 * d.r1x[k]);
    real rvn = d.nx * rvx + d.ny * rvy;
    real vb = real(0);
    if (rvn < -P.rest_vel) vb += -e * rvn;
    d.vb[k] = vb;
  }
}
// warm start
template <typename real>
__device__ __forceinline__ void con_warm(const CPar<real>& P, const ConVel<real>& d, Body2<real>& B) {
  const real im2 = d.tb < 0 ? real(0) : P.im, ii2 = d.tb < 0 ? real(0) : P.ii;
  const real nx = d.nx, ny = d.ny, tx = ny, ty = -nx;
#pragma unroll
  for (int k = 0; k < 2; k++)
    if (k < d.n)
      body2_apply(P.im, P.ii, im2, ii2, B, d.r1x[k], d.r1y[k], d.r2x[k], d.r2y[k], nx * d.jn[k] + tx * d.jt[k],
                  ny * d.jn[k] + ty * d.jt[k]);
}
// SequentialImpulses.solveVelocityConstraints (dyn4j) for one constraint: friction of every point
// first, then the normal constraint(s) (block solver for two points)
template <typename real>
__device__ __forceinline__ void con_solve(const CPar<real>& P, ConVel<real>& d, Body2<real>& B) {
  const bool gnd = d.tb < 0;
  const real im = P.im, ii = P.ii, im2 = gnd ? real(0) : im, ii2 = gnd ? real(0) : ii;
  const real mu = gnd ? P.mu : P.mu_s;
  const real nx = d.nx, ny = d.ny, tx = ny, ty = -nx;
  const bool two = d.n == 2;
#define VSR_RELV(k, rvx, rvy)                                        \
  real rvx = (B.v2x - B.w2 * d.r2y[k]) - (B.v1x - B.w1 * d.r1y[k]);  \
  real rvy = (B.v2y + B.w2 * d.r2x[k]) - (B.v1y + B.w1 * d.r1x[k]);
#pragma unroll
  for (int k = 0; k < 2; k++) {
    if (k == 0 || two) {
      VSR_RELV(k, rvx, rvy)
      real tn = tx * rvx + ty * rvy;
      real jt = d.mT[k] * (-tn);
      real maxJt = mu * d.jn[k];
      real J0 = d.jt[k];
      d.jt[k] = r_max(-maxJt, r_min(J0 + jt, maxJt));
      jt = d.jt[k] - J0;
      body2_apply(im, ii, im2, ii2, B, d.r1x[k], d.r1y[k], d.r2x[k], d.r2y[k], tx * jt, ty * jt);
    }
  }
  if (!two) {
    // (a two-point constraint always carries its block matrix: an ill-conditioned one was cut to
    // one point at initialisation)
    VSR_RELV(0, rvx, rvy)
    real rvn = nx * rvx + ny * rvy;
    real j = -d.mN[0] * (rvn - d.vb[0]);
    real j0 = d.jn[0];
    d.jn[0] = r_max(j0 + j, real(0));
    j = d.jn[0] - j0;
    body2_apply(im, ii, im2, ii2, B, d.r1x[0], d.r1y[0], d.r2x[0], d.r2y[0], nx * j, ny * j);
  } else {
    // 2-point block solver (the Box2D / dyn4j LCP enumeration)
    real ax = d.jn[0], ay = d.jn[1];
    real rvn1, rvn2;
    {
      VSR_RELV(0, rvx, rvy)
      rvn1 = nx * rvx + ny * rvy;
    }
    {
      VSR_RELV(1, rvx, rvy)
      rvn2 = nx * rvx + ny * rvy;
    }
    real bx = rvn1 - d.vb[0], by = rvn2 - d.vb[1];
    bx -= d.K00 * ax + d.K01 * ay;
    by -= d.K01 * ax + d.K11 * ay;
    real x = -(d.iK00 * bx + d.iK01 * by), y = -(d.iK01 * bx + d.iK11 * by);
    bool ok = (x >= real(0) && y >= real(0));
    if (!ok) {
      x = -d.mN[0] * bx;
      y = real(0);
      real v2 = d.K01 * x + by;
      ok = (x >= real(0) && v2 >= real(0));
    }
    if (!ok) {
      x = real(0);
      y = -d.mN[1] * by;
      real v1 = d.K01 * y + bx;
      ok = (y >= real(0) && v1 >= real(0));
    }
    if (!ok) {
      x = real(0);
      y = real(0);
      ok = (bx >= real(0) && by >= real(0));
    }
    if (ok) {
      real dx = x - ax, dy = y - ay;
      body2_apply(im, ii, im2, ii2, B, d.r1x[0], d.r1y[0], d.r2x[0], d.r2y[0], nx * dx, ny * dx);
      body2_apply(im, ii, im2, ii2, B, d.r1x[1], d.r1y[1], d.r2x[1], d.r2y[1], nx * dy, ny * dy);
      d.jn[0] = x;
      d.jn[1] = y;
    }
  }
#undef VSR_RELV
}

// SequentialImpulses.solvePositionContraints (dyn4j) for one constraint; the table holds
// (x, y, theta).  Returns the minimum separation seen.
template <typename real>
__device__ __forceinline__ real con_position(const CPar<real>& P, const ConPos<real>& d, real* mt, int mts, int ta) {
  const bool gnd = d.tb < 0;
  const int ka = d.ka, kb = d.kb, tb = gnd ? ta : d.tb;
  real x1 = mt[VSR_MTI(ka, 0, ta)], y1 = mt[VSR_MTI(ka, 1, ta)], t1 = mt[VSR_MTI(ka, 2, ta)];
  real x2 = gnd ? real(0) : mt[VSR_MTI(kb, 0, tb)], y2 = gnd ? real(0) : mt[VSR_MTI(kb, 1, tb)],
       t2 = gnd ? real(0) : mt[VSR_MTI(kb, 2, tb)];
  const real im = P.im, ii = P.ii, im2 = gnd ? real(0) : im, ii2 = gnd ? real(0) : ii;
  const real nx = d.nx, ny = d.ny;
  real minSep = real(0);
#pragma unroll
  for (int k = 0; k < 2; k++) {
    if (k < d.n) {
      real s1, c1, s2 = real(0), c2 = real(1);
      r_sincos(t1, &s1, &c1);
      if (!gnd) r_sincos(t2, &s2, &c2);
      real r1x = c1 * d.l1x[k] - s1 * d.l1y[k], r1y = s1 * d.l1x[k] + c1 * d.l1y[k];
      real r2x = c2 * d.l2x[k] - s2 * d.l2y[k], r2y = s2 * d.l2x[k] + c2 * d.l2y[k];
      real p1x = x1 + r1x, p1y = y1 + r1y;
      real p2x = x2 + r2x, p2y = y2 + r2y;
      real sep = (p2x - p1x) * nx + (p2y - p1y) * ny - d.depth[k];
      minSep = sep < minSep ? sep : minSep;
      real cl = r_clamp<real>(sep + P.lin_tol, -P.max_corr, real(0));
      real cp = P.baumgarte * cl;
      real rn1 = r1x * ny - r1y * nx, rn2 = r2x * ny - r2y * nx;
      real K = im + im2 + ii * rn1 * rn1 + ii2 * rn2 * rn2;
      real jp = -cp / K;
      real Jx = nx * jp, Jy = ny * jp;
      x1 -= Jx * im; y1 -= Jy * im; t1 -= ii * (r1x * Jy - r1y * Jx);
      x2 += Jx * im2; y2 += Jy * im2; t2 += ii2 * (r2x * Jy - r2y * Jx);
    }
  }
  mt[VSR_MTI(ka, 0, ta)] = x1; mt[VSR_MTI(ka, 1, ta)] = y1; mt[VSR_MTI(ka, 2, ta)] = t1;
  if (!gnd) {
    mt[VSR_MTI(kb, 0, tb)] = x2; mt[VSR_MTI(kb, 1, tb)] = y2; mt[VSR_MTI(kb, 2, tb)] = t2;
  }
  return minSep;
}

// Narrowphase for one candidate pair + ContactConstraint.update (warm start by point id) + the
// position-dependent half of SequentialImpulses.initialize (the bodies do not move between this
// detection and the next step's initialisation).  Returns 1 when *out holds a constraint.
template <typename real>
__device__ __noinline__ int self_pair(const CPar<real> P, real half, real ax, real ay, real ac, real as, real bx,
                                      real by, real bc, real bs, int ka, int tb, int kb, const SCon<real>* old,
                                      unsigned old_mask, SCon<real>* out) {
  // the record is built in place (the owner's list in the arena)
  SCon<real>& c = *out;
  const int n_c = collide_mass_segment(ax, ay, ac, as, half, bx, by, bc, bs, real(0), 1, c);
  if (!n_c) return 0;
  c.ka = ka; c.tb = tb; c.kb = kb; c.block = 0; c.seg = -1;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    if (k >= n_c) { c.p[k].jn = real(0); c.p[k].jt = real(0); }
    c.p[k].vb = real(0);
  }
  // ContactConstraint.update: the first slot of the previous list that held this pair; its points' impulses are
  // carried over by id.  The keys of the slots in use are loaded together (independent, predicated loads), then
  // the matching record in one batch: two L2 round trips instead of one or more per slot.
  {
    int mi = -1;
#pragma unroll
    for (int i = VSR_MAX_CAND - 1; i >= 0; i--) {
      int oka = -1, otb = -2, okb = -1;
      if ((old_mask >> i) & 1u) { oka = old[i].ka; otb = old[i].tb; okb = old[i].kb; }
      if (oka == ka && otb == tb && okb == kb) mi = i;  // descending: the first match wins
    }
    if (mi >= 0) {
      const SCon<real>& o = old[mi];
      const int on = o.n, oid[2] = {o.p[0].id, o.p[1].id}, idn[2] = {c.p[0].id, c.p[1].id};
      const real ojn[2] = {o.p[0].jn, o.p[1].jn}, ojt[2] = {o.p[0].jt, o.p[1].jt};
#pragma unroll
      for (int k = 0; k < 2; k++) {
        real jn = real(0), jt = real(0);
        bool hit = false;
#pragma unroll
        for (int l = 0; l < 2; l++)
          if (k < n_c && l < on && oid[l] == idn[k]) { jn = ojn[l]; jt = ojt[l]; hit = true; }
        if (hit) { c.p[k].jn = jn; c.p[k].jt = jt; }
      }
    }
  }
  const real im = P.im, ii = P.ii;
  const real nx = c.nx, ny = c.ny, tx = ny, ty = -nx;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    SPoint<real>& p = c.p[k];
    if (k >= n_c) { p.px = ax; p.py = ay; }  // unused slot: finite numbers
    p.r1x = p.px - ax; p.r1y = p.py - ay;
    p.r2x = p.px - bx; p.r2y = p.py - by;
    p.l1x = ac * p.r1x + as * p.r1y; p.l1y = -as * p.r1x + ac * p.r1y;
    p.l2x = bc * p.r2x + bs * p.r2y; p.l2y = -bs * p.r2x + bc * p.r2y;
    real rn1 = p.r1x * ny - p.r1y * nx, rn2 = p.r2x * ny - p.r2y * nx;
    p.massN = real(1) / (im + im + ii * rn1 * rn1 + ii * rn2 * rn2);
    real rt1 = p.r1x * ty - p.r1y * tx, rt2 = p.r2x * ty - p.r2y * tx;
    p.massT = real(1) / (im + im + ii * rt1 * rt1 + ii * rt2 * rt2);
  }
  c.K00 = c.K01 = c.K11 = c.iK00 = c.iK01 = c.iK11 = real(0);
  if (n_c == 2) {
    real rn1A = c.p[0].r1x * ny - c.p[0].r1y * nx, rn1B = c.p[0].r2x * ny - c.p[0].r2y * nx;
    real rn2A = c.p[1].r1x * ny - c.p[1].r1y * nx, rn2B = c.p[1].r2x * ny - c.p[1].r2y * nx;
    real mm = im + im;
    c.K00 = mm + ii * rn1A * rn1A + ii * rn1B * rn1B;
    c.K01 = mm + ii * rn1A * rn2A + ii * rn1B * rn2B;
    c.K11 = mm + ii * rn2A * rn2A + ii * rn2B * rn2B;
    real det = c.K00 * c.K11 - c.K01 * c.K01;
    if (c.K00 * c.K00 < real(1000) * det) {
      c.block = 1;
      real id = real(1) / det;
      c.iK00 = id * c.K11;
      c.iK01 = -id * c.K01;
      c.iK11 = id * c.K00;
    } else {
      c.n = 1;
    }
  }
  return 1;
}

// ---------------------------------------------------------------- the block's constraint list
// An item: (owner thread << 8) | kind-specific bits | kind.  Ground (kind 0): (mass * 2 + buffer)
// << 5 | index << 1 -> the owner's CBody.  Intra-robot (kind 1): list << 5 | index << 1 -> the
// owner's SList.
#define VSR_ITEM_OWNER(code) ((code) >> 8)
template <typename real>
__device__ __forceinline__ SCon<real>& item_record(CBody<real>* blk_cb, SList<real>* blk_sl, int code) {
  const int ow = VSR_ITEM_OWNER(code), sub = (code >> 5) & 7, idx = (code >> 1) & 15;
  SCon<real>* g = blk_cb[(size_t)ow * 8 + sub].c + idx;
  SCon<real>* s = blk_sl[(size_t)ow * 2 + (sub & 1)].c + idx;
  return *((code & 1) ? s : g);
}
enum : int { ROUND_WARM = 0, ROUND_VELOCITY = 1, ROUND_POSITION = 2 };
// One round per dependency level, every thread of the scope (the warp; the block of a big robot)
// takes items: the items of a level
// touch disjoint masses, so solving them at once equals solving the whole list in its order (per
// robot: the ground constraints mass by mass, then the intra-robot constraints in the robot's
// enumeration).  Items below `cap` live in the shared-memory pool, the rest in their records.
// `aux` (position rounds): [0][thread] robot already solved, [1][thread] a constraint of this
// thread is still violated.
template <int MODE, bool BLK, typename real>
__device__ __forceinline__ void contact_rounds(const CPar<real> cpar, real* mt, int mts, const int* lhd, const int* items,
                                               int n_lv, int st, int sdim, int bdim, CBody<real>* blk_cb,
                                               SList<real>* blk_sl, real* aux, real* pool, int cap, int bar_id) {
  for (int l = 0, lo = 0; l < n_lv; l++) {
    const int hi = lhd[2 + l];
    for (int i = lo + st; i < hi; i += sdim) {
      const int code = items[i];
      const int ow = VSR_ITEM_OWNER(code);
      const bool pooled = i < cap;
      if (MODE == ROUND_POSITION) {
        if (aux[ow] != real(0)) continue;
        const ConPos<real> d = pooled ? con_pos_from_pool(pool, cap, i) : con_pos_from_record(item_record(blk_cb, blk_sl, code));
        const real ms = con_position(cpar, d, mt, mts, ow);
        if (ms < real(-3) * cpar.lin_tol) aux[bdim + ow] = real(1);
      } else {
        ConVel<real> d = pooled ? con_vel_from_pool(pool, cap, i) : con_vel_from_record(item_record(blk_cb, blk_sl, code));
        Body2<real> B = body2_load(mt, mts, ow, d.ka, d.tb, d.kb);
        if (MODE == ROUND_WARM) {
          con_warm(cpar, d, B);
        } else {
          con_solve(cpar, d, B);
          if (pooled) {
            pool[10 * cap + i] = d.jn[0]; pool[11 * cap + i] = d.jt[0];
            pool[19 * cap + i] = d.jn[1]; pool[20 * cap + i] = d.jt[1];
          } else {
            SCon<real>& c = item_record(blk_cb, blk_sl, code);
            c.p[0].jn = d.jn[0]; c.p[0].jt = d.jt[0];
            c.p[1].jn = d.jn[1]; c.p[1].jt = d.jt[1];
          }
        }
        body2_store(mt, mts, ow, d.ka, d.tb, d.kb, B);
      }
    }
    lo = hi;
    if (BLK) robot_sync(bar_id, sdim); else __syncwarp();
  }
}

// ---------------------------------------------------------------- springs
// The 18 DistanceJoints of a voxel, indices as in `allSpringJoints` (Voxel.java:303-470).
// Listed in the canonical SOLVE order: inside a voxel the sweep alternates the perfect matchings
// of K4 ({01,23}, {12,30}, {02,13}) so that consecutive solves touch disjoint masses (two
// independent dependency chains); `nxt` is the index of the spring solved next (prefetch).
// X(index, nxt, body1, body2, anchor1 sign x,y, anchor2 sign x,y, range class)
#define VSR_SPRINGS(X)                  \
  X(0, 2, 0, 1, +1, -1, -1, -1, 0)      \
  X(2, 1, 2, 3, -1, +1, +1, +1, 0)      \
  X(1, 3, 1, 2, -1, -1, -1, +1, 0)      \
  X(3, 4, 3, 0, +1, +1, +1, -1, 0)      \
  X(4, 6, 0, 1, +1, +1, -1, +1, 0)      \
  X(6, 5, 2, 3, -1, -1, +1, -1, 0)      \
  X(5, 7, 1, 2, +1, -1, +1, +1, 0)      \
  X(7, 8, 3, 0, -1, +1, -1, -1, 0)      \
  X(8, 12, 0, 1, +1, +1, -1, -1, 1)     \
  X(12, 9, 2, 3, -1, +1, +1, -1, 1)     \
  X(9, 13, 0, 1, +1, -1, -1, +1, 1)     \
  X(13, 10, 2, 3, -1, -1, +1, +1, 1)    \
  X(10, 14, 1, 2, +1, -1, -1, +1, 1)    \
  X(14, 11, 3, 0, -1, +1, +1, -1, 1)    \
  X(11, 15, 1, 2, -1, -1, +1, +1, 1)    \
  X(15, 16, 3, 0, +1, +1, -1, -1, 1)    \
  X(16, 17, 0, 2, 0, 0, 0, 0, 2)        \
  X(17, 0, 1, 3, 0, 0, 0, 0, 2)

template <typename real>
__device__ __forceinline__ real activation(int a, real x) {
  switch (a) {
    case 0: return x < real(0) ? real(0) : x;
    case 1: return real(1) / (real(1) + r_exp(-x));
    case 2: return r_sin(x);
    case 3: return r_tanh(x);
    case 4: return x > real(0) ? real(1) : (x < real(0) ? real(-1) : x);
    default: return x;
  }
}

// One soft DistanceJoint velocity solve (DistanceJoint.solveVelocityConstraints, dyn4j):
//   Jv = n.(v1 - v2) + cr1 w1 - cr2 w2 ;  j = -softMass (Jv + bias + gamma Lambda) ; apply.
// fp64 (parity build): scalar, in the oracle's operation order.
__device__ __forceinline__ void spring_solve(double& vx1, double& vy1, double& w1, double& vx2, double& vy2, double& w2,
                                             double& imp, const double4 q, const double2 p, double g, double im, double ii) {
  double Jv = q.x * (vx1 - vx2) + q.y * (vy1 - vy2) + q.z * w1 - q.w * w2;
  double j = -p.x * (Jv + p.y + g * imp);
  imp += j;
  double jx = q.x * j, jy = q.y * j;
  vx1 += jx * im; vy1 += jy * im; w1 += ii * q.z * j;
  vx2 -= jx * im; vy2 -= jy * im; w2 -= ii * q.w * j;
}
// fp32 (product path): Blackwell packed-FP32 instructions (FFMA2 / FMUL2 on aligned register
// pairs, sm_100+) for the x/y halves: 17 instructions instead of 22.
__device__ __forceinline__ void spring_solve(float& vx1, float& vy1, float& w1, float& vx2, float& vy2, float& w2,
                                             float& imp, const float4 q, const float2 p, float g, float im, float ii) {
  const float2 n = make_float2(q.x, q.y);
  float2 v1 = make_float2(vx1, vy1), v2 = make_float2(vx2, vy2);
  const float2 dv = __ffma2_rn(v2, make_float2(-1.0f, -1.0f), v1);
  const float2 t = __fmul2_rn(n, dv);
  float Jv = t.x + t.y;
  Jv = fmaf(q.z, w1, Jv);
  Jv = fmaf(-q.w, w2, Jv);
  const float j = -p.x * (Jv + fmaf(g, imp, p.y));
  imp += j;
  const float jm = j * im;
  v1 = __ffma2_rn(n, make_float2(jm, jm), v1);
  v2 = __ffma2_rn(n, make_float2(-jm, -jm), v2);
  const float ji = j * ii;
  w1 = fmaf(q.z, ji, w1);
  w2 = fmaf(-q.w, ji, w2);
  vx1 = v1.x; vy1 = v1.y; vx2 = v2.x; vy2 = v2.y;
}

// WeldJoint rigid 3x3 solve with localAnchor1 = 0 (Robot.java:66-71 puts the anchor at
// body1's centre).  K = [[2im + ii r2y^2, -ii r2x r2y, -ii r2y], [., 2im + ii r2x^2, ii r2x],
// [., ., 2ii]]; K * lambda = -C solved in closed form (Schur complement on the angular row,
// then Sherman-Morrison with p = (-r2y, r2x), |p| = weld_len): algebraically the
// Matrix33.solve33 answer of dyn4j.  Used for velocities (C = Cdot) and positions (C = C).
// ONE function for both lanes that own the two masses, so both compute identical bits.
template <typename real>
__device__ __forceinline__ void weld_solve33(real cx, real cy, real cz, real r2x, real r2y, real inertia,
                                             real weld_alpha, real inv_two_im, real& lx, real& ly, real& lz) {
  real px = -r2y, py = r2x;
  real bx = -cx + real(0.5) * px * cz, by = -cy + real(0.5) * py * cz;
  real pb = px * bx + py * by;
  lx = (bx - weld_alpha * px * pb) * inv_two_im;
  ly = (by - weld_alpha * py * pb) * inv_two_im;
  lz = real(0.5) * (-cz * inertia - (px * lx + py * ly));
}

// ---------------------------------------------------------------- the episode kernel
#ifndef VSR_MAX_THREADS
#define VSR_MAX_THREADS 384
// reals of shared memory reserved for the header of the block's contact list
#define VSR_LHD_REALS 80
#define VSR_MAX_LEVELS 32
#define VSR_ITEMS_PER_THREAD 28
static_assert(4 * MAXC + VSR_MAX_CAND <= VSR_ITEMS_PER_THREAD, "contact list capacity");
#endif
#ifndef VSR_MIN_BLOCKS
#define VSR_MIN_BLOCKS 1
#endif
// MODE 1 (FAST) = the default voxel (all four scaffoldings, reduced-mass springs, no limits): no per-spring
// branch, constant gamma.  MODE 0 = generic (partial scaffoldings, effective-mass springs).  MODE 2 (LIM) = generic
// + everything that is off in the reference's default configuration, so that none of it weighs on the default kernel:
// the DistanceJoint limits of a finite areaRatioPassiveRange (Voxel.java:331-338: two more accumulated impulses per
// spring, a one-sided rigid row per limit after each spring row, a position correction of the violated limit; 256 /
// 128 thread columns per block: the per-tick constants take 36 bytes more per spring), BreakableVoxels, and the
// continuous-detection pass (vsr_settings.continuous_detection).  MODE 3 = MODE 1 + the continuous-detection pass (the
// reference's default voxel under dyn4j's default ContinuousDetectionMode).
// BLK = big robots (33..128 voxels): a robot takes P.wpr whole warps, a block holds P.rpb of them; the robots of a
// block meet at the block-wide phase barriers only (instruction-fetch locality), everything inside a phase is
// synchronised per robot with a named barrier; neighbour exchange through shared memory and
// block-wide votes instead of warp shuffles / ballots.
template <typename real, int CTRL, int MODE, bool BLK>
__global__ void __launch_bounds__(VSR_MAX_THREADS, VSR_MIN_BLOCKS) vsr_episode_kernel(const Params<real> P) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ int ccd_n;  // continuous detection: length of the block's candidate list (MODE 2 / 3 only)
  const ShapeDev& S = *P.shape;
  constexpr bool FAST = MODE == 1 || MODE == 3, LIM = MODE == 2, CCD = MODE >= 2;
  constexpr bool FULLSPR = FAST;
  const bool EFFMASS = FAST ? false : (P.spring_mode != 0);
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  typedef typename Real4T<real>::type R4;
  typedef typename Real2T<real>::type R2;
  // compile-time stride (= the largest block) so that every access has an immediate offset
  constexpr int SPR_STRIDE = (BLK || LIM) ? (sizeof(real) == 4 ? 256 : 128) : (sizeof(real) == 4 ? VSR_MAX_THREADS : VSR_MAX_THREADS / 2);
  // the robot of this thread inside its block (BLK): its index, its thread count, this thread's / warp's index in it
  const int wpr = BLK ? P.wpr : 1, rpb = BLK ? P.rpb : 1;
  const int rthreads = wpr << 5;
  const int rib = BLK ? (int)(threadIdx.x >> 5) / wpr : 0;
  const int rt = BLK ? (int)threadIdx.x - rib * rthreads : 0;
  const int wir = BLK ? (int)(threadIdx.x >> 5) - rib * wpr : 0;
  const int bar_id = rib + 1;
  const int tid = threadIdx.x;
  R4* spr = reinterpret_cast<R4*>(smem_raw);                                  // [18][SPR_STRIDE]
  R2* spr2 = reinterpret_cast<R2*>(spr + 18 * SPR_STRIDE);                     // [18][SPR_STRIDE]
  real* sgam = reinterpret_cast<real*>(spr2 + 18 * SPR_STRIDE);               // [18][SPR_STRIDE], effective-mass mode only
  R2* const slim = reinterpret_cast<R2*>(sgam + (EFFMASS ? 18 * SPR_STRIDE : 0));  // [18][SPR_STRIDE] (rigid mass, length), LIM only
  real* const sfree = sgam + (EFFMASS ? 18 * SPR_STRIDE : 0) + (LIM ? 2 * 18 * SPR_STRIDE : 0);  // end of the spring tables
  real* stage = sfree + (size_t)wib * P.smem_per_warp;
  // Contact exchange, block-wide.  Contacts are sparse and irregular (a few masses of a robot touch
  // the ground or each other), so the block pools them: every thread posts the state of its masses
  // in the mass table mt[16 fields][MTS threads] (see VSR_MTI), the block's constraints -- listed
  // once per tick in `items`, ordered by dependency level -- are solved one per thread, and the
  // owners read their masses back.
  constexpr int MTS = SPR_STRIDE;  // the largest block of this instantiation
  constexpr int mts = MTS;
  real* const mt = sfree + (size_t)(blockDim.x >> 5) * P.smem_per_warp;
  // list header: [0] ground items, [1] levels, [2..33] end offset of each level, [34..] scratch
  int* const lhd = reinterpret_cast<int*>(mt + 16 * MTS) + (BLK ? rib : wib) * VSR_LHD_REALS;
  // neighbour exchange of the one-block-per-robot variant: [2 buffers][20 fields][128 threads]
  real* const ex = mt + 16 * MTS + (size_t)(BLK ? rpb : (blockDim.x >> 5)) * VSR_LHD_REALS;
  // the tick's constraints as the solver rounds read them: [VSR_POOL_FIELDS][pool_cap]
  const int cap = P.pool_cap;  // per scope
  real* const pool = ex + (BLK ? 2 * 20 * SPR_STRIDE + rib * VSR_POOL_FIELDS * cap : wib * VSR_POOL_FIELDS * cap);
  const int st = BLK ? rt : lane, sdim = BLK ? rthreads : 32;  // thread index / size of the scope
  real* const stage_blk = sfree;
  // 4 reals per thread of block-wide scratch (the controller staging is idle outside control)
  real* const aux = stage_blk;
  const int bdim = blockDim.x;
  // (a thread lists at most 4 * MAXC ground + VSR_MAX_CAND intra-robot constraints)
  int* const items = P.lbuf + (size_t)blockIdx.x * VSR_ITEMS_PER_THREAD * MTS + (BLK ? rib * VSR_ITEMS_PER_THREAD * rthreads : wib * VSR_ITEMS_PER_THREAD * 32);
  const int mt_t = tid;
#define SSYNC() { if (BLK) robot_sync(bar_id, rthreads); else __syncwarp(); }
  int exsel = 0;
#define EXP(slot_, value_) if (BLK) ex[(exsel * 20 + (slot_)) * SPR_STRIDE + tid] = (value_);
#define EXSYNC() if (BLK) robot_sync(bar_id, rthreads);
#define EXG(slot_, value_, src_) (BLK ? ex[(exsel * 20 + (slot_)) * SPR_STRIDE + (src_)] : shfl(value_, src_))
#define EXFLIP() if (BLK) exsel ^= 1;
  const int nvox = S.nvox;
  const int G = BLK ? 1 : 32 / nvox;
  const int slot = BLK ? 0 : lane / nvox;
  const int v = BLK ? rt : lane - slot * nvox;
  const bool lane_ok = BLK ? (rt < nvox) : (slot < G);
  const int lane0 = BLK ? rib * rthreads : slot * nvox;  // first lane (block thread, BLK) of this lane's robot
  // lanes of this robot, for robot-wide votes
  const unsigned robot_mask = (!BLK && lane_ok) ? (((nvox == 32) ? 0xffffffffu : ((1u << nvox) - 1u)) << lane0) : 0u;
  // neighbour lanes (own lane when absent: the shuffled value is then ignored)
  int nl[4];
  bool has[4];
#pragma unroll
  for (int s = 0; s < 4; s++) {
    int nb = lane_ok ? (int)S.nbr[s][v] : -1;
    has[s] = nb >= 0;
    nl[s] = has[s] ? lane0 + nb : (BLK ? tid : lane);
  }
  real hm[4];
#pragma unroll
  for (int s = 0; s < 4; s++) hm[s] = has[s] ? real(1) : real(0);
  const real im = P.inv_m, ii = P.inv_i, dt = P.dt;
  const real two_im = real(2) * im;
  const real L2w = P.weld_len * P.weld_len;
  const real weld_alpha = (real(0.5) * ii) / (two_im + real(0.5) * ii * L2w);
  const real inv_two_im = real(1) / two_im;
  const real inertia = real(1) / ii;
  const real lin_f = r_clamp<real>(real(1) - dt * P.lin_damp, real(0), real(1));
  const real ang_f = r_clamp<real>(real(1) - dt * P.ang_damp, real(0), real(1));
  const real max_trans2 = P.max_trans * P.max_trans;
  const TPar<real> tpar = {P.terr_x, P.terr_y, P.terr_n, P.half, P.base_y};
  const CPar<real> cpar = {P.inv_m, P.inv_i, P.mu, P.rest_vel, P.rest_e, P.lin_tol, P.max_corr, P.baumgarte,
                           P.mu_self, P.rest_e_self};
  // reduced-mass springs: k, d and gamma are constants (same expression as the per-spring one)
  const real gam_c = real(1) / (dt * (real(2) * P.red_mass * P.spring_zeta * P.spring_w + dt * (P.red_mass * P.spring_w * P.spring_w)));
  const int sync_mode = P.sync_mode;

  const long long n_groups = (P.n_robots + G - 1) / G;
  // Every warp of a block runs the same number of passes (idle ones on a dummy robot) so that
  // the block-wide barriers below are legal.  The barriers carry no data: they keep the warps of
  // an SM inside the same stretch of the (large, fully unrolled) instruction stream, which turns
  // instruction-cache misses of one warp into hits for the others.
  const long long warps_total = BLK ? (long long)gridDim.x * rpb : (long long)gridDim.x * wpb;
  const long long n_pass = (n_groups + warps_total - 1) / warps_total;
  for (long long pass = 0; pass < n_pass; pass++) {
    const long long group = pass * warps_total + (BLK ? (long long)blockIdx.x * rpb + rib : (long long)blockIdx.x * wpb + wib);
    const long long rb = group * G + slot;  // bucket index of this lane's robot
    const bool active = lane_ok && rb < P.n_robots;
    const long long gl = rb * nvox + v;     // global voxel lane (ring addressing)
    const long long n_gl = P.n_robots * (long long)nvox;
    const real* par = P.params + (active ? rb : 0) * (long long)S.n_params;

    // ---- Voxel.assemble / Robot.assemble / placement: masses at rest
    real x[4], y[4], th[4], vx[4], vy[4], w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      x[k] = active ? (real)S.init_x[v * 4 + k] + (P.place_off ? P.place_off[2 * rb] : real(0)) : real(0);
      y[k] = active ? (real)S.init_y[v * 4 + k] + (P.place_off ? P.place_off[2 * rb + 1] : real(0)) : real(k);
      th[k] = vx[k] = vy[k] = w[k] = real(0);
    }
    // cos/sin of the 4 rotation angles, kept current across the tick (recomputed only where an
    // angle changed and is about to be used)
    real cs_[4] = {real(1), real(1), real(1), real(1)}, sn_[4] = {real(0), real(0), real(0), real(0)};
    real simp[18];
#pragma unroll
    for (int s = 0; s < 18; s++) simp[s] = real(0);
    real rest[3] = {P.rng_rest[0], P.rng_rest[1], P.rng_rest[2]};  // Voxel.reset: applyForce(0)
    // accumulated impulses of the lower / upper limit rows (LIM)
    real slo[LIM ? 18 : 1], shi[LIM ? 18 : 1];
#pragma unroll
    for (int s = 0; s < (LIM ? 18 : 1); s++) slo[s] = shi[s] = real(0);
    const real inv_dt = real(1) / dt;
    // BreakableVoxel (MODE 2 only): generator, malfunction state (2 bits per component: ACTUATOR, SENSORS, STRUCTURE;
    // bit 6 = the springs are rigid, bit 7 = a readings snapshot exists), trigger counters, bookkeeping of act()
    const bool brk = LIM && S.has_breakable && active && S.brk_enabled[v];
    unsigned long long bseed = brk ? (((unsigned long long)S.brk_seed[v] ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1)) : 0ull;
    unsigned bstate = 0;
    real bcnt[3] = {real(0), real(0), real(0)};
    double b_last_t = 0.0, b_last_break = 0.0;
    real b_last_ce = real(0), b_last_are = real(0);
    real snap[LIM ? MAX_READ : 1], rd_lo[LIM ? MAX_READ : 1], rd_hi[LIM ? MAX_READ : 1];
#pragma unroll 1
    for (int q = 0; q < (LIM ? MAX_READ : 1); q++) snap[q] = rd_lo[q] = rd_hi[q] = real(0);
    // weld impulses: [side L,R,D,U][pair][x,y,z]
    real wix[4][2], wiy[4][2], wiz[4][2];
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
      for (int k = 0; k < 2; k++) wix[s][k] = wiy[s][k] = wiz[s][k] = real(0);
    // contact state of this thread's 4 masses: a dense, double-buffered arena in global memory
    // (thread-local memory is lane-interleaved: one lane's 300-byte struct would touch 75 lines)
    CBody<real>* const cbase = reinterpret_cast<CBody<real>*>(P.cbuf) + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    unsigned flip = 0;
#define CBODY(k) cbase[2 * (k) + ((flip >> (k)) & 1u)]
    int hint[4];
#pragma unroll
    for (int k = 0; k < 4; k++) hint[k] = 0;
    unsigned cmask = 0;  // masses with ground constraints
    // intra-robot contacts owned by this thread: two lists (this step's / the previous step's) and,
    // packed in one word, the mask of the slots that hold a constraint (bits 0-15) and the current list (31)
    SList<real>* const sbase = reinterpret_cast<SList<real>*>(P.sbuf) + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    unsigned sst = 0;
#define S_MASK (sst & 0xffffu)
#define S_CUR (sbase + (sst >> 31))
#define VSR_MTK(k, c) (((k) * 4 + (c)) * MTS + mt_t)
#define VSR_MT_PUT(a0, a1, a2)                                                                     \
  _Pragma("unroll") for (int k = 0; k < 4; k++) {                                                  \
    mt[VSR_MTK(k, 0)] = a0[k]; mt[VSR_MTK(k, 1)] = a1[k]; mt[VSR_MTK(k, 2)] = a2[k];               \
  }
#define VSR_MT_GET(a0, a1, a2)                                                                     \
  _Pragma("unroll") for (int k = 0; k < 4; k++) {                                                  \
    a0[k] = mt[VSR_MTK(k, 0)]; a1[k] = mt[VSR_MTK(k, 1)]; a2[k] = mt[VSR_MTK(k, 2)];               \
  }
    CBody<real>* const blk_cb = reinterpret_cast<CBody<real>*>(P.cbuf) + (size_t)blockIdx.x * blockDim.x * 8;
    SList<real>* const blk_sl = reinterpret_cast<SList<real>*>(P.sbuf) + (size_t)blockIdx.x * blockDim.x * 2;
    int status = 0;
    real last_f = real(0), ar_energy = real(0), ctrl_energy = real(0);
    real readings[MAX_READ];
#pragma unroll
    for (int q = 0; q < MAX_READ; q++) readings[q] = real(0);
    double first_cx = 0, first_cy = 0, first_are = 0, first_ce = 0;
    // DistributedSensing: the signals this voxel emitted at the previous tick, [towards N, E, S, W][4]
    real lsig[CTRL == CTRL_MLP ? 16 : 1];
#pragma unroll
    for (int q = 0; q < (CTRL == CTRL_MLP ? 16 : 1); q++) lsig[q] = real(0);

    SSYNC()
    for (int q = st; q < 66; q += sdim) lhd[q] = 0;  // no contacts yet
    SSYNC()
    // tick -1 is detection only: a dyn4j World starts with updateRequired = true, i.e. the first step()
    // begins with detect() -- contacts that exist at the placement (a robot developed next to a wall,
    // DevoLocomotion.java) are solved from the first step on
    for (int tick = -1; tick < P.n_ticks; tick++) {
      // REQUIRED barrier (not only code locality): `aux` aliases the controller staging of the whole block
      // (stage_blk), so the position-pass / broadphase scratch of one warp lands in regions other warps use
      // for their MLP activations and first/last-tick sums; this barrier and the one before Robot.act
      // separate those phases block-wide.
      __syncthreads();
      if (CCD && tick < 0 && threadIdx.x == 0) ccd_n = 0;  // (its first use is a barrier away)
      if (tick >= 0) {
      // =========================== World.step(1) ===========================
      // (continuous detection: Body.transform0 = the poses before the islands are solved)
      if (CCD && P.ccd) {  // (compiled into the MODE 2 / 3 instantiations only: the default kernel stays as it is)
        // (the scratch column of this thread: recomputed where it is used -- kept across the velocity loop its four
        // registers pushed 17 values of that loop into local memory)
        const size_t ccd_t = (size_t)gridDim.x * blockDim.x, ccd_i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          P.ccd_buf[(k * 3 + 0) * ccd_t + ccd_i] = x[k];
          P.ccd_buf[(k * 3 + 1) * ccd_t + ccd_i] = y[k];
          P.ccd_buf[(k * 3 + 2) * ccd_t + ccd_i] = th[k];
          P.ccd_buf[(36 + k) * ccd_t + ccd_i] = P.half * (r_abs(cs_[k]) + r_abs(sn_[k]));  // half extent of its box
        }
      }
      // ---- integrate velocities (Island.solve / integrateVelocity, dyn4j)
#pragma unroll
      for (int k = 0; k < 4; k++) {
        vx[k] += dt * P.grav_x;
        vy[k] += dt * P.grav_y;
        vx[k] *= lin_f;
        vy[k] *= lin_f;
        w[k] *= ang_f;
      }
      // ---- contacts: SequentialImpulses.initialize.  The position-dependent half was done at
      //      detection; here the restitution bias of every constraint from the velocities before
      //      any warm start, then the warm starts.
      const int n_lv = lhd[1];  // levels of this scope's constraint list (0: no contacts); uniform
      if (n_lv > 0) {
        VSR_MT_PUT(vx, vy, w)
        SSYNC()
        for (int i_ = st, hi_ = lhd[2 + n_lv - 1]; i_ < hi_; i_ += sdim) {
          const int code = items[i_];
          const int ow = VSR_ITEM_OWNER(code);
          SCon<real>& c = item_record(blk_cb, blk_sl, code);
          ConVel<real> d = con_vel_from_record(c);
          const Body2<real> B = body2_load(mt, mts, ow, d.ka, d.tb, d.kb);
          con_bias(cpar, d, B);
          if (i_ < cap) {
            con_vel_to_pool(d, pool, cap, i_);
          } else {
            c.p[0].vb = d.vb[0];
            c.p[1].vb = d.vb[1];
          }
        }
        SSYNC()
        contact_rounds<ROUND_WARM, BLK>(cpar, mt, MTS, lhd, items, n_lv, st, sdim, bdim, blk_cb, blk_sl, aux, pool, cap, bar_id);
        VSR_MT_GET(vx, vy, w)
      }
      // ---- springs: DistanceJoint.initializeConstraints + warm start
      real ch[4], sh[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        ch[k] = cs_[k] * P.half;
        sh[k] = sn_[k] * P.half;
      }
      // per-spring constants of this tick go to shared memory: (nx, ny, cr1, cr2) as one 4-vector and
      // (softMass, bias) as one 2-vector per spring and thread ([spring][thread]: a warp reads
      // consecutive 16- / 8-byte words, conflict-free).  That frees ~110 registers.
      const real sw = P.spring_w;
#define VSR_SPRING_INIT(s, nxt, b1, b2, a1x, a1y, a2x, a2y, cls)                                        \
  if (FULLSPR || (P.spring_mask & (1u << s))) {                                                    \
    real r1x = real(a1x) * ch[b1] - real(a1y) * sh[b1], r1y = real(a1x) * sh[b1] + real(a1y) * ch[b1]; \
    real r2x = real(a2x) * ch[b2] - real(a2y) * sh[b2], r2y = real(a2x) * sh[b2] + real(a2y) * ch[b2]; \
    real dx = r1x + x[b1] - (r2x + x[b2]), dy = r1y + y[b1] - (r2y + y[b2]);                       \
    real len = r_sqrt(dx * dx + dy * dy);                                                          \
    real inv = len < P.lin_tol ? real(0) : real(1) / len;                                          \
    real nx = dx * inv, ny = dy * inv;                                                             \
    real cr1 = r1x * ny - r1y * nx, cr2 = r2x * ny - r2y * nx;                                     \
    real invMass = im + ii * cr1 * cr1;                                                            \
    invMass += im + ii * cr2 * cr2;                                                                \
    real m = EFFMASS ? real(1) / invMass : P.red_mass;                                       \
    real dc = real(2) * m * P.spring_zeta * sw;                                                    \
    real kk = m * sw * sw;                                                                         \
    real gamma = real(1) / (dt * (dc + dt * kk));                                                  \
    real bias = (len - rest[cls]) * dt * kk * gamma;                                               \
    if (LIM && (bstate & 64u)) { gamma = real(0); bias = real(0); } /* STRUCTURE FROZEN: a rigid joint */ \
    real soft = real(1) / (invMass + gamma);                                                       \
    {                                                                                              \
      R4 q_; q_.x = nx; q_.y = ny; q_.z = cr1; q_.w = cr2;                                         \
      spr[s * SPR_STRIDE + tid] = q_;                                                              \
      R2 p_; p_.x = soft; p_.y = bias;                                                             \
      spr2[s * SPR_STRIDE + tid] = p_;                                                             \
      if (EFFMASS) sgam[s * SPR_STRIDE + tid] = gamma;                                             \
      if (LIM) {                                                                                   \
        R2 l_; l_.x = real(1) / invMass; l_.y = len; /* the rigid row's effective mass, currentDistance */ \
        slim[s * SPR_STRIDE + tid] = l_;                                                           \
      }                                                                                            \
    }                                                                                              \
    /* warm start: spring + lower - upper along n (DistanceJoint.initializeConstraints) */         \
    const real wi_ = LIM ? simp[s] + slo[LIM ? s : 0] - shi[LIM ? s : 0] : simp[s];                \
    real jx = nx * wi_, jy = ny * wi_;                                                             \
    vx[b1] += jx * im; vy[b1] += jy * im; w[b1] += ii * cr1 * wi_;                                 \
    vx[b2] -= jx * im; vy[b2] -= jy * im; w[b2] -= ii * cr2 * wi_;                                 \
  }
      VSR_SPRINGS(VSR_SPRING_INIT)
#undef VSR_SPRING_INIT

      if (sync_mode & 16) __syncthreads();
      // ---- welds: WeldJoint.initializeConstraints + warm start.
      // Side s of this voxel: 0 left (own masses 0,3 are body1; partner masses 1,2 body2),
      // 1 right (own 1,2 are body2; partner 0,3 body1), 2 down (own 3,2 body1; partner 0,1
      // body2), 3 up (own 0,1 body2; partner 3,2 body1).  localAnchor1 = 0 (Robot.java:66-71),
      // localAnchor2 = (+len,0) for left/right pairs and (0,+len) for down/up pairs.
      real wr2x[4][2], wr2y[4][2];
      {
        // body2's rotation: partner's for sides 0 and 2, own for sides 1 and 3
        real pc, ps;
        // side 0: partner masses 1, 2
        EXP(0, cs_[0]) EXP(1, sn_[0]) EXP(2, cs_[1]) EXP(3, sn_[1]) EXP(4, cs_[2]) EXP(5, sn_[2])
        EXSYNC()
        pc = EXG(2, cs_[1], nl[0]); ps = EXG(3, sn_[1], nl[0]);
        wr2x[0][0] = pc * P.weld_len; wr2y[0][0] = ps * P.weld_len;
        pc = EXG(4, cs_[2], nl[0]); ps = EXG(5, sn_[2], nl[0]);
        wr2x[0][1] = pc * P.weld_len; wr2y[0][1] = ps * P.weld_len;
        // side 1: own masses 1, 2 are body2
        wr2x[1][0] = cs_[1] * P.weld_len; wr2y[1][0] = sn_[1] * P.weld_len;
        wr2x[1][1] = cs_[2] * P.weld_len; wr2y[1][1] = sn_[2] * P.weld_len;
        // side 2: partner masses 0, 1; anchor (0, len) -> R*(0,len) = (-s*len, c*len)
        pc = EXG(0, cs_[0], nl[2]); ps = EXG(1, sn_[0], nl[2]);
        wr2x[2][0] = -ps * P.weld_len; wr2y[2][0] = pc * P.weld_len;
        pc = EXG(2, cs_[1], nl[2]); ps = EXG(3, sn_[1], nl[2]);
        wr2x[2][1] = -ps * P.weld_len; wr2y[2][1] = pc * P.weld_len;
        // side 3: own masses 0, 1 are body2
        wr2x[3][0] = -sn_[0] * P.weld_len; wr2y[3][0] = cs_[0] * P.weld_len;
        wr2x[3][1] = -sn_[1] * P.weld_len; wr2y[3][1] = cs_[1] * P.weld_len;
        EXFLIP()
      }
      // warm start: body1 += (imp), body2 -= (imp), r1 = 0.  Each weld is owned (accumulated and
      // solved) by the lane of its body1 -- sides 0 (left) and 2 (down); the lane of body2 fetches the
      // accumulated impulse: its masses 1, 2 belong to the RIGHT neighbour's side-0 welds and its
      // masses 0, 1 to the UPPER neighbour's side-2 welds.  Impulses of absent welds stay exactly 0.
      {
        EXP(0, wix[0][0]) EXP(1, wiy[0][0]) EXP(2, wiz[0][0]) EXP(3, wix[0][1]) EXP(4, wiy[0][1]) EXP(5, wiz[0][1])
        EXP(6, wix[2][0]) EXP(7, wiy[2][0]) EXP(8, wiz[2][0]) EXP(9, wix[2][1]) EXP(10, wiy[2][1]) EXP(11, wiz[2][1])
        EXSYNC()
#define VSR_WELD_WARM_B1(side, k, own) \
  vx[own] += wix[side][k] * im; vy[own] += wiy[side][k] * im; w[own] += ii * wiz[side][k];
#define VSR_WELD_WARM_B2(side, k, own, src, slot0)                                              \
  {                                                                                              \
    const real ix_ = hm[side] * EXG(slot0, wix[src][k], nl[side]), iy_ = hm[side] * EXG(slot0 + 1, wiy[src][k], nl[side]), \
               iz_ = hm[side] * EXG(slot0 + 2, wiz[src][k], nl[side]);                          \
    vx[own] -= ix_ * im; vy[own] -= iy_ * im;                                                    \
    w[own] -= ii * (wr2x[side][k] * iy_ - wr2y[side][k] * ix_ + iz_);                            \
  }
        VSR_WELD_WARM_B1(0, 0, 0) VSR_WELD_WARM_B1(0, 1, 3)
        VSR_WELD_WARM_B1(2, 0, 3) VSR_WELD_WARM_B1(2, 1, 2)
        VSR_WELD_WARM_B2(1, 0, 1, 0, 0) VSR_WELD_WARM_B2(1, 1, 2, 0, 3)
        VSR_WELD_WARM_B2(3, 0, 0, 2, 6) VSR_WELD_WARM_B2(3, 1, 1, 2, 9)
        EXFLIP()
#undef VSR_WELD_WARM_B1
#undef VSR_WELD_WARM_B2
      }

      // ---- velocity iterations: all joints, then all contacts
#pragma unroll 1
      for (int it = 0; it < P.vel_iters; it++) {
        if (sync_mode & 2) __syncthreads();
#define VSR_SPRING_SOLVE(s, nxt, b1, b2, a1x, a1y, a2x, a2y, cls)                                          \
  /* software prefetch, one spring ahead (unconditional, so that an absent spring keeps the chain  \
     intact); the fences stop ptxas from hoisting all 36 loads to the top of the loop, which costs  \
     108 registers and spills them straight back to local memory */                                 \
  const R4 q##s = qn_;                                                                                \
  const R2 p##s = pn_;                                                                                \
  asm volatile("" ::: "memory");                                                                      \
  if (s != 17) {                                                                                      \
    qn_ = spr[(nxt) * SPR_STRIDE + tid];                                                              \
    pn_ = spr2[(nxt) * SPR_STRIDE + tid];                                                             \
  }                                                                                                   \
  asm volatile("" ::: "memory");                                                                      \
  if (FULLSPR || (P.spring_mask & (1u << s))) {                                                       \
    const R4 q_ = q##s;                                                                               \
    const R2 p_ = p##s;                                                                               \
    const real g_ = (LIM && (bstate & 64u)) ? real(0) : (EFFMASS ? sgam[s * SPR_STRIDE + tid] : gam_c); \
    spring_solve(vx[b1], vy[b1], w[b1], vx[b2], vy[b2], w[b2], simp[s], q_, p_, g_, im, ii);          \
    if (LIM) {                                                                                        \
      /* the limit rows of DistanceJoint.solveVelocityConstraints (dyn4j 4.2 = Box2D 2.4's joint; the     \
         oracle's spring_limits_solve): one-sided, accumulated impulse >= 0, bias = max(0, C) / dt */  \
      const R2 l_ = slim[s * SPR_STRIDE + tid];                                                       \
      if (P.lim_lo_on) {                                                                              \
        const real C_ = l_.y - P.lim_lo[cls];                                                         \
        const real bias_ = r_max(C_, real(0)) * inv_dt;                                               \
        const real Jv_ = q_.x * (vx[b1] - vx[b2]) + q_.y * (vy[b1] - vy[b2]) + q_.z * w[b1] - q_.w * w[b2]; \
        real j_ = -l_.x * (Jv_ + bias_);                                                              \
        const real nw_ = r_max(slo[LIM ? s : 0] + j_, real(0));                                       \
        j_ = nw_ - slo[LIM ? s : 0];                                                                  \
        slo[LIM ? s : 0] = nw_;                                                                       \
        const real jx_ = q_.x * j_, jy_ = q_.y * j_;                                                  \
        vx[b1] += jx_ * im; vy[b1] += jy_ * im; w[b1] += ii * q_.z * j_;                              \
        vx[b2] -= jx_ * im; vy[b2] -= jy_ * im; w[b2] -= ii * q_.w * j_;                              \
      }                                                                                               \
      if (P.lim_hi_on) {                                                                              \
        const real C_ = P.lim_hi[cls] - l_.y;                                                         \
        const real bias_ = r_max(C_, real(0)) * inv_dt;                                               \
        const real Jv_ = -(q_.x * (vx[b1] - vx[b2]) + q_.y * (vy[b1] - vy[b2]) + q_.z * w[b1] - q_.w * w[b2]); \
        real j_ = -l_.x * (Jv_ + bias_);                                                              \
        const real nw_ = r_max(shi[LIM ? s : 0] + j_, real(0));                                       \
        j_ = nw_ - shi[LIM ? s : 0];                                                                  \
        shi[LIM ? s : 0] = nw_;                                                                       \
        j_ = -j_;                                                                                     \
        const real jx_ = q_.x * j_, jy_ = q_.y * j_;                                                  \
        vx[b1] += jx_ * im; vy[b1] += jy_ * im; w[b1] += ii * q_.z * j_;                              \
        vx[b2] -= jx_ * im; vy[b2] -= jy_ * im; w[b2] -= ii * q_.w * j_;                              \
      }                                                                                               \
    }                                                                                                 \
  }
        R4 qn_ = spr[tid];
        R2 pn_ = spr2[tid];
        VSR_SPRINGS(VSR_SPRING_SOLVE)
#undef VSR_SPRING_SOLVE

        if (sync_mode & 8) __syncthreads();
        // WeldJoint.solveVelocityConstraints.  All partner values of a phase are gathered
        // BEFORE any lane applies an update, so the two lanes that own the two masses of a
        // weld compute the very same impulse.
#define VSR_PUBLISH_V()                                                                       \
  EXP(0, vx[0]) EXP(1, vy[0]) EXP(2, w[0]) EXP(3, vx[1]) EXP(4, vy[1]) EXP(5, w[1])                \
  EXP(6, vx[2]) EXP(7, vy[2]) EXP(8, w[2]) EXP(9, vx[3]) EXP(10, vy[3]) EXP(11, w[3]) EXSYNC()
#define VSR_GATHER_V(tag, part, side)                                                         \
  const real tag##vx = EXG(part * 3 + 0, vx[part], nl[side]), tag##vy = EXG(part * 3 + 1, vy[part], nl[side]), \
             tag##w = EXG(part * 3 + 2, w[part], nl[side]);
// body1's lane solves the weld (own mass `own` is body1, the gathered `tag` mass is body2) ...
#define VSR_WELD_VEL_B1(side, k, own, tag, L)                                                   \
  real L##x, L##y, L##z;                                                                        \
  {                                                                                             \
    real r2x = wr2x[side][k], r2y = wr2y[side][k];                                              \
    weld_solve33<real>(vx[own] - (tag##vx - tag##w * r2y), vy[own] - (tag##vy + tag##w * r2x), w[own] - tag##w, r2x, r2y, \
                       inertia, weld_alpha, inv_two_im, L##x, L##y, L##z);                      \
    /* branch-free: a voxel without a neighbour on this side scales the impulse by 0 */         \
    L##x *= hm[side]; L##y *= hm[side]; L##z *= hm[side];                                       \
    wix[side][k] += L##x; wiy[side][k] += L##y; wiz[side][k] += L##z;                           \
    vx[own] += L##x * im; vy[own] += L##y * im; w[own] += ii * L##z;                            \
  }
// ... and body2's lane applies the impulse it receives from that lane
#define VSR_WELD_VEL_B2(side, k, own, slot0, L)                                                 \
  {                                                                                             \
    const real bx_ = hm[side] * EXG(slot0, L##x, nl[side]), by_ = hm[side] * EXG(slot0 + 1, L##y, nl[side]), \
               bz_ = hm[side] * EXG(slot0 + 2, L##z, nl[side]);                                 \
    vx[own] -= bx_ * im; vy[own] -= by_ * im;                                                   \
    w[own] -= ii * (wr2x[side][k] * by_ - wr2y[side][k] * bx_ + bz_);                           \
  }
        {  // horizontal welds (left/right sides): each mass has at most one
          VSR_PUBLISH_V()
          VSR_GATHER_V(l1, 1, 0) VSR_GATHER_V(l2, 2, 0)
          VSR_WELD_VEL_B1(0, 0, 0, l1, la) VSR_WELD_VEL_B1(0, 1, 3, l2, lb)
          EXFLIP()
          EXP(0, lax) EXP(1, lay) EXP(2, laz) EXP(3, lbx) EXP(4, lby) EXP(5, lbz) EXSYNC()
          VSR_WELD_VEL_B2(1, 0, 1, 0, la) VSR_WELD_VEL_B2(1, 1, 2, 3, lb)
          EXFLIP()
        }
        {  // vertical welds (down/up sides)
          VSR_PUBLISH_V()
          VSR_GATHER_V(d0, 0, 2) VSR_GATHER_V(d1, 1, 2)
          VSR_WELD_VEL_B1(2, 0, 3, d0, da) VSR_WELD_VEL_B1(2, 1, 2, d1, db)
          EXFLIP()
          EXP(0, dax) EXP(1, day) EXP(2, daz) EXP(3, dbx) EXP(4, dby) EXP(5, dbz) EXSYNC()
          VSR_WELD_VEL_B2(3, 0, 0, 0, da) VSR_WELD_VEL_B2(3, 1, 1, 3, db)
          EXFLIP()
        }
#undef VSR_GATHER_V
#undef VSR_PUBLISH_V
#undef VSR_WELD_VEL_B1
#undef VSR_WELD_VEL_B2

        if (sync_mode & 8) __syncthreads();
        // contacts (SequentialImpulses.solveVelocityConstraints), pooled over the block
        if (n_lv > 0) {
          VSR_MT_PUT(vx, vy, w)
          SSYNC()
          contact_rounds<ROUND_VELOCITY, BLK>(cpar, mt, MTS, lhd, items, n_lv, st, sdim, bdim, blk_cb, blk_sl, aux, pool, cap, bar_id);
          VSR_MT_GET(vx, vy, w)
        }
      }

      if (n_lv > 0) {
        // pooled constraints: the accumulated impulses go back to their records (the next step's
        // warm start), the pool now caches what the position rounds read
        const int n_it = lhd[2 + n_lv - 1];
        for (int i_ = st; i_ < n_it && i_ < cap; i_ += sdim) {
          SCon<real>& c = item_record(blk_cb, blk_sl, items[i_]);
          c.p[0].jn = pool[10 * cap + i_]; c.p[0].jt = pool[11 * cap + i_];
          c.p[1].jn = pool[19 * cap + i_]; c.p[1].jt = pool[20 * cap + i_];
          con_pos_to_pool(con_pos_from_record(c), pool, cap, i_);
        }
        SSYNC()
      }
      if (sync_mode & 32) __syncthreads();
      // ---- integrate positions (integratePosition, dyn4j; translation/rotation caps)
#pragma unroll
      for (int k = 0; k < 4; k++) {
        real tx = vx[k] * dt, ty = vy[k] * dt;
        real m2 = tx * tx + ty * ty;
        if (m2 > max_trans2) {
          real ratio = P.max_trans / r_sqrt(m2);
          vx[k] *= ratio; vy[k] *= ratio; tx *= ratio; ty *= ratio;
        }
        real rot = w[k] * dt;
        if (r_abs(rot) > P.max_rot) {
          real ratio = P.max_rot / r_abs(rot);
          w[k] *= ratio; rot *= ratio;
        }
        x[k] += tx; y[k] += ty; th[k] += rot;
      }
      r_sincos(th[1], &sn_[1], &cs_[1]);
      r_sincos(th[2], &sn_[2], &cs_[2]);

      // ---- position iterations: contacts, then joints; stop when the whole robot
      //      (= one island) reports solved
      {
        bool done = BLK ? !(rb < P.n_robots) : !active;  // BLK: idle threads follow their block
        for (int it = 0; it < P.pos_iters; it++) {
          const bool wdone = BLK ? done : (bool)__all_sync(0xffffffffu, done);  // BLK: done is robot-uniform
          if (wdone) break;
          bool solved = true;
          if (n_lv > 0) {
            VSR_MT_PUT(x, y, th)
            aux[0 * bdim + tid] = done ? real(1) : real(0);
            aux[1 * bdim + tid] = real(0);
            SSYNC()
            contact_rounds<ROUND_POSITION, BLK>(cpar, mt, MTS, lhd, items, n_lv, st, sdim, bdim, blk_cb, blk_sl, aux, pool, cap, bar_id);
            if (!done) {
              VSR_MT_GET(x, y, th)
              r_sincos(th[1], &sn_[1], &cs_[1]);
              r_sincos(th[2], &sn_[2], &cs_[2]);
              solved = aux[1 * bdim + tid] == real(0);
            }
          }
          if (LIM && !wdone) {
            // DistanceJoint.solvePositionConstraints with limits (the oracle's spring_limits_position): the violated
            // limit is corrected rigidly with the effective mass of this step's initialisation; the joints come in
            // the island's order, the springs of a voxel before the welds
            r_sincos(th[0], &sn_[0], &cs_[0]);
            r_sincos(th[3], &sn_[3], &cs_[3]);
            // (idle threads of a block robot carry dummy masses: they must neither move nor vote "unsolved")
            const real pm_ = (done || !active) ? real(0) : real(1);
#define VSR_SPRING_POS(s, nxt, b1, b2, a1x, a1y, a2x, a2y, cls)                                            \
  if (P.spring_mask & (1u << s)) {                                                                    \
    const real h_ = P.half;                                                                           \
    const real r1x = (real(a1x) * cs_[b1] - real(a1y) * sn_[b1]) * h_, r1y = (real(a1x) * sn_[b1] + real(a1y) * cs_[b1]) * h_; \
    const real r2x = (real(a2x) * cs_[b2] - real(a2y) * sn_[b2]) * h_, r2y = (real(a2x) * sn_[b2] + real(a2y) * cs_[b2]) * h_; \
    real nx = r1x + x[b1] - (r2x + x[b2]), ny = r1y + y[b1] - (r2y + y[b2]);                          \
    const real len = r_sqrt(nx * nx + ny * ny);                                                       \
    const real inv = len > real(0) ? real(1) / len : real(0);                                         \
    nx *= inv; ny *= inv;                                                                             \
    const bool rg_ = (bstate & 64u) != 0u; /* rigid (STRUCTURE FROZEN): corrected to the rest distance on both sides */ \
    const bool lo_ = !rg_ && P.lim_lo_on && len < P.lim_lo[cls], hi_ = !rg_ && !lo_ && P.lim_hi_on && len > P.lim_hi[cls]; \
    if (rg_ || lo_ || hi_) {                                                                          \
      real C_ = rg_ ? len - rest[cls] : (lo_ ? len - P.lim_lo[cls] : len - P.lim_hi[cls]);            \
      C_ = r_clamp<real>(C_, -P.max_corr, P.max_corr);                                                \
      const real imp_ = -slim[s * SPR_STRIDE + tid].x * C_ * pm_;                                     \
      const real Jx = nx * imp_, Jy = ny * imp_;                                                      \
      x[b1] += Jx * im; y[b1] += Jy * im; th[b1] += ii * (r1x * Jy - r1y * Jx);                       \
      x[b2] -= Jx * im; y[b2] -= Jy * im; th[b2] -= ii * (r2x * Jy - r2y * Jx);                       \
      r_sincos(th[b1], &sn_[b1], &cs_[b1]);                                                           \
      r_sincos(th[b2], &sn_[b2], &cs_[b2]);                                                           \
      solved = solved && (!active || r_abs(C_) < P.lin_tol);                                          \
    }                                                                                                 \
  }
            VSR_SPRINGS(VSR_SPRING_POS)
#undef VSR_SPRING_POS
          }
          if (!wdone) {
          // WeldJoint.solvePositionConstraints: C = (p1 - p2, theta1 - theta2 - ref)
#define VSR_PUBLISH_P()                                                                       \
  EXP(0, x[0]) EXP(1, y[0]) EXP(2, th[0]) EXP(3, x[1]) EXP(4, y[1]) EXP(5, th[1])                  \
  EXP(6, x[2]) EXP(7, y[2]) EXP(8, th[2]) EXP(9, x[3]) EXP(10, y[3]) EXP(11, th[3])                \
  EXP(12, cs_[0]) EXP(13, sn_[0]) EXP(14, cs_[1]) EXP(15, sn_[1]) EXP(16, cs_[2]) EXP(17, sn_[2]) EXSYNC()
#define VSR_GATHER_P(tag, part, side)                                                         \
  const real tag##x = EXG(part * 3 + 0, x[part], nl[side]), tag##y = EXG(part * 3 + 1, y[part], nl[side]), \
             tag##t = EXG(part * 3 + 2, th[part], nl[side]);
#define VSR_GATHER_CS(tag, part, side) \
  const real tag##c = EXG(12 + part * 2, cs_[part], nl[side]), tag##s = EXG(13 + part * 2, sn_[part], nl[side]);
#define VSR_OWN_CS(tag, own) const real tag##c = cs_[own], tag##s = sn_[own];
#define VSR_WELD_POS_B1(side, own, tag, horiz, L)                                               \
  real L##x, L##y, L##z;                                                                        \
  {                                                                                             \
    const real s2 = tag##s, c2 = tag##c; /* body2's rotation */                                 \
    real r2x = horiz ? c2 * P.weld_len : -s2 * P.weld_len;                                      \
    real r2y = horiz ? s2 * P.weld_len : c2 * P.weld_len;                                       \
    real Cx = x[own] - (tag##x + r2x), Cy = y[own] - (tag##y + r2y), Cz = th[own] - tag##t;     \
    const real PI_ = real(3.14159265358979323846);                                              \
    if (Cz < -PI_) Cz += real(2) * PI_;                                                         \
    if (Cz > PI_) Cz -= real(2) * PI_;                                                          \
    real linErr = r_sqrt(Cx * Cx + Cy * Cy), angErr = r_abs(Cz);                                \
    weld_solve33<real>(Cx, Cy, Cz, r2x, r2y, inertia, weld_alpha, inv_two_im, L##x, L##y, L##z); \
    const real pm = done ? real(0) : hm[side];                                                  \
    L##x *= pm; L##y *= pm; L##z *= pm;                                                         \
    x[own] += L##x * im; y[own] += L##y * im; th[own] += ii * L##z;                             \
    solved = solved && (!has[side] || ((linErr <= P.lin_tol) && (angErr <= P.ang_tol)));        \
  }
#define VSR_WELD_POS_B2(side, own, horiz, slot0, L)                                             \
  {                                                                                             \
    const real pm = done ? real(0) : hm[side];                                                  \
    const real bx_ = pm * EXG(slot0, L##x, nl[side]), by_ = pm * EXG(slot0 + 1, L##y, nl[side]), \
               bz_ = pm * EXG(slot0 + 2, L##z, nl[side]);                                       \
    const real r2x = horiz ? cs_[own] * P.weld_len : -sn_[own] * P.weld_len;                    \
    const real r2y = horiz ? sn_[own] * P.weld_len : cs_[own] * P.weld_len;                     \
    x[own] -= bx_ * im; y[own] -= by_ * im; th[own] -= ii * (r2x * by_ - r2y * bx_ + bz_);      \
  }
          {
            VSR_PUBLISH_P()
            VSR_GATHER_P(l1, 1, 0) VSR_GATHER_P(l2, 2, 0)
            VSR_GATHER_CS(l1, 1, 0) VSR_GATHER_CS(l2, 2, 0)
            VSR_WELD_POS_B1(0, 0, l1, true, la) VSR_WELD_POS_B1(0, 3, l2, true, lb)
            EXFLIP()
            EXP(0, lax) EXP(1, lay) EXP(2, laz) EXP(3, lbx) EXP(4, lby) EXP(5, lbz) EXSYNC()
            VSR_WELD_POS_B2(1, 1, true, 0, la) VSR_WELD_POS_B2(1, 2, true, 3, lb)
            EXFLIP()
          }
          // masses 0, 1 are body2 of the vertical welds: their angles just changed
          r_sincos(th[0], &sn_[0], &cs_[0]);
          r_sincos(th[1], &sn_[1], &cs_[1]);
          {
            VSR_PUBLISH_P()
            VSR_GATHER_P(d0, 0, 2) VSR_GATHER_P(d1, 1, 2)
            VSR_GATHER_CS(d0, 0, 2) VSR_GATHER_CS(d1, 1, 2)
            VSR_WELD_POS_B1(2, 3, d0, false, da) VSR_WELD_POS_B1(2, 2, d1, false, db)
            EXFLIP()
            EXP(0, dax) EXP(1, day) EXP(2, daz) EXP(3, dbx) EXP(4, dby) EXP(5, dbz) EXSYNC()
            VSR_WELD_POS_B2(3, 0, false, 0, da) VSR_WELD_POS_B2(3, 1, false, 3, db)
            EXFLIP()
          }
#undef VSR_GATHER_P
#undef VSR_PUBLISH_P
#undef VSR_GATHER_CS
#undef VSR_OWN_CS
#undef VSR_WELD_POS_B1
#undef VSR_WELD_POS_B2
          // masses 1, 2 are body2 of the next iteration's horizontal welds
          r_sincos(th[1], &sn_[1], &cs_[1]);
          r_sincos(th[2], &sn_[2], &cs_[2]);
          }  // !wdone
          // robot-wide (island-wide) vote
          if (BLK) {
            if (robot_sync_and(bar_id, rthreads, solved || done)) done = true;
          } else {
            unsigned bal = __ballot_sync(0xffffffffu, solved || done);
            if ((bal & robot_mask) == robot_mask) done = true;
          }
        }
      }

      r_sincos(th[0], &sn_[0], &cs_[0]);
      r_sincos(th[3], &sn_[3], &cs_[3]);
      // ---- World.solveTOI (vsr_settings.continuous_detection): every mass against the ground
      //      polygons it is about to touch.  A mass whose swept box stays above the terrain's height map is skipped, and
      //      so is a mass in contact whose swept box lies over ONE ground polygon (it is in contact with that one: a
      //      resting foot on a long segment).  The few that remain are taken one after the other by the WHOLE warp
      //      (ccd_mass_warp); their poses and velocities travel through the global scratch ccd_buf (rows 0-11 pose at
      //      the start of the step, 12-23 pose now / the result, 24-35 velocity, 36-39 box half extent at the start).
      if (CCD && P.ccd) {
        const size_t ccd_t = (size_t)gridDim.x * blockDim.x, ccd_i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        unsigned ccand = 0;
        int cseg[4] = {0, 0, 0, 0};
#if defined(VSR_CCD_EXP) && VSR_CCD_EXP == 2
        if (false) {
#elif defined(VSR_CCD_EXP) && VSR_CCD_EXP == 4
        if (active && P.n_ticks < 0) {  // (never: the code stays, nothing of it runs)
#else
        if (active) {
#endif
          // Phase A, all four masses at once and without a branch, so that the twelve scratch loads are ONE round trip
          // to L2 and the height-map loads another (mass after mass, with an early `continue` each, the lockstep block
          // paid four dependent chains per tick: 13 ms of the launch): which swept boxes reach down to the terrain?
          const real rr_ = P.half * real(1.4142135623730951) + real(0.01);
          real p0x_[4], p0y_[4], ylo4_[4];
          unsigned near_ = 0;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            p0x_[k] = P.ccd_buf[(k * 3 + 0) * ccd_t + ccd_i];
            p0y_[k] = P.ccd_buf[(k * 3 + 1) * ccd_t + ccd_i];
            ylo4_[k] = P.ccd_buf[(36 + k) * ccd_t + ccd_i];  // (the box half extent at the start, for now)
          }
#pragma unroll
          for (int k = 0; k < 4; k++) {
            int c0 = (int)floorf((float)((r_min(p0x_[k], x[k]) - rr_ - P.hgrid_x0) * P.hgrid_inv));
            int c1 = (int)floorf((float)((r_max(p0x_[k], x[k]) + rr_ - P.hgrid_x0) * P.hgrid_inv));
            c0 = max(0, min(c0, P.hgrid_n - 1));
            c1 = max(0, min(c1, P.hgrid_n - 1));
            real top = r_max(P.hgrid[c0], P.hgrid[c1]);
            for (int c = c0 + 1; c < c1; c++) top = r_max(top, P.hgrid[c]);  // (cells are 2 wide: rarely taken)
            // the swept box proper (the pass itself skips a polygon whose top lies below the box; 1e-4 and the 0.01 in
            // rr_: float rounding)
            const real e1_ = P.half * (r_abs(cs_[k]) + r_abs(sn_[k]));
            ylo4_[k] = r_min(p0y_[k] - ylo4_[k], y[k] - e1_) - real(1e-4);
            if (!(ylo4_[k] > top)) near_ |= 1u << k;
          }
          // Phase B, the few that remain: against the polygons under the box, as the pass itself does.  The height map
          // alone is not enough: the robots start one unit away from the border slope, whose cell is 100 high.
#pragma unroll
          for (int k = 0; k < 4; k++) {
#if defined(VSR_CCD_EXP) && VSR_CCD_EXP == 5
            if (!((near_ >> k) & 1u) || P.n_ticks > 0) continue;  // (phase A only)
#else
            if (!((near_ >> k) & 1u)) continue;
#endif
            const real p0x = p0x_[k], p0y = p0y_[k], ylo_ = ylo4_[k];
            const real xlo_ = r_min(p0x, x[k]) - rr_, xhi_ = r_max(p0x, x[k]) + rr_;
            int s_ = hint[k];
            while (s_ > 0 && P.terr_x[s_] > xlo_) --s_;
            while (s_ + 2 < P.terr_n && P.terr_x[s_ + 1] < xlo_) ++s_;
            // a mass in contact whose swept box lies over ONE polygon is in contact with that one (first: it is what
            // the feet on the ground are, every tick)
            if (((cmask >> k) & 1u) && P.terr_x[s_] <= xlo_ && !(s_ + 2 < P.terr_n && P.terr_x[s_ + 1] <= xhi_)) continue;
            bool hit_ = false;
            // (and a polygon further away than the mass can travel yields no time of impact: see ccd_mass_warp)
            const real trav_ = (r_sqrt(vx[k] * vx[k] + vy[k] * vy[k]) + rr_ * r_abs(w[k])) * dt + real(1e-3);
            real s0_, c0_;
            r_sincos(P.ccd_buf[(k * 3 + 2) * ccd_t + ccd_i], &s0_, &c0_);  // (the pose at the start of the step)
            for (int seg = s_; seg + 1 < P.terr_n && P.terr_x[seg] <= xhi_; ++seg) {
              const real sx0_ = P.terr_x[seg], sy0_ = P.terr_y[seg], sy1_ = P.terr_y[seg + 1], ex_ = P.terr_x[seg + 1] - sx0_, ey_ = sy1_ - sy0_;
              const real il_ = real(1) / r_sqrt(ex_ * ex_ + ey_ * ey_), lnx_ = -ey_ * il_, lny_ = ex_ * il_;
              const real gap_ = (p0x - sx0_) * lnx_ + (p0y - sy0_) * lny_
                                - P.half * (r_abs(c0_ * lnx_ + s0_ * lny_) + r_abs(c0_ * lny_ - s0_ * lnx_)) - real(0.01);
              if (!(ylo_ > r_max(sy0_, sy1_)) && !(gap_ > trav_)) hit_ = true;
            }
            if (!hit_) continue;
            ccand |= 1u << k;
            cseg[k] = s_;  // (the first polygon under the swept box: where the pass starts its walk)
            P.ccd_buf[(12 + k * 3 + 0) * ccd_t + ccd_i] = x[k];
            P.ccd_buf[(12 + k * 3 + 1) * ccd_t + ccd_i] = y[k];
            P.ccd_buf[(12 + k * 3 + 2) * ccd_t + ccd_i] = th[k];
            P.ccd_buf[(24 + k * 3 + 0) * ccd_t + ccd_i] = vx[k];
            P.ccd_buf[(24 + k * 3 + 1) * ccd_t + ccd_i] = vy[k];
            P.ccd_buf[(24 + k * 3 + 2) * ccd_t + ccd_i] = w[k];
          }
        }
#ifdef VSR_CCD_COUNT
        if (ccand) atomicAdd(&g_ccd_count, (unsigned long long)__popc(ccand));
#endif
        // Block-wide pool: the candidates of all the block's robots go into one list and the warps take them in
        // turn.  (Per warp, the block waited at the next barrier for the warp whose robot stands on the border slope:
        // 8 candidates there every tick, none elsewhere.)  The list lives in rows 62-71 of the spring tables, which
        // are dead for every warp once the block has passed the first barrier below.
#if defined(VSR_CCD_EXP) && VSR_CCD_EXP == 3
        const bool ccd_any = __syncthreads_or(ccand != 0u) != 0 && P.n_ticks < 0;
#else
        const bool ccd_any = __syncthreads_or(ccand != 0u) != 0;  // (every warp is past its position loop)
#endif
        if (ccd_any) {
          real* const park = reinterpret_cast<real*>(spr);
          int* const clist = reinterpret_cast<int*>(park + 62 * SPR_STRIDE);
          const int ccap = 10 * SPR_STRIDE;
#pragma unroll
          for (int k = 0; k < 4; k++)
            if ((ccand >> k) & 1u) {
              const int pos_ = atomicAdd(&ccd_n, 1);
              if (pos_ < ccap)
                clist[pos_] = (int)threadIdx.x | (k << 10) | (int)(((flip >> k) & 1u) << 12) | (int)(((cmask >> k) & 1u) << 13) | (cseg[k] << 14);
              else
                status |= VSR_STATUS_CONTACT_OVERFLOW;  // (cannot happen: 4 candidates per thread, 10 rows)
            }
          // The pass is a call into double-precision code: every register its call tree may touch is one the kernel
          // cannot keep a long-lived value in, and ptxas answers by keeping values of the VELOCITY LOOP in local memory
          // for the whole tick.  So there is ONE call site, and the live ranges are split by hand around it: the
          // long-lived state is parked in rows 0-61 of the spring tables ([field][thread]: conflict-free).
#pragma unroll
          for (int q = 0; q < 18; q++) park[q * SPR_STRIDE + tid] = simp[q];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            park[(18 + q) * SPR_STRIDE + tid] = vx[q]; park[(22 + q) * SPR_STRIDE + tid] = vy[q]; park[(26 + q) * SPR_STRIDE + tid] = w[q];
            park[(30 + q) * SPR_STRIDE + tid] = cs_[q]; park[(34 + q) * SPR_STRIDE + tid] = sn_[q];
#pragma unroll
            for (int z = 0; z < 2; z++) {
              park[(38 + q * 2 + z) * SPR_STRIDE + tid] = wix[q][z]; park[(46 + q * 2 + z) * SPR_STRIDE + tid] = wiy[q][z];
              park[(54 + q * 2 + z) * SPR_STRIDE + tid] = wiz[q][z];
            }
          }
          __syncthreads();  // the list and the scratch columns are complete
          const int cn_ = min(ccd_n, ccap);
#pragma unroll 1
          for (int ci = (int)(threadIdx.x >> 5); ci < cn_; ci += (int)(blockDim.x >> 5)) {  // (warp-uniform)
            const int e_ = clist[ci];
            const int ot = e_ & 1023, k = (e_ >> 10) & 3;
            const size_t oi = (size_t)blockIdx.x * blockDim.x + ot;  // the owner's column of the scratch
            Pose3<real> p0, p1;
            p0.x = P.ccd_buf[(k * 3 + 0) * ccd_t + oi]; p0.y = P.ccd_buf[(k * 3 + 1) * ccd_t + oi]; p0.th = P.ccd_buf[(k * 3 + 2) * ccd_t + oi];
            p1.x = P.ccd_buf[(12 + k * 3 + 0) * ccd_t + oi]; p1.y = P.ccd_buf[(12 + k * 3 + 1) * ccd_t + oi]; p1.th = P.ccd_buf[(12 + k * 3 + 2) * ccd_t + oi];
            const real ovx = P.ccd_buf[(24 + k * 3 + 0) * ccd_t + oi], ovy = P.ccd_buf[(24 + k * 3 + 1) * ccd_t + oi],
                       ow_ = P.ccd_buf[(24 + k * 3 + 2) * ccd_t + oi];
            const CBody<real>* const ob = blk_cb + (size_t)ot * 8 + 2 * k + ((e_ >> 12) & 1);
#ifdef VSR_CCD_COUNT
            const long long clk0_ = clock64();
#endif
#if defined(VSR_CCD_EXP) && VSR_CCD_EXP == 1
            const Pose3<real> r_ = p1;
#else
            const Pose3<real> r_ = ccd_mass_warp(tpar, p0, p1, ovx, ovy, ow_, dt, ob, (e_ >> 13) & 1, P.lin_tol, P.max_corr, im, ii,
                                                 (int)((unsigned)e_ >> 14));
#endif
#ifdef VSR_CCD_COUNT
            if (lane == 0) atomicAdd(&g_ccd_count2, (unsigned long long)(clock64() - clk0_));
#endif
            if (lane == 0 && (r_.x != p1.x || r_.y != p1.y || r_.th != p1.th)) {
              P.ccd_buf[(12 + k * 3 + 0) * ccd_t + oi] = r_.x;
              P.ccd_buf[(12 + k * 3 + 1) * ccd_t + oi] = r_.y;
              P.ccd_buf[(12 + k * 3 + 2) * ccd_t + oi] = r_.th;
            }
          }
#pragma unroll
          for (int q = 0; q < 18; q++) simp[q] = park[q * SPR_STRIDE + tid];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            vx[q] = park[(18 + q) * SPR_STRIDE + tid]; vy[q] = park[(22 + q) * SPR_STRIDE + tid]; w[q] = park[(26 + q) * SPR_STRIDE + tid];
            cs_[q] = park[(30 + q) * SPR_STRIDE + tid]; sn_[q] = park[(34 + q) * SPR_STRIDE + tid];
#pragma unroll
            for (int z = 0; z < 2; z++) {
              wix[q][z] = park[(38 + q * 2 + z) * SPR_STRIDE + tid]; wiy[q][z] = park[(46 + q * 2 + z) * SPR_STRIDE + tid];
              wiz[q][z] = park[(54 + q * 2 + z) * SPR_STRIDE + tid];
            }
          }
          __syncthreads();  // every candidate's result is in the scratch
          if (threadIdx.x == 0) ccd_n = 0;  // (the next append is two barriers away)
#pragma unroll
          for (int k = 0; k < 4; k++)
            if ((ccand >> k) & 1u) {
              const real nx_ = P.ccd_buf[(12 + k * 3 + 0) * ccd_t + ccd_i], ny_ = P.ccd_buf[(12 + k * 3 + 1) * ccd_t + ccd_i],
                         nth_ = P.ccd_buf[(12 + k * 3 + 2) * ccd_t + ccd_i];
              if (nx_ != x[k] || ny_ != y[k] || nth_ != th[k]) {
                x[k] = nx_; y[k] = ny_; th[k] = nth_;
                r_sincos(th[k], &sn_[k], &cs_[k]);
              }
            }
        }
      }
      }  // tick >= 0
      if (sync_mode & 64) __syncthreads();
      // ---- World.detect: new manifolds for Touch and for the next step
      const unsigned old_cmask = cmask;
      cmask = 0;
      unsigned gcn = 0;  // ground constraints per mass, 4 bits each
      if (cap * VSR_POOL_FIELDS >= 416 * wpr) {
        // Only a few masses of a warp are near the ground.  Those (by the coarse height map; a mass that
        // had contacts always) are listed for the whole warp and detected one per lane, writing straight
        // into their owners' records; the pool -- idle during detection -- is the scratch space.
        int* const scr = reinterpret_cast<int*>(pool) + (BLK ? wir * 416 : 0);  // [32][5] hints + flags, [128] list, [32][4] results
#pragma unroll
        for (int k = 0; k < 4; k++) {
          mt[VSR_MTK(k, 0)] = x[k]; mt[VSR_MTK(k, 1)] = y[k]; mt[VSR_MTK(k, 2)] = cs_[k]; mt[VSR_MTK(k, 3)] = sn_[k];
          scr[lane * 5 + k] = hint[k];
        }
        scr[lane * 5 + 4] = (int)((flip & 15u) | (old_cmask << 4));
        unsigned cand = 0;
        if (active) {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const real e = P.half * (r_abs(cs_[k]) + r_abs(sn_[k])) + real(0.01);  // (+ a margin: the cell index is rounded)
            int c0 = (int)floorf((float)((x[k] - e - P.hgrid_x0) * P.hgrid_inv)), c1 = (int)floorf((float)((x[k] + e - P.hgrid_x0) * P.hgrid_inv));
            c0 = max(0, min(c0, P.hgrid_n - 1));
            c1 = max(0, min(c1, P.hgrid_n - 1));
            real hm = P.hgrid[c0];
            for (int c = c0 + 1; c <= c1; c++) hm = r_max(hm, P.hgrid[c]);
            if (((old_cmask >> k) & 1u) || !(y[k] - e > hm)) cand |= 1u << k;
          }
        }
        const int ncand = __popc(cand);
        int incl = ncand;
#pragma unroll
        for (int d2 = 1; d2 < 32; d2 <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, d2);
          if (lane >= d2) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        {
          int pos = incl - ncand;
#pragma unroll
          for (int k = 0; k < 4; k++)
            if (cand & (1u << k)) scr[160 + pos++] = (lane << 2) | k;
        }
        __syncwarp();
        for (int i = lane; i < total; i += 32) {
          const int code = scr[160 + i];
          const int ol = code >> 2, k = code & 3;
          const int fl = scr[ol * 5 + 4];
          int h = scr[ol * 5 + k];
          const int fb = (fl >> k) & 1, had = (fl >> (4 + k)) & 1;
          const int tl = (wib << 5) + ol;
          CBody<real>* const ob = blk_cb + (size_t)tl * 8;
          int cn = 0;
          const int ov = detect_mass(tpar, mt[VSR_MTI(k, 0, tl)], mt[VSR_MTI(k, 1, tl)], mt[VSR_MTI(k, 2, tl)], mt[VSR_MTI(k, 3, tl)],
                                     h, ob[2 * k + fb], had, ob[2 * k + (fb ^ 1)], cn, im, ii, k);
          scr[288 + ol * 4 + k] = cn | (ov << 4) | (h << 8);
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (cand & (1u << k)) {
            const int r = scr[288 + lane * 4 + k];
            hint[k] = r >> 8;
            flip ^= 1u << k;
            if (r & 16) status |= STATUS_CONTACT_OVERFLOW;
            const int cn = r & 15;
            if (cn > 0) cmask |= 1u << k;
            gcn |= (unsigned)cn << (4 * k);
          }
        __syncwarp();
      } else if (active) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          // cheap reject: far above the highest point of the terrain near this mass
          int cn = 0;
          int ov = detect_mass(tpar, x[k], y[k], cs_[k], sn_[k], hint[k], CBODY(k), (int)((old_cmask >> k) & 1u),
                               cbase[2 * k + (((flip >> k) & 1u) ^ 1u)], cn, im, ii, k);
          flip ^= 1u << k;
          if (ov) status |= STATUS_CONTACT_OVERFLOW;
          if (cn > 0) cmask |= 1u << k;
          gcn |= (unsigned)cn << (4 * k);
#ifdef VSR_DEBUG_DUMP
          if (rb == P.debug_robot && tick >= P.debug_tick && tick <= P.debug_tick_end)
            for (int q = 0; q < cn; q++) {
              const SCon<real>& c = CBODY(k).c[q];
              printf("[kernel] tick %d body %d seg %d n %d normal (%.9g, %.9g) pose (%.9g %.9g %.9g)\n", tick, v * 4 + k,
                     c.seg, c.n, (double)c.nx, (double)c.ny, (double)x[k] + P.origin_x, (double)y[k], (double)th[k]);
              for (int z = 0; z < c.n; z++)
                printf("           p (%.9g, %.9g) depth %.6e id %d jn %.6e jt %.6e\n", (double)c.p[z].px + P.origin_x,
                       (double)c.p[z].py, (double)c.p[z].depth, c.p[z].id, (double)c.p[z].jn, (double)c.p[z].jt);
            }
#endif
        }
      }

      // ---- intra-robot pairs: every mass against every later mass of the robot that is not welded
      //      to it.  Order inside a robot = (owner voxel, partner voxel, partner mass, own mass),
      //      the oracle's canonical enumeration.
      const int rbase = BLK ? lane0 : (wib << 5) + lane0;  // block thread of this robot's first voxel
      unsigned smask = 0;  // slots of the new list that hold a constraint
      unsigned long long lvpack[2] = {0ull, 0ull};  // dependency level of each of this thread's constraints, 8 bits each
      if (P.self_coll) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          mt[VSR_MTK(k, 0)] = x[k]; mt[VSR_MTK(k, 1)] = y[k]; mt[VSR_MTK(k, 2)] = cs_[k]; mt[VSR_MTK(k, 3)] = sn_[k];
        }
        const real vcx = real(0.25) * ((x[0] + x[1]) + (x[2] + x[3])), vcy = real(0.25) * ((y[0] + y[1]) + (y[2] + y[3]));
        real vrad = real(0);
#pragma unroll
        for (int k = 0; k < 4; k++) vrad = r_max(vrad, r_max(r_abs(x[k] - vcx), r_abs(y[k] - vcy)));
        aux[0 * bdim + tid] = vcx; aux[1 * bdim + tid] = vcy; aux[2 * bdim + tid] = vrad;
        SSYNC()
        SCon<real>* const newc = (sbase + ((sst >> 31) ^ 1u))->c;
        // Candidate q of a thread = (partner thread << 4) | partner mass << 2 | own mass, bound for slot q of the thread's
        // list.  With the pool as scratch space the candidates are written there as they are found and the warp's
        // candidates are spread over its lanes afterwards (one narrowphase round); without it every thread runs the
        // narrowphase of a candidate on the spot.
        const bool pooled_np = cap * VSR_POOL_FIELDS >= 576 * wpr;
        // scratch: [32][16] candidates and the [512] list as 16-bit words, [32] meta and [32] results as words
        int* const scr = reinterpret_cast<int*>(pool) + (BLK ? wir * 576 : 0);
        unsigned short* const scand = reinterpret_cast<unsigned short*>(scr);        // words [0, 256)
        unsigned short* const slist = reinterpret_cast<unsigned short*>(scr + 288);  // words [288, 544)
        int n_cand = 0;
        if (active) {
          const real rr = real(2.8284271247461903) * P.half;  // two half diagonals
          real ae[4];
#pragma unroll
          for (int k = 0; k < 4; k++) ae[k] = P.half * (r_abs(cs_[k]) + r_abs(sn_[k]));
          const real vreach = vrad + r_max(r_max(ae[0], ae[1]), r_max(ae[2], ae[3]));  // mass centres + their extents
          const int j_right = has[1] ? nl[1] - lane0 : -1, j_up = has[3] ? nl[3] - lane0 : -1;
          for (int j = v; j < nvox; j++) {
            const int tj = rbase + j;
            if (j != v) {
              const real lim = vrad + aux[2 * bdim + tj] + rr;
              if (r_abs(aux[0 * bdim + tj] - vcx) > lim || r_abs(aux[1 * bdim + tj] - vcy) > lim) continue;
            }
            // welded pairs, bit (own mass * 4 + partner mass): right neighbour 1-0, 2-3; upper 0-3, 1-2
            const unsigned wm = j == j_right ? ((1u << 4) | (1u << 11)) : (j == j_up ? ((1u << 3) | (1u << 6)) : 0u);
#pragma unroll 1
            for (int kb = 0; kb < 4; kb++) {
              const real bx = mt[VSR_MTI(kb, 0, tj)], by = mt[VSR_MTI(kb, 1, tj)];
              // half extent of the partner's bounding box
              const real be = P.half * (r_abs(mt[VSR_MTI(kb, 2, tj)]) + r_abs(mt[VSR_MTI(kb, 3, tj)]));
              // the partner mass against this voxel's box first: most masses of a neighbour are out of reach
              if (r_abs(bx - vcx) > vreach + be || r_abs(by - vcy) > vreach + be) continue;
#pragma unroll
              for (int ka = 0; ka < 4; ka++) {
                if (j == v && kb <= ka) continue;
                if ((wm >> (ka * 4 + kb)) & 1u) continue;
                // disjoint bounding boxes: separated (exact, unlike the narrowphase margin: safe side)
                const real lim = be + ae[ka];
                if (r_abs(x[ka] - bx) > lim || r_abs(y[ka] - by) > lim) continue;
                if (n_cand == VSR_MAX_CAND) { status |= STATUS_PAIR_OVERFLOW; continue; }
                if (pooled_np) {
                  scand[lane * VSR_MAX_CAND + n_cand] = (unsigned short)((tj << 4) | (kb << 2) | ka);
                } else if (self_pair(cpar, P.half, x[ka], y[ka], cs_[ka], sn_[ka], bx, by, mt[VSR_MTI(kb, 2, tj)],
                                     mt[VSR_MTI(kb, 3, tj)], ka, tj, kb, S_CUR->c, S_MASK, &newc[n_cand])) {
                  smask |= 1u << n_cand;
                }
                n_cand++;
              }
            }
          }
        }
        if (pooled_np) {
          scr[256 + lane] = (int)S_MASK;
          int incl = n_cand;
#pragma unroll
          for (int d2 = 1; d2 < 32; d2 <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d2);
            if (lane >= d2) incl += t;
          }
          const int total = __shfl_sync(0xffffffffu, incl, 31);
          for (int q = 0, pos = incl - n_cand; q < n_cand; q++) slist[pos + q] = (unsigned short)((lane << 4) | q);
          scr[544 + lane] = 0;
          __syncwarp();
          const unsigned cur = sst >> 31;
          for (int i = lane; i < total; i += 32) {
            const int it = slist[i];
            const int ol = it >> 4, q = it & 15;
            const int code = scand[ol * VSR_MAX_CAND + q];
            const int tj = code >> 4, kb = (code >> 2) & 3, ka = code & 3;
            const int tl = (BLK ? 0 : 0) + (wib << 5) + ol;
            SList<real>* const own = blk_sl + (size_t)tl * 2;
            const int got = self_pair(cpar, P.half, mt[VSR_MTI(ka, 0, tl)], mt[VSR_MTI(ka, 1, tl)], mt[VSR_MTI(ka, 2, tl)],
                                      mt[VSR_MTI(ka, 3, tl)], mt[VSR_MTI(kb, 0, tj)], mt[VSR_MTI(kb, 1, tj)], mt[VSR_MTI(kb, 2, tj)],
                                      mt[VSR_MTI(kb, 3, tj)], ka, tj, kb, own[cur].c, (unsigned)scr[256 + ol], &own[cur ^ 1u].c[q]);
            if (got) atomicOr(&scr[544 + ol], 1 << q);
          }
          __syncwarp();
          smask = (unsigned)scr[544 + lane];
          __syncwarp();
        }
        const int n_new = __popc(smask);
        // position of this thread's constraints in its robot's order
        int incl = n_new;
#pragma unroll
        for (int d2 = 1; d2 < 32; d2 <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, d2);
          if (lane >= d2) incl += t;
        }
        int s_start, s_tot;
        if (BLK) {
          if (lane == 31) lhd[66 + wir] = incl;
          SSYNC()
          int before = 0, all = 0;
          for (int q = 0; q < wpr; q++) {
            const int c = lhd[66 + q];
            before += q < wir ? c : 0;
            all += c;
          }
          s_start = before + incl - n_new;
          s_tot = all;
        } else {
          const int first = lane_ok ? lane0 : lane, last = lane_ok ? lane0 + nvox - 1 : lane;
          const int below = __shfl_sync(0xffffffffu, incl, first > 0 ? first - 1 : 0);
          const int upto = __shfl_sync(0xffffffffu, incl, last);
          const int base = first > 0 ? below : 0;
          s_start = incl - n_new - base;
          s_tot = (int)__reduce_max_sync(0xffffffffu, (unsigned)(lane_ok ? upto - base : 0));
        }
        // dependency levels: a constraint runs one round after the last earlier constraint that
        // touches one of its two masses (ground constraints of a mass count as round 0).  Walk the
        // robot's list in order, one constraint at a time; the counters replace the mass table.
        SSYNC()
        real* const cnt = mt;  // small integers held as reals: each thread keeps to its own columns
#pragma unroll
        for (int k = 0; k < 4; k++) cnt[k * MTS + tid] = (real)((gcn >> (4 * k)) & 15u);
        SSYNC()
        for (int i = 0; i < s_tot; i++) {
          const int li = i - s_start;
          if (li >= 0 && li < n_new) {
            const int sl = (int)__fns(smask, 0, li + 1);  // the li-th slot in use
            const int ia = newc[sl].ka * MTS + tid, ib = newc[sl].kb * MTS + newc[sl].tb;
            const int ca = (int)cnt[ia], cb2 = (int)cnt[ib];
            int lv = ca > cb2 ? ca : cb2;
            cnt[ia] = (real)(lv + 1);
            cnt[ib] = (real)(lv + 1);
            if (lv >= VSR_MAX_LEVELS) { lv = VSR_MAX_LEVELS - 1; status |= STATUS_LEVEL_OVERFLOW; }
            lvpack[sl >> 3] |= (unsigned long long)lv << (8 * (sl & 7));
          }
          SSYNC()
        }
        sst = smask | ((sst ^ 0x80000000u) & 0x80000000u);
      }

      // ---- the scope's constraint list for the next step, by dependency level: the q-th ground
      //      constraint of a mass has level q, the intra-robot ones follow (levels above)
      if (!BLK && !__any_sync(0xffffffffu, gcn != 0u || smask != 0u)) {
        if (lane == 0) lhd[1] = 0;  // nothing touches anything in this warp
        __syncwarp();
      } else {
        SSYNC()
        for (int q = st; q < 66; q += sdim) lhd[q] = 0;
        SSYNC()
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int cn = (int)((gcn >> (4 * k)) & 15u);
          for (int q = 0; q < cn; q++) atomicAdd(&lhd[2 + q], 1);
        }
        for (unsigned m_ = smask; m_; m_ &= m_ - 1u) {
          const int li = __ffs((int)m_) - 1;
          atomicAdd(&lhd[2 + ((unsigned)((lvpack[li >> 3] >> (8 * (li & 7))) & 255ull))], 1);
        }
        SSYNC()
        if (!BLK) {
          // one level per lane (VSR_MAX_LEVELS = 32): prefix sum of the counts by shuffles
          const int c = lhd[2 + lane];
          int incl = c;
#pragma unroll
          for (int d2 = 1; d2 < 32; d2 <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d2);
            if (lane >= d2) incl += t;
          }
          const unsigned used = __ballot_sync(0xffffffffu, c > 0);
          lhd[34 + lane] = incl - c;  // cursor: start of level `lane`
          lhd[2 + lane] = incl;       // end offset of level `lane`
          if (lane == 0) lhd[1] = used ? 32 - __clz(used) : 0;
        } else if (st == 0) {
          int run = 0, nl_ = 0;
          for (int l = 0; l < VSR_MAX_LEVELS; l++) {
            const int c = lhd[2 + l];
            lhd[34 + l] = run;  // cursor: start of level l
            run += c;
            lhd[2 + l] = run;   // end offset of level l
            if (c > 0) nl_ = l + 1;
          }
          lhd[1] = nl_;
        }
        SSYNC()
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int cn = (int)((gcn >> (4 * k)) & 15u);
          for (int q = 0; q < cn; q++)
            items[atomicAdd(&lhd[34 + q], 1)] = (tid << 8) | ((k * 2 + (int)((flip >> k) & 1u)) << 5) | (q << 1);
        }
        for (unsigned m_ = smask; m_; m_ &= m_ - 1u) {
          const int li = __ffs((int)m_) - 1;
          items[atomicAdd(&lhd[34 + ((unsigned)((lvpack[li >> 3] >> (8 * (li & 7))) & 255ull))], 1)] = (tid << 8) | (int)((sst >> 31) << 5) | (li << 1) | 1;
        }
        SSYNC()
      }

      // =========================== Robot.act(t) ===========================
      __syncthreads();  // REQUIRED: detection's use of `aux` ends here, the controllers' staging begins (see the tick's top)
      if (tick < 0) continue;
      // ---- Voxel.act: energies; geometry shared by the sensors
      const real* s4 = sn_;
      const real* c4 = cs_;
      // outer corners getIndexedVertex(k, 3-k): NW(-h,+h) NE(+h,+h) SE(+h,-h) SW(-h,-h)
      real ox[4], oy[4];
      {
        const real h = P.half;
        ox[0] = x[0] + c4[0] * (-h) - s4[0] * (+h); oy[0] = y[0] + s4[0] * (-h) + c4[0] * (+h);
        ox[1] = x[1] + c4[1] * (+h) - s4[1] * (+h); oy[1] = y[1] + s4[1] * (+h) + c4[1] * (+h);
        ox[2] = x[2] + c4[2] * (+h) - s4[2] * (-h); oy[2] = y[2] + s4[2] * (+h) + c4[2] * (-h);
        ox[3] = x[3] + c4[3] * (-h) - s4[3] * (-h); oy[3] = y[3] + s4[3] * (-h) + c4[3] * (-h);
      }
      real area = real(0);
#pragma unroll
      for (int i = 0; i < 4; i++) area = area + ox[i] * (oy[(i + 1) & 3] - oy[(i + 3) & 3]);
      area = real(0.5) * r_abs(area);
      const real area_ratio = area / P.side / P.side;
      ar_energy = ar_energy + area_ratio * area_ratio;
      ctrl_energy = ctrl_energy + last_f * last_f;

      // ---- sensors (CompositeSensor.act: inner first)
      if (active && S.n_sens[v] > 0) {
        real mvx = (((vx[0] + vx[1]) + vx[2]) + vx[3]) / real(4);
        real mvy = (((vy[0] + vy[1]) + vy[2]) + vy[3]) / real(4);
        real ang = (r_atan2(y[1] - y[0], x[1] - x[0]) + r_atan2(y[2] - y[3], x[2] - x[3])) / real(2);
        int ro = 0;
        const int ns = S.n_sens[v];
        for (int si = 0; si < ns; si++) {
          // the record's scalars by four independent 8-byte loads (one latency); a struct copy would put the
          // whole record, ray directions included, on the stack because `dirs` is indexed dynamically
          const SensorDev* const sp = &S.sens[v][si];
          SensorHdr sd;
          {
            const int2 hw = __ldg(reinterpret_cast<const int2*>(sp));
            const int2 iw = __ldg(reinterpret_cast<const int2*>(sp) + 1);
            sd.max_norm = __ldg(&sp->max_norm);
            sd.noise = __ldg(&sp->noise);
            sd.base = (int)(int8_t)(hw.x & 255); sd.axes = (int)(int8_t)((hw.x >> 8) & 255);
            sd.rotated = (int)(int8_t)((hw.x >> 16) & 255); sd.agg = (int)(int8_t)((hw.x >> 24) & 255);
            sd.soft = (int)(int8_t)(hw.y & 255); sd.win = (int)(int8_t)((hw.y >> 8) & 255);
            sd.ch = (int)(int8_t)((hw.y >> 16) & 255); sd.dim = (int)(int8_t)((hw.y >> 24) & 255);
            sd.interval = __int_as_float(iw.x);
          }
          real raw[2] = {0, 0}, dmin = real(0), dmax = real(1);
          if (sd.base == SENSOR_TOUCH) {
            raw[0] = cmask ? real(1) : real(0);
          } else if (sd.base == SENSOR_AREA) {
            raw[0] = area_ratio;
            dmin = real(0.5);
            dmax = real(1.5);
          } else if (sd.base == SENSOR_ANGLE) {
            raw[0] = ang;
            dmin = real(-3.14159265358979323846);
            dmax = real(3.14159265358979323846);
          } else if (sd.base == SENSOR_CONSTANT) {
            raw[0] = (real)sd.max_norm;
            dmin = r_min(real(0), raw[0]);
            dmax = r_max(real(1), raw[0]);
          } else if (sd.base == SENSOR_APPLIED_FORCE) {
            raw[0] = last_f;  // the force of the previous control step (sensors act before the controller)
            dmin = real(-1);
          } else if (sd.base == SENSOR_MALFUNCTION) {
            raw[0] = (LIM && (bstate & 63u)) ? real(1) : real(0);  // BreakableVoxel.isBroken, before this tick's update
          } else if (sd.base == SENSOR_TIME_SIN) {
            raw[0] = (real)sin(2.0 * 3.14159265358979323846 * (double)sd.max_norm * P.tick_t[tick]);
            dmin = real(-1);
          } else if (sd.base == SENSOR_CRUMPLING) {
            real cnt = real(0);
            const real thr = P.side * real(0.2);
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
              for (int j = i + 1; j < 4; j++) {
                const real ddx = x[i] - x[j], ddy = y[i] - y[j];
                if (r_sqrt(ddx * ddx + ddy * ddy) < thr) cnt = cnt + real(1);
              }
            raw[0] = real(2) * cnt / real(12);
          } else if (sd.base == SENSOR_CONTROL_POWER) {
            const double tprev = tick > 0 ? P.tick_t[tick - 1] : 0.0;
            raw[0] = ctrl_energy / (real)(P.tick_t[tick] - tprev);
            dmax = (real)(1.0 / sd.max_norm);
          } else if (sd.base == SENSOR_LIDAR) {
            dmax = (real)sd.max_norm;  // the rays are cast below, one reading each
          } else {
            real sa, ca;
            r_sincos(ang, &sa, &ca);
            if (sd.axes & 1) raw[0] = sd.rotated ? mvx * ca + mvy * sa : mvx;
            if (sd.axes & 2) {
              // unit(angle + pi/2) = (-sin, cos) (Velocity.java:75 builds it numerically: same to 1 ulp)
              const real vy_ = sd.rotated ? -mvx * sa + mvy * ca : mvy;
              if (sd.axes & 1) raw[1] = vy_; else raw[0] = vy_;  // (static indices: raw stays in registers)
            }
            dmin = -(real)sd.max_norm;
            dmax = (real)sd.max_norm;
          }
          real val[2] = {raw[0], raw[1]};
          if (sd.agg != AGG_NONE) {
            const int depth = S.ring_depth;
            const int f = P.win_first[sd.win * P.n_ticks + tick];
#pragma unroll
            for (int d = 0; d < 2; d++) {  // aggregated sensors have one or two readings
              if (d >= sd.dim) break;
              real* ring = P.ring + ((size_t)(sd.ch + d)) * n_gl + gl;
              const size_t stride = (size_t)S.n_channels * n_gl;
              ring[(size_t)(tick % depth) * stride] = raw[d];
              if (sd.agg == AGG_AVERAGE) {
                real sum = real(0);
                for (int q = f; q < tick; q++) sum = sum + ring[(size_t)(q % depth) * stride];
                sum = sum + raw[d];
                val[d] = sum / (real)(tick - f + 1);
              } else if (sd.agg == AGG_DYNAMIC_NORM) {
                // DynamicNormalization.aggregate (:40-58): min / max over the kept window (the current reading
                // included); (r - min) / (max - min) is 0/0 = NaN while they coincide, and Java's Math.min / max
                // pass the NaN on (fminf / fmaxf would drop it)
                real mn = raw[d], mx = raw[d];
                for (int q = f; q < tick; q++) {
                  const real pv = ring[(size_t)(q % depth) * stride];
                  mn = r_min(mn, pv);
                  mx = r_max(mx, pv);
                }
                val[d] = mx > mn ? r_min(r_max((raw[d] - mn) / (mx - mn), real(0)), real(1)) : (real)NAN;
              } else {
                real li = (real)(P.tick_t[tick] - P.tick_t[f]);
                real first = (f == tick) ? raw[d] : ring[(size_t)(f % depth) * stride];
                val[d] = (f == tick) ? real(0) : (raw[d] - first) / li;
              }
            }
            if (sd.agg == AGG_TREND) {
              real ext = dmax - dmin;
              dmin = -ext / sd.interval;
              dmax = ext / sd.interval;
            } else if (sd.agg == AGG_DYNAMIC_NORM) {
              dmin = real(0);
              dmax = real(1);
            }
          }
          for (int d = 0; d < sd.dim; d++) {
            real o = d == 1 ? val[1] : val[0];
            if (sd.base == SENSOR_LIDAR) {
              // from Voxel.center (the mean of the 4 mass centres) at rayDirection + Voxel.getAngle
              const real ocx = (((x[0] + x[1]) + x[2]) + x[3]) / real(4), ocy = (((y[0] + y[1]) + y[2]) + y[3]) / real(4);
              real rs, rc;
              r_sincos((real)__ldg(&sp->dirs[d]) + ang, &rs, &rc);
              o = lidar_cast(tpar, ocx, ocy, rc, rs, (real)sd.max_norm, hint[0]);
            }
            const real o_in = o;
            if (sd.soft == 1) {
              real cl = r_min(r_max(o, dmin), dmax);
              real nv = (cl - dmin) / (dmax - dmin);
              o = r_tanh(((nv * real(2)) - real(1)) * real(2)) / real(2) + real(0.5);
            } else if (sd.soft == 2) {  // Normalization: DoubleRange.normalize
              o = (r_min(r_max(o, dmin), dmax) - dmin) / (dmax - dmin);
            }
            // (Java's Math.min / Math.max return NaN for a NaN argument; only DynamicNormalization produces one)
            if (sd.agg == AGG_DYNAMIC_NORM && o_in != o_in) o = o_in;
            if (sd.noise > 0.0) {
              // Noisy.sense (Noisy.java:57-62): every instance owns a Random(seed), so a sensor of
              // dimension dim draws gaussians number tick * dim + d of the same stream (host table)
              const real ext = sd.soft ? real(1) : dmax - dmin;
              o = o + (real)(P.noise_tab[(size_t)tick * sd.dim + d] * sd.noise) * ext;
            }
            if (LIM) {  // the reading's final domain (Sensor.getDomains): what a RANDOM sensor malfunction draws from
              rd_lo[ro] = sd.soft ? real(0) : dmin;
              rd_hi[ro] = sd.soft ? real(1) : dmax;
            }
            readings[ro++] = o;  // (dynamic index: a small thread-local array, not registers)
          }
        }
      }

      // ---- BreakableVoxel.act after super.act (BreakableVoxel.java:127-168; the oracle's breakable_act): the readings
      //      snapshot, the trigger counters, one draw per threshold in enum order, the break (component, type), restore
      real effr_[LIM ? MAX_READ : 1];
      real* const effr = LIM ? effr_ : readings;  // what Voxel.getSensorReadings returns to a controller
      if (LIM) {
        const int nr_ = (active && S.n_sens[v] > 0) ? ((v + 1 < nvox ? S.read_off[v + 1] : S.n_readings) - S.read_off[v]) : 0;
        if (brk) {
          const double t_ = P.tick_t[tick];
          if (((bstate >> 2) & 3u) == 0u || !(bstate & 128u)) {
#pragma unroll 1
            for (int q = 0; q < nr_; q++) snap[q] = readings[q];
            bstate |= 128u;
          }
          bcnt[2] = (real)(((double)bcnt[2] + t_) - b_last_t);  // (the reference's order: counter + t - lastT)
          bcnt[0] = bcnt[0] + (ctrl_energy - b_last_ce);
          bcnt[1] = bcnt[1] + (ar_energy - b_last_are);
          b_last_t = t_;
          b_last_ce = ctrl_energy;
          b_last_are = ar_energy;
          bool breaking = false;
#pragma unroll 1
          for (int g = 0; g < 3; g++) {
            const double thr = S.brk_thr[v][g];
            if (thr != thr) continue;
            const double u = jr_next_double(bseed);
            if (u < 1.0 - tanh(thr / (double)(g == 0 ? bcnt[0] : (g == 1 ? bcnt[1] : bcnt[2])))) {
              bcnt[0] = bcnt[1] = bcnt[2] = real(0);
              const unsigned m0 = S.brk_mal[v][0], m1 = S.brk_mal[v][1], m2 = S.brk_mal[v][2];
              const int ncomp = (m0 != 0u) + (m1 != 0u) + (m2 != 0u);
              if (ncomp > 0) {
                breaking = true;
                int ci = jr_next_int(bseed, ncomp);
                int c = 0;
                if (m0 != 0u && ci-- == 0) c = 0;
                else if (m1 != 0u && ci-- == 0) c = 1;
                else c = 2;
                const unsigned mc = c == 0 ? m0 : (c == 1 ? m1 : m2);
                const int ty = (int)__fns(mc, 0, jr_next_int(bseed, __popc(mc)) + 1);  // the n-th type of the set
                bstate = (bstate & ~(3u << (2 * c))) | ((unsigned)ty << (2 * c));
                // updateStructureMalfunctionType (:244-258): at a break only
                bstate = (bstate & ~64u) | ((((bstate >> 4) & 3u) == 2u) ? 64u : 0u);
              }
            }
          }
          if (breaking) b_last_break = t_;
          if (t_ - b_last_break > S.brk_restore[v]) bstate &= ~63u;
        }
        // BreakableVoxel.getSensorReadings (:183-193), asked for by a sensing controller only
        if (CTRL == CTRL_MLP) {
          const unsigned ss_ = brk ? ((bstate >> 2) & 3u) : 0u;
#pragma unroll 1
          for (int q = 0; q < nr_; q++)
            effr_[q] = !brk ? readings[q]
                            : (ss_ == 1u ? real(0)
                                         : (ss_ == 3u ? (real)(jr_next_double(bseed) * (double)(rd_hi[q] - rd_lo[q]) + (double)rd_lo[q])
                                                      : snap[q]));
        }
      }

      // ---- controller.control -> Voxel.applyForce
      real f = real(0);
      if (CTRL == CTRL_PHASE_SIN) {
        // PhaseSin.java:61, evaluated in double on the accumulated double t
        double fr = (double)par[0], am = (double)par[1], ph = (double)par[2 + (active ? v : 0)];
        f = (real)(sin(2.0 * 3.14159265358979323846 * fr * P.tick_t[tick] + ph) * am);
      } else if (CTRL == CTRL_TABULATED) {
        f = par[(size_t)tick * nvox + (active ? v : 0)];
      } else if (P.dist_signals >= 0) {
        // DistributedSensing.computeControlSignals (:150-186): every voxel runs its own
        // MultiLayerPerceptron on (own readings, what its N, E, S, W neighbours emitted towards it at
        // the previous tick); output 0 is the control signal, the rest is what it emits now.
        // Dir: N = (0, -1) = the voxel below, E = right, S = above, W = left; a neighbour's signals
        // towards this voxel sit at the opposite direction's index (Dir.adjacent).
        const int sg = P.dist_signals;
        real a[VSR_DIST_WIDTH], b2[VSR_DIST_WIDTH];
        int nin = 0;
        {
          const int base = S.read_off[active ? v : 0];
          const int nr = active ? ((v + 1 < nvox ? S.read_off[v + 1] : S.n_readings) - base) : 0;
#pragma unroll 1
          for (int q = 0; q < nr; q++) a[q] = effr[q];
          nin = nr;
          // gather: side (kernel numbering 0 left, 1 right, 2 down, 3 up) of Dir d, and the slot the
          // neighbour wrote for the opposite direction
          EXP(0, lsig[0]) EXP(1, lsig[1]) EXP(2, lsig[2]) EXP(3, lsig[3]) EXP(4, lsig[4]) EXP(5, lsig[5])
          EXP(6, lsig[6]) EXP(7, lsig[7]) EXP(8, lsig[8]) EXP(9, lsig[9]) EXP(10, lsig[10]) EXP(11, lsig[11])
          EXP(12, lsig[12]) EXP(13, lsig[13]) EXP(14, lsig[14]) EXP(15, lsig[15])
          EXSYNC()
#pragma unroll
          for (int dir = 0; dir < 4; dir++) {
            const int side = dir == 0 ? 2 : (dir == 1 ? 1 : (dir == 2 ? 3 : 0));
            const int opp = (dir + 2) & 3;
#pragma unroll
            for (int q = 0; q < 4; q++) {
              // (non-directional: everybody reads what the neighbour keeps in slot 0)
              const real gd = EXG(opp * 4 + q, lsig[opp * 4 + q], nl[side]);
              const real gn = EXG(q, lsig[q], nl[side]);
              const real g = P.dist_nd ? gn : gd;
              if (q < sg) a[nin + dir * sg + q] = has[side] ? g : real(0);
            }
          }
          nin += 4 * sg;
          EXFLIP()
        }
        const real* W = par + (active ? S.dist_off[v] : 0);
        int np = nin;
#pragma unroll 1
        for (int q = 0; q < np; q++) a[q] = activation<real>(P.mlp_act, a[q]);
        real* pa = a;
        real* pb = b2;
#pragma unroll 1
        for (int li = 1; li < S.n_layers; li++) {
          const int nn = S.neurons[li];
          if (active) {
#pragma unroll 1
            for (int j = 0; j < nn; j++) {
              real sum = W[0];
#pragma unroll 1
              for (int k = 1; k < np + 1; k++) sum = sum + pa[k - 1] * W[k];
              pb[j] = activation<real>(P.mlp_act, sum);
              W += np + 1;
            }
          }
          np = nn;
          real* t = pa; pa = pb; pb = t;
        }
        f = active ? pa[0] : real(0);
#pragma unroll
        for (int dir = 0; dir < 4; dir++)
#pragma unroll
          for (int q = 0; q < 4; q++)
            if (q < sg && !(P.dist_nd && dir > 0)) lsig[dir * 4 + q] = active ? pa[1 + dir * sg + q] : real(0);
      } else {
        // CentralizedSensing + MultiLayerPerceptron.apply: inputs = readings of all voxels
        // in voxel order; lane j of the robot evaluates neurons j, j+nvox, ...
        const int per_robot = BLK ? wpr * P.smem_per_warp : P.smem_per_warp / (G > 0 ? G : 1);
        real* act0 = BLK ? stage_blk + (size_t)rib * per_robot : stage + (size_t)(lane_ok ? slot : 0) * per_robot;
        real* act1 = act0 + per_robot / 2;
        if (active) {
          int base = S.read_off[v];
          int nr = (v + 1 < nvox ? S.read_off[v + 1] : S.n_readings) - base;
#pragma unroll 1
          for (int q = 0; q < nr; q++) act0[base + q] = activation<real>(P.mlp_act, effr[q]);
        }
        SSYNC()
        real* a = act0;
        real* b = act1;
        for (int li = 1; li < S.n_layers; li++) {
          const int np = S.neurons[li - 1], nn = S.neurons[li];
          const real* W = par + S.layer_off[li];  // transposed: [k = 0..np][neuron]
          if (active) {
            for (int j = v; j < nn; j += nvox) {
              real sum = W[j];
              for (int k = 1; k < np + 1; k++) sum = sum + a[k - 1] * W[(size_t)k * nn + j];
              b[j] = activation<real>(P.mlp_act, sum);
            }
          }
          SSYNC()
          real* t = a; a = b; b = t;
        }
        f = active ? a[v] : real(0);
        SSYNC()
      }
      if (LIM && brk) {  // BreakableVoxel.applyForce (:170-180)
        const unsigned as_ = bstate & 3u;
        if (as_ == 1u) f = real(0);
        else if (as_ == 2u) f = last_f;
        else if (as_ == 3u) f = (real)(jr_next_double(bseed) * 2.0 - 1.0);
      }
      if (r_abs(f) > real(1)) f = f > real(0) ? real(1) : real(-1);
      last_f = f;
#pragma unroll
      for (int c = 0; c < 3; c++)
        rest[c] = (f >= real(0)) ? P.rng_rest[c] - (P.rng_rest[c] - P.rng_min[c]) * f
                  : ((f < real(0)) ? P.rng_rest[c] + (P.rng_max[c] - P.rng_rest[c]) * -f
                                   : rest[c]);  // NaN: neither branch of Voxel.applyForce (:230-235), the rest distance stays

      // ---- optional per-tick dump (SnapshotListener / Observation path)
      if (P.trajectory && active && rb < P.traj_robots) {
        real* o = P.trajectory + (((size_t)rb * P.n_ticks + tick) * (size_t)(nvox * 4) + (size_t)v * 4) * 6;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          o[6 * k + 0] = x[k]; o[6 * k + 1] = y[k]; o[6 * k + 2] = th[k];
          o[6 * k + 3] = vx[k]; o[6 * k + 4] = vy[k]; o[6 * k + 5] = w[k];
        }
        // what VoxelPoly carries besides the geometry (Voxel.java:605-616): Touch.isTouchingGround and
        // getLastAppliedForce (the energies are running sums of these and of the area ratio)
        real* q = P.vtrace + (((size_t)rb * P.n_ticks + tick) * (size_t)nvox + (size_t)v) * 2;
        q[0] = cmask ? real(1) : real(0);
        q[1] = last_f;
        const int rbase_ = S.read_off[v];
        const int rn_ = (v + 1 < nvox ? S.read_off[v + 1] : S.n_readings) - rbase_;
        real* rt = P.rtrace + ((size_t)rb * P.n_ticks + tick) * (size_t)S.n_readings + rbase_;
        if (S.n_sens[v] > 0)
          for (int z = 0; z < rn_; z++) rt[z] = readings[z];
      }

      // ---- Observation: first and last tick only (Outcome.java:152-166)
      if (tick == 0 || tick == P.n_ticks - 1) {
        // BehaviorUtils.center: mean over voxels of the mean of the 4 outer corners,
        // summed by the robot's first lane in voxel order (the oracle's order)
        // (BLK: the robot's own slice of the staging area -- the robots of a block pass this point at different times)
        real* st = BLK ? stage_blk + (size_t)rib * wpr * P.smem_per_warp : stage;
        const int sidx = BLK ? rt : lane;
        const int sl0 = BLK ? 0 : lane0;
        SSYNC()
        if (lane_ok) {
          st[sidx * 4 + 0] = (((ox[0] + ox[1]) + ox[2]) + ox[3]) / real(4);
          st[sidx * 4 + 1] = (((oy[0] + oy[1]) + oy[2]) + oy[3]) / real(4);
          st[sidx * 4 + 2] = ar_energy;
          st[sidx * 4 + 3] = ctrl_energy;
        }
        SSYNC()
        if (active && v == 0) {
          double sx = 0, sy = 0, sa = 0, sc = 0;
          for (int q = 0; q < nvox; q++) {
            sx = sx + (double)st[(sl0 + q) * 4 + 0];
            sy = sy + (double)st[(sl0 + q) * 4 + 1];
            sa += (double)st[(sl0 + q) * 4 + 2];
            sc += (double)st[(sl0 + q) * 4 + 3];
          }
          double cx = sx / (double)nvox + P.origin_x, cy = sy / (double)nvox;
          if (tick == 0) {
            first_cx = cx; first_cy = cy; first_are = sa; first_ce = sc;
          }
          if (tick == P.n_ticks - 1) {
            double* R = P.results + (size_t)P.robot_ids[rb] * 11;
            R[0] = first_cx; R[1] = first_cy; R[2] = cx; R[3] = cy;
            R[4] = P.tick_t[0]; R[5] = P.tick_t[tick];
            R[6] = first_are; R[7] = sa; R[8] = first_ce; R[9] = sc;
            int bad = !(isfinite(cx) && isfinite(cy));
            // robot-wide status: gathered below
            int2 tail;
            tail.x = P.n_ticks;
            tail.y = status | (bad ? STATUS_NONFINITE : 0);
            *reinterpret_cast<int2*>(&R[10]) = tail;
          }
        }
        SSYNC()
        if (tick == P.n_ticks - 1) {
          // Robot.boundingBox (Robot.java:116-120) = the largest box of the voxels' outer corners
          // (Voxel.java:481-495): what a later stage of a task is placed by (DevoLocomotion.java)
          if (lane_ok) {
            st[sidx * 4 + 0] = r_min(r_min(ox[0], ox[1]), r_min(ox[2], ox[3]));
            st[sidx * 4 + 1] = r_min(r_min(oy[0], oy[1]), r_min(oy[2], oy[3]));
            st[sidx * 4 + 2] = r_max(r_max(ox[0], ox[1]), r_max(ox[2], ox[3]));
            st[sidx * 4 + 3] = r_max(r_max(oy[0], oy[1]), r_max(oy[2], oy[3]));
          }
          SSYNC()
          if (active && v == 0) {
            real bx0 = st[sl0 * 4 + 0], by0 = st[sl0 * 4 + 1], bx1 = st[sl0 * 4 + 2], by1 = st[sl0 * 4 + 3];
            for (int q = 1; q < nvox; q++) {
              bx0 = r_min(bx0, st[(sl0 + q) * 4 + 0]); by0 = r_min(by0, st[(sl0 + q) * 4 + 1]);
              bx1 = r_max(bx1, st[(sl0 + q) * 4 + 2]); by1 = r_max(by1, st[(sl0 + q) * 4 + 3]);
            }
            double* B = P.final_bbox + (size_t)P.robot_ids[rb] * 4;
            B[0] = (double)bx0 + P.origin_x; B[1] = (double)by0; B[2] = (double)bx1 + P.origin_x; B[3] = (double)by1;
          }
          SSYNC()
        }
      }
    }
    // robot-wide OR of the per-voxel status bits into the record
    if (active && status) atomicOr(reinterpret_cast<int*>(P.results + (size_t)P.robot_ids[rb] * 11 + 10) + 1, status);
  }
}

}  // namespace vsr
