/*
 * vsr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT.  See vsr_oracle.h.
 *
 * fp64, single-robot-at-a-time CPU restatement of the hot path of 2D-VSR-Sim:
 *
 *   Locomotion.apply            tasks/locomotion/Locomotion.java:168-207
 *   AbstractTask.updateWorld    tasks/AbstractTask.java:41-64
 *   Voxel.assemble/act/applyForce/getters  core/objects/Voxel.java:197-203,224-237,239-479,508-556
 *   Robot.assemble/act          core/objects/Robot.java:66-114
 *   Ground / yAt                core/objects/Ground.java:43-82,104-111
 *   sensors                     core/sensors/{Touch,AreaRatio,Velocity,Angle,Constant,AppliedForce,
 *                                             Malfunction,Crumpling,ControlPower,TimeFunction,Lidar,
 *                                             AggregatorSensor,Average,Trend,SoftNormalization,
 *                                             Normalization,Noisy}.java
 *   controllers                 core/controllers/{PhaseSin,TimeFunctions,CentralizedSensing,
 *                                                 DistributedSensing[NonDirectional],
 *                                                 MultiLayerPerceptron}.java
 *   Outcome                     tasks/locomotion/Outcome.java:42-58,126-142,152-200
 *
 * and, for World.step(1), the published algorithm of org.dyn4j:dyn4j:4.2.1
 * (pom.xml:31-35; Box2D lineage; NOT in /root/reference, NOT runnable here):
 * semi-implicit Euler, soft DistanceJoint, rigid WeldJoint, SAT/clipping manifolds,
 * sequential impulses with a 2-point block solver, Baumgarte position pass; contacts of a mass with
 * the ground and with the other masses of its robot (Voxel.java:178-185,474).
 * PARITY UNPINNED for that part: the one physics vector the reference ships (assets/frames.png, the
 * robot centre at 5 ticks of Example.main) is not reproduced (tests/test_oracle_cpu.py, xfail).
 *
 * Each numbered dyn4j item of SURVEY.md Appendix C sits behind one function so it
 * can be corrected if a dyn4j 4.2.1 run ever becomes available.
 *
 * the CCD / time-of-impact pass runs when vsr_settings.continuous_detection is set (the default, as
 * dyn4j's ContinuousDetectionMode.ALL; the kernel has the same pass) or when vsr_oracle_opts.ccd forces it; at-rest
 * sleeping sits behind vsr_oracle_opts.sleeping (off by default; the kernel has none: actuated robots are woken every
 * tick by setRestDistance).
 * DistanceJoint limits (a finite passive range, Voxel.java:66,264-267,331-338) are restated from the published
 * Box2D 2.4 joint that dyn4j 4.2 ported.
 */
#include "vsr_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.14159265358979323846
#define TWO_PI (2.0 * PI)
#define EPS_E 2.220446049250313e-16 /* dyn4j Epsilon.E = machine epsilon of double */
/* Exact geometric ties (a mass lying perfectly flat on a flat segment) are decided by rounding
   in dyn4j's GJK/EPA; here they are decided by rule, with this margin, identically in the
   oracle and in the CUDA kernel. */
#define TIE_EPS 1e-12

static __thread char g_err[512];
const char* vsr_oracle_last_error(void) { return g_err; }
static int fail(int code, const char* msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return code;
}

void vsr_oracle_default_opts(vsr_oracle_opts* o) {
  memset(o, 0, sizeof *o);
  o->order = VSR_ORACLE_ORDER_CANONICAL;
  o->n_threads = 1;
}

/* ------------------------------------------------------------------ ticks */
/* Locomotion.java:195-203 + AbstractTask.java:48: t = 0; while (t < finalT) t = t + dT. */
static int tick_times_from(double t_start, double final_t, double dt, double* t, int n) {
  double tt = t_start;
  int k = 0;
  while (tt < final_t) {
    tt = tt + dt;
    if (t && k < n) t[k] = tt;
    k++;
  }
  return k;
}
int vsr_oracle_tick_times(double final_t, double dt, double* t, int n) {
  double tt = 0.0;
  int k = 0;
  while (tt < final_t) {
    tt = tt + dt;
    if (t && k < n) t[k] = tt;
    k++;
  }
  return k;
}

/* AggregatorSensor.java:56-66: put(t); while (firstKey < t - interval) remove(firstKey). */
int vsr_oracle_window_first(double final_t, double dt, double interval, int32_t* first, int n) {
  int nt = vsr_oracle_tick_times(final_t, dt, NULL, 0);
  double* t = (double*)malloc(sizeof(double) * (size_t)(nt > 0 ? nt : 1));
  vsr_oracle_tick_times(final_t, dt, t, nt);
  int f = 0;
  for (int k = 0; k < nt && k < n; k++) {
    while (t[f] < t[k] - interval) f++;
    first[k] = f;
  }
  free(t);
  return nt;
}

/* ------------------------------------------------------------------ MLP */
/* MultiLayerPerceptron.java:89-95 */
static double act_fn(int a, double x) {
  switch (a) {
    case VSR_ACT_RELU: return (x < 0) ? 0.0 : x;
    case VSR_ACT_SIGMOID: return 1.0 / (1.0 + exp(-x));
    case VSR_ACT_SIN: return sin(x);
    case VSR_ACT_TANH: return tanh(x);
    case VSR_ACT_SIGN: return (x > 0) ? 1.0 : ((x < 0) ? -1.0 : x); /* Math.signum */
    default: return x;
  }
}

/* MultiLayerPerceptron.java:168-189: activation on the raw inputs too (:177);
   weights[layer][neuron][0] is the bias (:181); flat order :139-151. */
int vsr_oracle_mlp_apply(int activation, const int32_t* neurons, int n_layers,
                         const double* w, const double* input, double* output) {
  int maxn = 0;
  for (int i = 0; i < n_layers; i++)
    if (neurons[i] > maxn) maxn = neurons[i];
  double* a = (double*)malloc(sizeof(double) * (size_t)maxn * 2);
  double* b = a + maxn;
  for (int k = 0; k < neurons[0]; k++) a[k] = act_fn(activation, input[k]);
  size_t c = 0;
  for (int i = 1; i < n_layers; i++) {
    for (int j = 0; j < neurons[i]; j++) {
      double sum = w[c++]; /* bias */
      for (int k = 1; k < neurons[i - 1] + 1; k++) sum = sum + a[k - 1] * w[c++];
      b[j] = act_fn(activation, sum);
    }
    double* t = a;
    a = b;
    b = t;
  }
  for (int j = 0; j < neurons[n_layers - 1]; j++) output[j] = a[j];
  free(a < b ? a : b);
  return 0;
}

/* ------------------------------------------------------------------ model */
typedef struct {
  double x, y, th;     /* world centre + rotation angle (dyn4j keeps cos/sin; same map) */
  double vx, vy, w;
  double invM, invI;
} OBody;

typedef struct {
  int b1, b2;
  double a1x, a1y, a2x, a2y; /* body-local anchors */
  double rmin, rrest, rmax;  /* active SpringRange, Voxel.java:287-301 */
  double rest;               /* current rest distance (applyForce) */
  double imp;                /* accumulated impulse (warm start) */
  double nx, ny, cr1, cr2, gamma, bias, soft;
  int full; /* index 0..17 in Voxel.allSpringJoints with all four scaffoldings */
  /* passive-range limits ("rope", Voxel.java:331-338,370-377,433-440,460-467): DistanceJoint lower / upper limit */
  int lo_on, hi_on;
  double lo, hi;
  double lo_imp, hi_imp; /* accumulated limit impulses (warm start) */
  double mass, len;      /* rigid effective mass 1/K and the current length, from initializeConstraints */
  int rigid;             /* frequency 0 (BreakableVoxel.updateStructureMalfunctionType): a rigid distance joint */
} OSpring;

typedef struct {
  int b1, b2, horizontal;
  double a1x, a1y, a2x, a2y, ref;
  double ix, iy, iz;
  double r1x, r1y, r2x, r2y;
  double K[9];
} OWeld;

typedef struct {
  double px, py, depth;
  int id;
  double jn, jt;
  double r1x, r1y, r2x, r2y; /* world offsets from the body centres at initialise time */
  double l1x, l1y, l2x, l2y; /* the same in body-local frames (position solver) */
  double massN, massT, vb;
} OCPoint;

typedef struct {
  int b1;  /* dynamic mass (convex1) */
  int b2;  /* other mass for self collision, or -1 = ground */
  int seg; /* ground segment when b2 == -1 */
  int n;
  double nx, ny; /* manifold normal, from body1 to body2 */
  double mu, e;
  OCPoint p[2];
  int block;
  double K00, K01, K11, iK00, iK01, iK11;
} OContact;

/* BreakableVoxel.java: the parameters (vsr_breakable) and the transient state of one voxel */
typedef struct {
  int enabled;
  unsigned mal[3];
  double thr[3];
  double restore;
  unsigned long long seed; /* java.util.Random state (Random(randomSeed), :213) */
  int state[3];            /* VSR_MALFUNCTION_* per VSR_COMPONENT_* */
  int struct_frozen;       /* what updateStructureMalfunctionType (:244-258) last applied to the springs */
  double cnt[3];           /* triggerCounters per VSR_TRIGGER_* */
  double lastT, lastBreakT, lastCE, lastARE;
  int have_snapshot;       /* sensorReadings != null */
} OBreak;

typedef struct {
  int nv, nb, ns, nw;
  int next_full;
  /* at-rest detection (vsr_oracle_opts.sleeping): the robot is one island */
  int asleep;
  double* rest_time; /* [nb] Body.atRestTime */
  const vsr_material* mat; /* the robot's material: its shape class' own (vsr_shape.material) or the batch's */
  double half;             /* half the side of a mass */
  OBreak* brk;     /* [nv] or NULL: no BreakableVoxel in this robot */
  double* snap;    /* [n_readings] BreakableVoxel.sensorReadings of every voxel */
  double* eff;     /* [n_readings] what Voxel.getSensorReadings returns at this control step */
  double* rd_min;  /* [n_readings] final domain of every reading (Sensor.getDomains) */
  double* rd_max;
  int* vox_gx;
  int* vox_gy;
  OBody* bodies;
  OSpring* springs;
  OWeld* welds;
  int* jorder; /* joint order: >=0 spring index, <0 -> weld index ~j */
  int nj;
  /* VSR_ORACLE_ORDER_DFS: per-body joint lists in world insertion order (ConstraintGraphNode.joints) and the
     island's contact order of the step */
  int* adj_start; /* [nb + 1] */
  int* adj;       /* joints of each body, encoded like jorder */
  int* corder;    /* contact order of this step (indices into contacts), or NULL = detection order */
  int cap_corder;
  OContact* contacts;
  int ncontacts, cap_contacts;
  OContact* old;
  int nold, cap_old;
  /* per voxel */
  double* last_force;
  double* ar_energy;
  double* ctrl_energy;
  unsigned char* touch;
  /* sensors */
  int n_readings;
  double* readings;
  /* aggregator histories: per sensor instance, ticks x dim */
  int n_sens;
  const vsr_sensor** sens;
  int* sens_voxel;
  int* sens_off;    /* offset in readings */
  double** hist;    /* [n_sens] -> [ticks][dim] */
  int* hist_first;  /* index of the oldest kept tick */
  /* Noisy: one java.util.Random per sensor instance (Noisy.java:49) */
  unsigned long long* rng_seed; /* [n_sens] */
  double* rng_spare;            /* [n_sens] nextNextGaussian */
  unsigned char* rng_have;      /* [n_sens] haveNextNextGaussian */
  /* MLP */
  int n_layers;
  int32_t* neurons;
  /* DistributedSensing */
  int* vox_read_off; /* [nv + 1] first reading of each voxel */
  int* vox_at;       /* [w * h] voxel index at grid cell, or -1 */
  int grid_w, grid_h;
  double* last_sig;  /* [nv][4 * signals] lastSignalsGrid */
  double* cur_sig;   /* [nv][4 * signals] currentSignalsGrid */
} ORobot;

typedef struct {
  const vsr_batch_desc* d;
  vsr_oracle_opts o;
  const vsr_oracle_probe* probe;
  vsr_result* out;
  int nticks;
  double* ticks;
  double half; /* half mass side */
  double base_y;
  volatile size_t next;
  pthread_mutex_t mu;
  int status;
} ORun;

static int shape_nvox(const vsr_shape* s) {
  int n = 0;
  for (int i = 0; i < s->w * s->h; i++) n += s->mask[i] ? 1 : 0;
  return n;
}
static int sensor_dim(const vsr_sensor* s) {
  if (s->base == VSR_SENSOR_VELOCITY) return ((s->axes & VSR_AXIS_X) ? 1 : 0) + ((s->axes & VSR_AXIS_Y) ? 1 : 0);
  if (s->base == VSR_SENSOR_LIDAR) return s->axes;
  return 1;
}

/* Ground.yAt, Ground.java:104-111 */
static double ground_y_at(const vsr_batch_desc* d, double x) {
  for (int i = 1; i < d->terrain_n; i++) {
    if (d->terrain_xs[i - 1] <= x && x <= d->terrain_xs[i])
      return (x - d->terrain_xs[i - 1]) * (d->terrain_ys[i] - d->terrain_ys[i - 1]) /
                 (d->terrain_xs[i] - d->terrain_xs[i - 1]) +
             d->terrain_ys[i - 1];
  }
  return NAN;
}

static void add_spring(ORobot* r, int b1, int b2, double a1x, double a1y, double a2x, double a2y,
                       double rmin, double rrest, double rmax, double pmin, double pmax) {
  OSpring* s = &r->springs[r->ns++];
  memset(s, 0, sizeof *s);
  /* joint.setLowerLimit / setUpperLimit only for a finite passive bound (Voxel.java:331-338) */
  s->lo_on = pmin > -INFINITY;
  s->hi_on = pmax < INFINITY;
  s->lo = pmin;
  s->hi = pmax;
  s->b1 = b1;
  s->b2 = b2;
  s->a1x = a1x;
  s->a1y = a1y;
  s->a2x = a2x;
  s->a2y = a2y;
  s->rmin = rmin;
  s->rrest = rrest;
  s->rmax = rmax;
  s->rest = rrest; /* Voxel.java:473 */
  s->full = r->next_full++;
}

/* Robot.join, Robot.java:66-71: anchor = body1's centre (sic). */
static void add_weld(ORobot* r, int b1, int b2, int horizontal) {
  OWeld* w = &r->welds[r->nw++];
  memset(w, 0, sizeof *w);
  w->b1 = b1;
  w->b2 = b2;
  w->horizontal = horizontal;
  OBody* B1 = &r->bodies[b1];
  OBody* B2 = &r->bodies[b2];
  /* localAnchor_i = body_i.getLocalPoint(anchor); bodies are unrotated at assembly */
  w->a1x = 0.0;
  w->a1y = 0.0;
  w->a2x = B1->x - B2->x;
  w->a2y = B1->y - B2->y;
  w->ref = B1->th - B2->th;
}

static int robot_build(ORun* R, ORobot* r, size_t ridx) {
  const vsr_batch_desc* d = R->d;
  const vsr_shape* sh = &d->shapes[d->robot_shape ? d->robot_shape[ridx] : 0];
  const vsr_material* m = sh->material ? sh->material : &d->material;
  memset(r, 0, sizeof *r);
  r->mat = m;
  r->half = m->side_length * m->mass_side_length_ratio / 2.0;
  int nv = shape_nvox(sh);
  r->nv = nv;
  r->nb = 4 * nv;
  r->vox_gx = (int*)malloc(sizeof(int) * nv);
  r->vox_gy = (int*)malloc(sizeof(int) * nv);
  int* grid = (int*)malloc(sizeof(int) * sh->w * sh->h);
  int v = 0;
  for (int i = 0; i < sh->w * sh->h; i++) {
    grid[i] = -1;
    if (sh->mask[i]) {
      r->vox_gx[v] = i % sh->w;
      r->vox_gy[v] = i / sh->w;
      grid[i] = v++;
    }
  }
  r->bodies = (OBody*)calloc((size_t)r->nb, sizeof(OBody));
  r->springs = (OSpring*)calloc((size_t)nv * 18, sizeof(OSpring));
  r->welds = (OWeld*)calloc((size_t)nv * 4 + 4, sizeof(OWeld));

  /* --- Voxel.assemble, Voxel.java:239-479 --- */
  double L = m->side_length;
  double ms = L * m->mass_side_length_ratio;                /* :241 */
  double density = (m->mass / 4) / (ms * ms);               /* :242 */
  double bmass = density * ms * ms;                         /* Rectangle.createMass (dyn4j) */
  double binertia = bmass * (ms * ms + ms * ms) / 12.0;
  double off = L / 2.0 - ms / 2.0;                          /* :253-256 */
  double h = ms / 2.0;
  static const double sx[4] = {-1, +1, +1, -1}; /* NW NE SE SW */
  static const double sy[4] = {+1, +1, -1, -1};
  /* ranges :262-301: the active ranges move the rest distances, the passive ones are the joints' limits */
  double amin = sqrt(L * L * m->area_ratio_active_min), amax = sqrt(L * L * m->area_ratio_active_max);
  double pmin = sqrt(L * L * m->area_ratio_passive_min), pmax = sqrt(L * L * m->area_ratio_passive_max); /* NaN / inf when infinite */
  int plo = isfinite(m->area_ratio_passive_min), phi = isfinite(m->area_ratio_passive_max);
  double pp_min = plo ? pmin - 2.0 * ms : -INFINITY, pp_max = phi ? pmax - 2.0 * ms : INFINITY;
  double pc_min = plo ? sqrt(ms * ms + pp_min * pp_min) : -INFINITY, pc_max = phi ? sqrt(ms * ms + pp_max * pp_max) : INFINITY;
  double pd_min = plo ? (pmin - ms) * sqrt(2.0) : -INFINITY, pd_max = phi ? (pmax - ms) * sqrt(2.0) : INFINITY;
  double sp_min = amin - 2.0 * ms, sp_rest = L - 2.0 * ms, sp_max = amax - 2.0 * ms;
  double sc_min = sqrt(ms * ms + sp_min * sp_min), sc_rest = sqrt(ms * ms + sp_rest * sp_rest),
         sc_max = sqrt(ms * ms + sp_max * sp_max);
  double cc_min = (amin - ms) * sqrt(2.0), cc_rest = (L - ms) * sqrt(2.0), cc_max = (amax - ms) * sqrt(2.0);
  for (v = 0; v < nv; v++) {
    int b0 = 4 * v;
    for (int k = 0; k < 4; k++) {
      OBody* B = &r->bodies[b0 + k];
      /* Voxel.java:253-256 then Robot.java:99 translate by (gx*L, gy*L) */
      B->x = sx[k] * off + (double)r->vox_gx[v] * L;
      B->y = sy[k] * off + (double)r->vox_gy[v] * L;
      B->th = 0.0;
      B->invM = 1.0 / bmass;
      B->invI = 1.0 / binertia;
    }
    uint32_t sc = m->spring_scaffoldings;
    if (sc & VSR_SCAFFOLD_SIDE_INTERNAL) { /* :303-341 */
      r->next_full = 0;
      add_spring(r, b0 + 0, b0 + 1, +h, -h, -h, -h, sp_min, sp_rest, sp_max, pp_min, pp_max);
      add_spring(r, b0 + 1, b0 + 2, -h, -h, -h, +h, sp_min, sp_rest, sp_max, pp_min, pp_max);
      add_spring(r, b0 + 2, b0 + 3, -h, +h, +h, +h, sp_min, sp_rest, sp_max, pp_min, pp_max);
      add_spring(r, b0 + 3, b0 + 0, +h, +h, +h, -h, sp_min, sp_rest, sp_max, pp_min, pp_max);
    }
    if (sc & VSR_SCAFFOLD_SIDE_EXTERNAL) { /* :342-380 */
      r->next_full = 4;
      add_spring(r, b0 + 0, b0 + 1, +h, +h, -h, +h, sp_min, sp_rest, sp_max, pp_min, pp_max);
      add_spring(r, b0 + 1, b0 + 2, +h, -h, +h, +h, sp_min, sp_rest, sp_max, pp_min, pp_max);
      add_spring(r, b0 + 2, b0 + 3, -h, -h, +h, -h, sp_min, sp_rest, sp_max, pp_min, pp_max);
      add_spring(r, b0 + 3, b0 + 0, -h, +h, -h, -h, sp_min, sp_rest, sp_max, pp_min, pp_max);
    }
    if (sc & VSR_SCAFFOLD_SIDE_CROSS) { /* :381-443 */
      r->next_full = 8;
      add_spring(r, b0 + 0, b0 + 1, +h, +h, -h, -h, sc_min, sc_rest, sc_max, pc_min, pc_max);
      add_spring(r, b0 + 0, b0 + 1, +h, -h, -h, +h, sc_min, sc_rest, sc_max, pc_min, pc_max);
      add_spring(r, b0 + 1, b0 + 2, +h, -h, -h, +h, sc_min, sc_rest, sc_max, pc_min, pc_max);
      add_spring(r, b0 + 1, b0 + 2, -h, -h, +h, +h, sc_min, sc_rest, sc_max, pc_min, pc_max);
      add_spring(r, b0 + 2, b0 + 3, -h, +h, +h, -h, sc_min, sc_rest, sc_max, pc_min, pc_max);
      add_spring(r, b0 + 2, b0 + 3, -h, -h, +h, +h, sc_min, sc_rest, sc_max, pc_min, pc_max);
      add_spring(r, b0 + 3, b0 + 0, -h, +h, +h, -h, sc_min, sc_rest, sc_max, pc_min, pc_max);
      add_spring(r, b0 + 3, b0 + 0, +h, +h, -h, -h, sc_min, sc_rest, sc_max, pc_min, pc_max);
    }
    if (sc & VSR_SCAFFOLD_CENTRAL_CROSS) { /* :444-470 */
      r->next_full = 16;
      add_spring(r, b0 + 0, b0 + 2, 0, 0, 0, 0, cc_min, cc_rest, cc_max, pd_min, pd_max);
      add_spring(r, b0 + 1, b0 + 3, 0, 0, 0, 0, cc_min, cc_rest, cc_max, pd_min, pd_max);
    }
  }
  /* --- Robot.assemble, Robot.java:91-114: gx outer, gy inner; left pair then lower pair --- */
  for (int gx = 0; gx < sh->w; gx++) {
    for (int gy = 0; gy < sh->h; gy++) {
      int vv = grid[gy * sh->w + gx];
      if (vv < 0) continue;
      if (gx > 0 && grid[gy * sh->w + gx - 1] >= 0) {
        int a = grid[gy * sh->w + gx - 1];
        add_weld(r, 4 * vv + 0, 4 * a + 1, 1);
        add_weld(r, 4 * vv + 3, 4 * a + 2, 1);
      }
      if (gy > 0 && grid[(gy - 1) * sh->w + gx] >= 0) {
        int a = grid[(gy - 1) * sh->w + gx];
        add_weld(r, 4 * vv + 3, 4 * a + 0, 0);
        add_weld(r, 4 * vv + 2, 4 * a + 1, 0);
      }
    }
  }
  free(grid);

  /* --- placement, Locomotion.java:180-187 --- */
  /* robot.boundingBox(): voxel bbox = outer corners (Voxel.java:481-495), unrotated here */
  double minx = INFINITY;
  for (v = 0; v < nv; v++) {
    double x0 = r->bodies[4 * v + 0].x - h; /* NW body's NW corner */
    double x3 = r->bodies[4 * v + 3].x - h;
    if (x0 < minx) minx = x0;
    if (x3 < minx) minx = x3;
  }
  double dx = (d->robot_placement ? d->robot_placement[ridx] : d->initial_placement) - minx;
  for (int b = 0; b < r->nb; b++) r->bodies[b].x += dx;
  double min_gap = INFINITY;
  for (v = 0; v < nv; v++) {
    double cx = 0, miny = INFINITY;
    for (int k = 0; k < 4; k++) {
      cx += r->bodies[4 * v + k].x;
      double yy = r->bodies[4 * v + k].y - h;
      if (yy < miny) miny = yy;
    }
    cx /= 4.0; /* Voxel.center, Voxel.java:497-506 */
    double gap = miny - ground_y_at(d, cx);
    if (gap < min_gap) min_gap = gap;
  }
  if (nv == 0) min_gap = 0.0;
  double dy = 1.0 /* INITIAL_PLACEMENT_Y_GAP */ - min_gap;
  for (int b = 0; b < r->nb; b++) r->bodies[b].y += dy;

  /* --- joint order --- */
  r->nj = r->ns + r->nw;
  r->jorder = (int*)malloc(sizeof(int) * (size_t)(r->nj > 0 ? r->nj : 1));
  int c = 0;
  if (R->o.order == VSR_ORACLE_ORDER_CANONICAL) {
    /* Inside a voxel the canonical sweep alternates the perfect matchings of K4 ({01,23}, {12,30},
       {02,13}: SURVEY.md Appendix B), so that consecutive solves touch disjoint masses -- the order
       the CUDA kernel's unrolled block uses (two independent dependency chains).  Springs of
       different voxels share no mass, so the voxel order is immaterial. */
    static const int PERM[18] = {0, 2, 1, 3, 4, 6, 5, 7, 8, 12, 9, 13, 10, 14, 11, 15, 16, 17};
    int per = r->nv > 0 ? r->ns / r->nv : 0;
    for (int v = 0; v < r->nv; v++)
      for (int i = 0; i < 18; i++)
        for (int k = 0; k < per; k++)
          if (r->springs[v * per + k].full == PERM[i]) r->jorder[c++] = v * per + k;
  } else {
    for (int s = 0; s < r->ns; s++) r->jorder[c++] = s;
  }
  if (R->o.order == VSR_ORACLE_ORDER_CANONICAL) {
    for (int w = 0; w < r->nw; w++)
      if (r->welds[w].horizontal) r->jorder[c++] = ~w;
    for (int w = 0; w < r->nw; w++)
      if (!r->welds[w].horizontal) r->jorder[c++] = ~w;
  } else {
    for (int w = 0; w < r->nw; w++) r->jorder[c++] = ~w;
    if (R->o.order == VSR_ORACLE_ORDER_SHUFFLED) {
      uint64_t st = 0x9E3779B97F4A7C15ull ^ ((uint64_t)R->o.seed * 0xD1342543DE82EF95ull) ^ (uint64_t)ridx;
      for (int i = r->nj - 1; i > 0; i--) {
        st = st * 6364136223846793005ull + 1442695040888963407ull;
        int j = (int)((st >> 33) % (uint64_t)(i + 1));
        int t = r->jorder[i];
        r->jorder[i] = r->jorder[j];
        r->jorder[j] = t;
      }
    }
  }

  /* per-body joint lists, in the order world.addJoint saw the joints (Voxel.java:216-221: a voxel's 18 springs;
     Robot.java:81-88: then all welds): ConstraintGraphNode.joints of dyn4j */
  r->adj_start = (int*)calloc((size_t)r->nb + 1, sizeof(int));
  r->adj = (int*)malloc(sizeof(int) * (size_t)(2 * r->nj > 0 ? 2 * r->nj : 1));
  r->corder = NULL;
  r->cap_corder = 0;
  {
    for (int si = 0; si < r->ns; si++) { r->adj_start[r->springs[si].b1 + 1]++; r->adj_start[r->springs[si].b2 + 1]++; }
    for (int w = 0; w < r->nw; w++) { r->adj_start[r->welds[w].b1 + 1]++; r->adj_start[r->welds[w].b2 + 1]++; }
    for (int b = 0; b < r->nb; b++) r->adj_start[b + 1] += r->adj_start[b];
    int* fill = (int*)calloc((size_t)r->nb + 1, sizeof(int));
    for (int si = 0; si < r->ns; si++) {
      int b1 = r->springs[si].b1, b2 = r->springs[si].b2;
      r->adj[r->adj_start[b1] + fill[b1]++] = si;
      r->adj[r->adj_start[b2] + fill[b2]++] = si;
    }
    for (int w = 0; w < r->nw; w++) {
      int b1 = r->welds[w].b1, b2 = r->welds[w].b2;
      r->adj[r->adj_start[b1] + fill[b1]++] = ~w;
      r->adj[r->adj_start[b2] + fill[b2]++] = ~w;
    }
    free(fill);
  }

  /* --- per voxel state: Voxel.reset, Voxel.java:640-651 --- */
  r->last_force = (double*)calloc((size_t)nv, sizeof(double));
  r->ar_energy = (double*)calloc((size_t)nv, sizeof(double));
  r->ctrl_energy = (double*)calloc((size_t)nv, sizeof(double));
  r->touch = (unsigned char*)calloc((size_t)nv, 1);

  /* --- sensors --- */
  int ns = sh->sensor_offsets ? sh->sensor_offsets[nv] : 0;
  r->n_sens = ns;
  r->sens = (const vsr_sensor**)malloc(sizeof(void*) * (size_t)(ns + 1));
  r->sens_voxel = (int*)malloc(sizeof(int) * (size_t)(ns + 1));
  r->sens_off = (int*)malloc(sizeof(int) * (size_t)(ns + 1));
  r->hist = (double**)calloc((size_t)(ns + 1), sizeof(double*));
  r->hist_first = (int*)calloc((size_t)(ns + 1), sizeof(int));
  int off_r = 0, si = 0;
  for (v = 0; v < nv && ns > 0; v++) {
    for (int k = sh->sensor_offsets[v]; k < sh->sensor_offsets[v + 1]; k++) {
      r->sens[si] = &sh->sensors[k];
      r->sens_voxel[si] = v;
      r->sens_off[si] = off_r;
      int dim = sensor_dim(&sh->sensors[k]);
      if (sh->sensors[k].aggregator != VSR_AGG_NONE)
        r->hist[si] = (double*)malloc(sizeof(double) * (size_t)R->nticks * (size_t)dim);
      off_r += dim;
      si++;
    }
  }
  r->n_readings = off_r;
  r->readings = (double*)calloc((size_t)(off_r + 1), sizeof(double));
  r->rest_time = (double*)calloc((size_t)r->nb + 1, sizeof(double));
  r->brk = NULL;
  r->snap = r->eff = r->rd_min = r->rd_max = NULL;
  if (sh->breakable) {
    /* BreakableVoxel constructor + reset() (:97-104,216-229): counters 0, every component NONE, springs soft */
    r->brk = (OBreak*)calloc((size_t)nv + 1, sizeof(OBreak));
    r->snap = (double*)calloc((size_t)(off_r + 1), sizeof(double));
    r->eff = (double*)calloc((size_t)(off_r + 1), sizeof(double));
    r->rd_min = (double*)calloc((size_t)(off_r + 1), sizeof(double));
    r->rd_max = (double*)calloc((size_t)(off_r + 1), sizeof(double));
    for (int vv = 0; vv < nv; vv++) {
      const vsr_breakable* q = &sh->breakable[vv];
      OBreak* B = &r->brk[vv];
      B->enabled = q->enabled != 0;
      for (int z = 0; z < 3; z++) {
        B->mal[z] = q->malfunctions[z];
        B->thr[z] = q->thresholds[z];
      }
      B->restore = q->restore_time;
      B->seed = ((unsigned long long)q->random_seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
    }
  }
  r->rng_seed = (unsigned long long*)calloc((size_t)ns + 1, sizeof(unsigned long long));
  r->rng_spare = (double*)calloc((size_t)ns + 1, sizeof(double));
  r->rng_have = (unsigned char*)calloc((size_t)ns + 1, 1);
  for (int q = 0; q < r->n_sens; q++)
    r->rng_seed[q] = ((unsigned long long)r->sens[q]->noise_seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);

  /* --- DistributedSensing: per-voxel reading ranges, grid lookup, signal grids (reset() = zeros) --- */
  r->vox_read_off = (int*)calloc((size_t)nv + 1, sizeof(int));
  {
    int acc = 0, si2 = 0;
    for (int v = 0; v < nv; v++) {
      r->vox_read_off[v] = acc;
      if (sh->sensor_offsets)
        for (int k = sh->sensor_offsets[v]; k < sh->sensor_offsets[v + 1]; k++, si2++) acc += sensor_dim(&sh->sensors[k]);
    }
    r->vox_read_off[nv] = acc;
  }
  r->grid_w = sh->w;
  r->grid_h = sh->h;
  r->vox_at = (int*)malloc(sizeof(int) * (size_t)(sh->w * sh->h));
  for (int i = 0; i < sh->w * sh->h; i++) r->vox_at[i] = -1;
  for (int v = 0; v < nv; v++) r->vox_at[r->vox_gy[v] * sh->w + r->vox_gx[v]] = v;
  r->last_sig = r->cur_sig = NULL;
  if (d->controller_kind >= VSR_CTRL_DISTRIBUTED) {
    size_t ns4 = (size_t)nv * 4 * (size_t)(d->distributed_signals > 0 ? d->distributed_signals : 1);
    r->last_sig = (double*)calloc(ns4, sizeof(double));
    r->cur_sig = (double*)calloc(ns4, sizeof(double));
    r->n_layers = 2 + sh->n_hidden;
    r->neurons = (int32_t*)malloc(sizeof(int32_t) * (size_t)r->n_layers);
    for (int i = 0; i < sh->n_hidden; i++) r->neurons[1 + i] = sh->hidden[i];
  }
  /* --- MLP topology: CentralizedSensing.nOfInputs/nOfOutputs, :68-83 --- */
  if (d->controller_kind == VSR_CTRL_MLP) {
    r->n_layers = 2 + sh->n_hidden;
    r->neurons = (int32_t*)malloc(sizeof(int32_t) * (size_t)r->n_layers);
    r->neurons[0] = r->n_readings;
    for (int i = 0; i < sh->n_hidden; i++) r->neurons[1 + i] = sh->hidden[i];
    r->neurons[r->n_layers - 1] = nv;
  }
  return 0;
}

static void robot_free(ORobot* r) {
  free(r->rng_seed);
  free(r->rng_spare);
  free(r->rng_have);
  free(r->vox_read_off);
  free(r->vox_at);
  free(r->last_sig);
  free(r->cur_sig);
  free(r->vox_gx);
  free(r->vox_gy);
  free(r->bodies);
  free(r->springs);
  free(r->welds);
  free(r->jorder);
  free(r->adj_start);
  free(r->adj);
  free(r->corder);
  free(r->contacts);
  free(r->old);
  free(r->last_force);
  free(r->ar_energy);
  free(r->ctrl_energy);
  free(r->touch);
  free(r->readings);
  free(r->rest_time);
  free(r->brk);
  free(r->snap);
  free(r->eff);
  free(r->rd_min);
  free(r->rd_max);
  for (int i = 0; i < r->n_sens; i++) free(r->hist[i]);
  free(r->hist);
  free(r->hist_first);
  free((void*)r->sens);
  free(r->sens_voxel);
  free(r->sens_off);
  free(r->neurons);
}

/* ------------------------------------------------------------------ dyn4j C.2 */
/* AbstractPhysicsBody.integrateVelocity (dyn4j): no user forces are ever applied. */
static void integrate_velocity(const vsr_batch_desc* d, const vsr_material* mat, OBody* b, double dt) {
  const vsr_settings* s = &d->settings;
  b->vx += dt * s->gravity_x;
  b->vy += dt * s->gravity_y;
  double lin = 1.0 - dt * mat->mass_linear_damping;
  lin = lin < 0 ? 0 : (lin > 1 ? 1 : lin);
  if (mat->mass_linear_damping != 0.0) {
    b->vx *= lin;
    b->vy *= lin;
  }
  double ang = 1.0 - dt * mat->mass_angular_damping;
  ang = ang < 0 ? 0 : (ang > 1 ? 1 : ang);
  b->w *= ang;
}

/* AbstractPhysicsBody.integratePosition (dyn4j), with maxTranslation/maxRotation caps. */
static void integrate_position(const vsr_settings* s, OBody* b, double dt) {
  double tx = b->vx * dt, ty = b->vy * dt;
  double m2 = tx * tx + ty * ty;
  if (m2 > s->max_translation * s->max_translation) {
    double ratio = s->max_translation / sqrt(m2);
    b->vx *= ratio;
    b->vy *= ratio;
    tx *= ratio;
    ty *= ratio;
  }
  double rot = b->w * dt;
  if (fabs(rot) > s->max_rotation) {
    double ratio = s->max_rotation / fabs(rot);
    b->w *= ratio;
    rot *= ratio;
  }
  b->x += tx;
  b->y += ty;
  b->th += rot;
}

/* ------------------------------------------------------------------ dyn4j C.3 */
/* Which mass turns (frequency, dampingRatio) into (k, d): the single switch point. */
static double spring_mass(const vsr_settings* s, const OBody* B1, const OBody* B2, double eff_mass) {
  if (s->spring_mass_mode == VSR_SPRING_MASS_EFFECTIVE) return eff_mass;
  double m1 = B1->invM > 0 ? 1.0 / B1->invM : 0.0, m2 = B2->invM > 0 ? 1.0 / B2->invM : 0.0;
  if (m1 > 0.0 && m2 > 0.0) return m1 * m2 / (m1 + m2);
  return m1 > 0.0 ? m1 : m2;
}

/* DistanceJoint.initializeConstraints (dyn4j 4.2: the Box2D 2.4 joint), soft (frequency > 0), optional limits. */
static void spring_init(const vsr_batch_desc* d, ORobot* r, OSpring* s, double dt) {
  OBody* B1 = &r->bodies[s->b1];
  OBody* B2 = &r->bodies[s->b2];
  double c1 = cos(B1->th), s1 = sin(B1->th), c2 = cos(B2->th), s2 = sin(B2->th);
  double r1x = c1 * s->a1x - s1 * s->a1y, r1y = s1 * s->a1x + c1 * s->a1y;
  double r2x = c2 * s->a2x - s2 * s->a2y, r2y = s2 * s->a2x + c2 * s->a2y;
  double nx = r1x + B1->x - (r2x + B2->x), ny = r1y + B1->y - (r2y + B2->y);
  double len = sqrt(nx * nx + ny * ny);
  if (len < d->settings.linear_tolerance) {
    nx = 0;
    ny = 0;
  } else {
    nx *= 1.0 / len;
    ny *= 1.0 / len;
  }
  double cr1 = r1x * ny - r1y * nx, cr2 = r2x * ny - r2y * nx;
  double invMass = B1->invM + B1->invI * cr1 * cr1;
  invMass += B2->invM + B2->invI * cr2 * cr2;
  double mass = invMass <= EPS_E ? 0.0 : 1.0 / invMass;
  double x = len - s->rest;
  double w = TWO_PI * r->mat->spring_f;
  double m = spring_mass(&d->settings, B1, B2, mass);
  double dc = 2.0 * m * r->mat->spring_d * w;
  double k = m * w * w;
  double gamma = dt * (dc + dt * k);
  gamma = gamma <= EPS_E ? 0.0 : 1.0 / gamma;
  double bias = x * dt * k * gamma;
  if (s->rigid) { /* BreakableVoxel STRUCTURE FROZEN: setFrequency(0) -> a rigid DistanceJoint (gamma = bias = 0) */
    gamma = 0.0;
    bias = 0.0;
  }
  invMass += gamma;
  s->soft = invMass <= EPS_E ? 0.0 : 1.0 / invMass;
  s->gamma = gamma;
  s->bias = bias;
  s->nx = nx;
  s->ny = ny;
  s->cr1 = cr1;
  s->cr2 = cr2;
  s->mass = mass;
  s->len = len;
  /* warm start (dtRatio = 1): spring + lower - upper along n */
  double wi = s->imp + s->lo_imp - s->hi_imp;
  double jx = nx * wi, jy = ny * wi;
  B1->vx += jx * B1->invM;
  B1->vy += jy * B1->invM;
  B1->w += B1->invI * cr1 * wi;
  B2->vx -= jx * B2->invM;
  B2->vy -= jy * B2->invM;
  B2->w -= B2->invI * cr2 * wi;
}

/* DistanceJoint.solveVelocityConstraints (dyn4j): Jv = n.(v1 + w1 x r1 - v2 - w2 x r2),
   and n.(w x r) = w (r x n). */
static void spring_solve(ORobot* r, OSpring* s) {
  OBody* B1 = &r->bodies[s->b1];
  OBody* B2 = &r->bodies[s->b2];
  double Jv = s->nx * (B1->vx - B2->vx) + s->ny * (B1->vy - B2->vy) + s->cr1 * B1->w - s->cr2 * B2->w;
  double j = -s->soft * (Jv + s->bias + s->gamma * s->imp);
  s->imp += j;
  double jx = s->nx * j, jy = s->ny * j;
  B1->vx += jx * B1->invM;
  B1->vy += jy * B1->invM;
  B1->w += B1->invI * s->cr1 * j;
  B2->vx -= jx * B2->invM;
  B2->vy -= jy * B2->invM;
  B2->w -= B2->invI * s->cr2 * j;
}

/* The limits of DistanceJoint.solveVelocityConstraints (dyn4j 4.2 = b2DistanceJoint of Box2D 2.4, restated: the jar is
   not in the reference tree): after the spring row, a one-sided rigid row per enabled limit,
     lower: C = len - lo, bias = max(0, C) / dt, Cdot = n.(v1 - v2), impulse = -mass (Cdot + bias), accumulated >= 0;
     upper: C = hi - len, bias = max(0, C) / dt, Cdot = n.(v2 - v1), impulse = -mass (Cdot + bias), accumulated >= 0,
            applied along -n.
   (n points from body 2 to body 1, so "len grows" = Cdot > 0.) */
static void spring_limits_solve(ORobot* r, OSpring* s, double inv_dt) {
  OBody* B1 = &r->bodies[s->b1];
  OBody* B2 = &r->bodies[s->b2];
  if (s->lo_on) {
    double C = s->len - s->lo;
    double bias = (C > 0.0 ? C : 0.0) * inv_dt;
    double Jv = s->nx * (B1->vx - B2->vx) + s->ny * (B1->vy - B2->vy) + s->cr1 * B1->w - s->cr2 * B2->w;
    double j = -s->mass * (Jv + bias);
    double nw = s->lo_imp + j;
    if (nw < 0.0) nw = 0.0;
    j = nw - s->lo_imp;
    s->lo_imp = nw;
    double jx = s->nx * j, jy = s->ny * j;
    B1->vx += jx * B1->invM;
    B1->vy += jy * B1->invM;
    B1->w += B1->invI * s->cr1 * j;
    B2->vx -= jx * B2->invM;
    B2->vy -= jy * B2->invM;
    B2->w -= B2->invI * s->cr2 * j;
  }
  if (s->hi_on) {
    double C = s->hi - s->len;
    double bias = (C > 0.0 ? C : 0.0) * inv_dt;
    double Jv = -(s->nx * (B1->vx - B2->vx) + s->ny * (B1->vy - B2->vy) + s->cr1 * B1->w - s->cr2 * B2->w);
    double j = -s->mass * (Jv + bias);
    double nw = s->hi_imp + j;
    if (nw < 0.0) nw = 0.0;
    j = nw - s->hi_imp;
    s->hi_imp = nw;
    j = -j;
    double jx = s->nx * j, jy = s->ny * j;
    B1->vx += jx * B1->invM;
    B1->vy += jy * B1->invM;
    B1->w += B1->invI * s->cr1 * j;
    B2->vx -= jx * B2->invM;
    B2->vy -= jy * B2->invM;
    B2->w -= B2->invI * s->cr2 * j;
  }
}

/* DistanceJoint.solvePositionConstraints with limits: a spring-damper without limits reports "solved"; with limits
   the violated one is corrected rigidly (C clamped to +-maxLinearCorrection, the effective mass of
   initializeConstraints), solved when |C| < linearTolerance. */
static int spring_limits_position(const vsr_settings* S, ORobot* r, OSpring* s) {
  OBody* B1 = &r->bodies[s->b1];
  OBody* B2 = &r->bodies[s->b2];
  double c1 = cos(B1->th), s1 = sin(B1->th), c2 = cos(B2->th), s2 = sin(B2->th);
  double r1x = c1 * s->a1x - s1 * s->a1y, r1y = s1 * s->a1x + c1 * s->a1y;
  double r2x = c2 * s->a2x - s2 * s->a2y, r2y = s2 * s->a2x + c2 * s->a2y;
  double nx = r1x + B1->x - (r2x + B2->x), ny = r1y + B1->y - (r2y + B2->y);
  double len = sqrt(nx * nx + ny * ny);
  if (len > 0.0) { nx /= len; ny /= len; }
  double C;
  if (s->rigid) C = len - s->rest; /* a rigid joint is corrected to its rest distance on both sides */
  else if (s->lo_on && len < s->lo) C = len - s->lo;
  else if (s->hi_on && len > s->hi) C = len - s->hi;
  else return 1;
  if (C > S->max_linear_correction) C = S->max_linear_correction;
  if (C < -S->max_linear_correction) C = -S->max_linear_correction;
  double imp = -s->mass * C;
  double Jx = nx * imp, Jy = ny * imp;
  B1->x += Jx * B1->invM;
  B1->y += Jy * B1->invM;
  B1->th += B1->invI * (r1x * Jy - r1y * Jx);
  B2->x -= Jx * B2->invM;
  B2->y -= Jy * B2->invM;
  B2->th -= B2->invI * (r2x * Jy - r2y * Jx);
  return fabs(C) < S->linear_tolerance;
}

/* ------------------------------------------------------------------ dyn4j C.4 */
static void weld_K(const OBody* B1, const OBody* B2, double r1x, double r1y, double r2x, double r2y, double* K) {
  double m = B1->invM + B2->invM, i1 = B1->invI, i2 = B2->invI;
  K[0] = m + r1y * r1y * i1 + r2y * r2y * i2;
  K[1] = -r1y * r1x * i1 - r2y * r2x * i2;
  K[2] = -r1y * i1 - r2y * i2;
  K[3] = K[1];
  K[4] = m + r1x * r1x * i1 + r2x * r2x * i2;
  K[5] = r1x * i1 + r2x * i2;
  K[6] = K[2];
  K[7] = K[5];
  K[8] = i1 + i2;
}
/* Matrix33.solve33 (dyn4j): Cramer's rule. */
static void solve33(const double* K, double bx, double by, double bz, double* x, double* y, double* z) {
  double det = K[0] * (K[4] * K[8] - K[5] * K[7]) - K[1] * (K[3] * K[8] - K[5] * K[6]) +
               K[2] * (K[3] * K[7] - K[4] * K[6]);
  if (fabs(det) > EPS_E) det = 1.0 / det;
  double m00 = K[4] * K[8] - K[5] * K[7], m01 = K[2] * K[7] - K[1] * K[8], m02 = K[1] * K[5] - K[2] * K[4];
  double m10 = K[5] * K[6] - K[3] * K[8], m11 = K[0] * K[8] - K[2] * K[6], m12 = K[2] * K[3] - K[0] * K[5];
  double m20 = K[3] * K[7] - K[4] * K[6], m21 = K[1] * K[6] - K[0] * K[7], m22 = K[0] * K[4] - K[1] * K[3];
  *x = det * (m00 * bx + m01 * by + m02 * bz);
  *y = det * (m10 * bx + m11 * by + m12 * bz);
  *z = det * (m20 * bx + m21 * by + m22 * bz);
}

/* WeldJoint.initializeConstraints (dyn4j), frequency = 0. */
static void weld_init(ORobot* r, OWeld* w) {
  OBody* B1 = &r->bodies[w->b1];
  OBody* B2 = &r->bodies[w->b2];
  double c1 = cos(B1->th), s1 = sin(B1->th), c2 = cos(B2->th), s2 = sin(B2->th);
  w->r1x = c1 * w->a1x - s1 * w->a1y;
  w->r1y = s1 * w->a1x + c1 * w->a1y;
  w->r2x = c2 * w->a2x - s2 * w->a2y;
  w->r2y = s2 * w->a2x + c2 * w->a2y;
  weld_K(B1, B2, w->r1x, w->r1y, w->r2x, w->r2y, w->K);
  /* warm start */
  B1->vx += w->ix * B1->invM;
  B1->vy += w->iy * B1->invM;
  B1->w += B1->invI * (w->r1x * w->iy - w->r1y * w->ix + w->iz);
  B2->vx -= w->ix * B2->invM;
  B2->vy -= w->iy * B2->invM;
  B2->w -= B2->invI * (w->r2x * w->iy - w->r2y * w->ix + w->iz);
}

/* WeldJoint.solveVelocityConstraints (dyn4j), rigid 3x3. */
static void weld_solve(ORobot* r, OWeld* w) {
  OBody* B1 = &r->bodies[w->b1];
  OBody* B2 = &r->bodies[w->b2];
  /* v + w x r, with w x r = (-w r.y, w r.x) */
  double cx = (B1->vx - B1->w * w->r1y) - (B2->vx - B2->w * w->r2y);
  double cy = (B1->vy + B1->w * w->r1x) - (B2->vy + B2->w * w->r2x);
  double cz = B1->w - B2->w;
  double ix, iy, iz;
  solve33(w->K, -cx, -cy, -cz, &ix, &iy, &iz);
  w->ix += ix;
  w->iy += iy;
  w->iz += iz;
  B1->vx += ix * B1->invM;
  B1->vy += iy * B1->invM;
  B1->w += B1->invI * (w->r1x * iy - w->r1y * ix + iz);
  B2->vx -= ix * B2->invM;
  B2->vy -= iy * B2->invM;
  B2->w -= B2->invI * (w->r2x * iy - w->r2y * ix + iz);
}

static double wrap_pi(double rr) {
  if (rr < -PI) rr += TWO_PI;
  if (rr > PI) rr -= TWO_PI;
  return rr;
}

/* WeldJoint.solvePositionConstraints (dyn4j). */
static int weld_position(const vsr_settings* s, ORobot* r, OWeld* w) {
  OBody* B1 = &r->bodies[w->b1];
  OBody* B2 = &r->bodies[w->b2];
  double c1 = cos(B1->th), s1 = sin(B1->th), c2 = cos(B2->th), s2 = sin(B2->th);
  double r1x = c1 * w->a1x - s1 * w->a1y, r1y = s1 * w->a1x + c1 * w->a1y;
  double r2x = c2 * w->a2x - s2 * w->a2y, r2y = s2 * w->a2x + c2 * w->a2y;
  double Cx = (B1->x + r1x) - (B2->x + r2x), Cy = (B1->y + r1y) - (B2->y + r2y);
  double Cz = wrap_pi(B1->th - B2->th - w->ref);
  double linErr = sqrt(Cx * Cx + Cy * Cy), angErr = fabs(Cz);
  double K[9];
  weld_K(B1, B2, r1x, r1y, r2x, r2y, K);
  double ix, iy, iz;
  solve33(K, -Cx, -Cy, -Cz, &ix, &iy, &iz);
  B1->x += ix * B1->invM;
  B1->y += iy * B1->invM;
  B1->th += B1->invI * (r1x * iy - r1y * ix + iz);
  B2->x -= ix * B2->invM;
  B2->y -= iy * B2->invM;
  B2->th -= B2->invI * (r2x * iy - r2y * ix + iz);
  return linErr <= s->linear_tolerance && angErr <= s->angular_tolerance;
}

/* ------------------------------------------------------------------ dyn4j C.5: collision */
typedef struct {
  int n;
  double vx[4], vy[4]; /* CCW world vertices */
  double nx[4], ny[4]; /* outward normal of edge i -> i+1 */
} OPoly;

static void poly_normals(OPoly* p) {
  for (int i = 0; i < p->n; i++) {
    int j = (i + 1) % p->n;
    double ex = p->vx[j] - p->vx[i], ey = p->vy[j] - p->vy[i];
    double l = sqrt(ex * ex + ey * ey);
    p->nx[i] = ey / l; /* left() of the CCW edge = outward */
    p->ny[i] = -ex / l;
  }
}

/* dyn4j Rectangle vertex order: (-h,-h) (h,-h) (h,h) (-h,h) */
static void mass_poly(const OBody* b, double h, OPoly* p) {
  static const double lx[4] = {-1, 1, 1, -1}, ly[4] = {-1, -1, 1, 1};
  double c = cos(b->th), s = sin(b->th);
  p->n = 4;
  for (int i = 0; i < 4; i++) {
    double x = lx[i] * h, y = ly[i] * h;
    p->vx[i] = b->x + c * x - s * y;
    p->vy[i] = b->y + s * x + c * y;
  }
  /* normals are the rotated axis directions */
  p->nx[0] = s;
  p->ny[0] = -c;
  p->nx[1] = c;
  p->ny[1] = s;
  p->nx[2] = -s;
  p->ny[2] = c;
  p->nx[3] = -c;
  p->ny[3] = -s;
}

/* Ground.java:64-70: Polygon((0,ys[i-1]), (0,baseY), (w,baseY), (w,ys[i])) translated by xs[i-1] */
static void ground_poly(const vsr_batch_desc* d, double base_y, int seg, OPoly* p) {
  double x0 = d->terrain_xs[seg], x1 = d->terrain_xs[seg + 1];
  p->n = 4;
  p->vx[0] = x0;
  p->vy[0] = d->terrain_ys[seg];
  p->vx[1] = x0;
  p->vy[1] = base_y;
  p->vx[2] = x1;
  p->vy[2] = base_y;
  p->vx[3] = x1;
  p->vy[3] = d->terrain_ys[seg + 1];
  poly_normals(p);
}

/* Narrowphase.  dyn4j runs GJK + EPA; for two convex polygons the EPA result is the
   minimum-translation vector, i.e. the face normal (of either polygon) with the
   smallest overlap = the SAT answer restated here.  Normal points from A to B. */
static int sat_penetration(const OPoly* A, const OPoly* B, int mm, double* nx, double* ny, double* depth) {
  /* mm = mass against mass (intra-robot).  Masses of neighbouring voxels touch corner to corner at
     rest, so every comparison there is decided by rule with a margin instead of by the last bit:
     an earlier face wins ties, and a contact needs a penetration above the margin. */
  const double in_eps = mm ? TIE_EPS : 0.0;
  double best = -INFINITY, bx = 0, by = 0;
  for (int i = 0; i < A->n; i++) {
    double s = INFINITY;
    for (int j = 0; j < B->n; j++) {
      double v = A->nx[i] * (B->vx[j] - A->vx[i]) + A->ny[i] * (B->vy[j] - A->vy[i]);
      if (v < s) s = v;
    }
    if (s > best + (i > 0 ? in_eps : 0.0)) {
      best = s;
      bx = A->nx[i];
      by = A->ny[i];
    }
  }
  for (int j = 0; j < B->n; j++) {
    double s = INFINITY;
    for (int i = 0; i < A->n; i++) {
      double v = B->nx[j] * (A->vx[i] - B->vx[j]) + B->ny[j] * (A->vy[i] - B->vy[j]);
      if (v < s) s = v;
    }
    if (s > best + TIE_EPS) { /* structural tie (flat mass on flat ground): keep the mass' own face */
      best = s;
      bx = -B->nx[j];
      by = -B->ny[j];
    }
  }
  if (!(best < -in_eps)) return 0;
  *nx = bx;
  *ny = by;
  *depth = -best;
  return 1;
}

typedef struct {
  double x1, y1, x2, y2; /* CCW: vertex1 -> vertex2 */
  int i1, i2;            /* vertex indices */
  double mx, my;         /* the farthest vertex */
  int index;             /* edge index */
} OEdge;

/* vsr_oracle_opts.dyn4j_labels of the run in progress (one run at a time per process; all threads share it) */
static int g_dyn4j_labels = 0;
/* vsr_oracle_opts.touching_separated, likewise */
static int g_touching_separated = 0;

/* Polygon.getFarthestFeature (dyn4j): farthest vertex, then the adjacent edge most
   perpendicular to n, keeping the CCW winding. */
static void farthest_edge(const OPoly* P, double nx, double ny, OEdge* e) {
  int idx = 0;
  double best = P->vx[0] * nx + P->vy[0] * ny;
  for (int i = 1; i < P->n; i++) {
    double v = P->vx[i] * nx + P->vy[i] * ny;
    if (v > best) {
      best = v;
      idx = i;
    }
  }
  int prev = idx == 0 ? P->n - 1 : idx - 1, next = idx == P->n - 1 ? 0 : idx + 1;
  double ln = P->nx[prev] * nx + P->ny[prev] * ny; /* edge prev->idx */
  double rn = P->nx[idx] * nx + P->ny[idx] * ny;   /* edge idx->next */
  e->mx = P->vx[idx];
  e->my = P->vy[idx];
  if (ln < rn) {
    e->x1 = P->vx[idx];
    e->y1 = P->vy[idx];
    e->i1 = idx;
    e->x2 = P->vx[next];
    e->y2 = P->vy[next];
    e->i2 = next;
    e->index = next == 0 ? P->n : next; /* dyn4j: idx + 1 */
  } else {
    e->x1 = P->vx[prev];
    e->y1 = P->vy[prev];
    e->i1 = prev;
    e->x2 = P->vx[idx];
    e->y2 = P->vy[idx];
    e->i2 = idx;
    /* dyn4j labels this edge `idx`, i.e. 0 for the wrap-around edge (n-1 -> 0) that the other
       branch labels n.  When n is the edge's own normal both end points tie as "farthest" and the
       label would flip with the rounding; one label per edge keeps the warm start deterministic. */
    e->index = (idx == 0 && !g_dyn4j_labels) ? P->n : idx;
  }
}

typedef struct {
  double x, y;
  int idx;
} OClipPt;

/* ClippingManifoldSolver.clip (dyn4j): keep what lies at n.p <= offset. */
static int clip_edge(const OClipPt* v1, const OClipPt* v2, double nx, double ny, double offset, OClipPt* out) {
  int n = 0;
  double d1 = nx * v1->x + ny * v1->y - offset, d2 = nx * v2->x + ny * v2->y - offset;
  if (d1 <= 0.0) out[n++] = *v1;
  if (d2 <= 0.0) out[n++] = *v2;
  if (d1 * d2 < 0.0) {
    double u = d1 / (d1 - d2);
    OClipPt p;
    p.x = v1->x + (v2->x - v1->x) * u;
    p.y = v1->y + (v2->y - v1->y) * u;
    p.idx = d1 > 0.0 ? v1->idx : v2->idx;
    if (n < 2) out[n++] = p;
  }
  return n;
}

/* ClippingManifoldSolver.getManifold (dyn4j), both features are edges. */
static int manifold(const OPoly* A, const OPoly* B, double nx, double ny, OContact* c) {
  OEdge ref, inc;
  farthest_edge(A, nx, ny, &ref);
  farthest_edge(B, -nx, -ny, &inc);
  int flipped = 0;
  double e1 = fabs((ref.x2 - ref.x1) * nx + (ref.y2 - ref.y1) * ny);
  double e2 = fabs((inc.x2 - inc.x1) * nx + (inc.y2 - inc.y1) * ny);
  if (e1 > e2 + TIE_EPS) { /* parallel edges tie: the reference stays on convex1 (the mass) */
    OEdge t = ref;
    ref = inc;
    inc = t;
    flipped = 1;
  }
  double rx = ref.x2 - ref.x1, ry = ref.y2 - ref.y1;
  double rl = sqrt(rx * rx + ry * ry);
  rx /= rl;
  ry /= rl;
  OClipPt a = {inc.x1, inc.y1, inc.i1}, b = {inc.x2, inc.y2, inc.i2}, c1[3], c2[3];
  double o1 = -(rx * ref.x1 + ry * ref.y1);
  if (clip_edge(&a, &b, -rx, -ry, o1, c1) < 2) return 0;
  double o2 = rx * ref.x2 + ry * ref.y2;
  if (clip_edge(&c1[0], &c1[1], rx, ry, o2, c2) < 2) return 0;
  /* inward normal of the CCW reference edge; a clipped point is a contact when it
     lies behind the reference face */
  double fx = -ry, fy = rx;
  double fo = fx * ref.mx + fy * ref.my;
  /* manifold normal: reference face's outward normal, oriented from A to B */
  c->nx = flipped ? fx : -fx;
  c->ny = flipped ? fy : -fy;
  c->n = 0;
  for (int i = 0; i < 2; i++) {
    double depth = fx * c2[i].x + fy * c2[i].y - fo;
    if (depth >= 0.0) {
      OCPoint* p = &c->p[c->n++];
      memset(p, 0, sizeof *p);
      p->px = c2[i].x;
      p->py = c2[i].y;
      p->depth = depth;
      /* IndexedManifoldPointId(referenceEdge, incidentEdge, incidentVertex, flipped) */
      p->id = (ref.index & 7) | ((inc.index & 7) << 3) | ((c2[i].idx & 7) << 6) | (flipped << 9);
    }
  }
  return c->n > 0;
}

static OContact* push_contact(ORobot* r) {
  if (r->ncontacts == r->cap_contacts) {
    r->cap_contacts = r->cap_contacts ? 2 * r->cap_contacts : 32;
    r->contacts = (OContact*)realloc(r->contacts, sizeof(OContact) * (size_t)r->cap_contacts);
  }
  return &r->contacts[r->ncontacts++];
}

/* ContactConstraint.update (dyn4j): carry jn/jt over for points with the same id. */
static void warm_match(ORobot* r, OContact* c) {
  for (int i = 0; i < r->nold; i++) {
    OContact* o = &r->old[i];
    if (o->b1 != c->b1 || o->b2 != c->b2 || o->seg != c->seg) continue;
    for (int k = 0; k < c->n; k++)
      for (int l = 0; l < o->n; l++)
        if (o->p[l].id == c->p[k].id) {
          c->p[k].jn = o->p[l].jn;
          c->p[k].jt = o->p[l].jt;
        }
    return;
  }
}

/* World.detect (dyn4j) restricted to what can collide here: every mass against the
   ground segments its AABB overlaps (and, optionally, other masses of the robot). */
static void detect(ORun* R, ORobot* r) {
  const vsr_batch_desc* d = R->d;
  /* swap current -> old */
  OContact* t = r->old;
  int tc = r->cap_old;
  r->old = r->contacts;
  r->nold = r->ncontacts;
  r->cap_old = r->cap_contacts;
  r->contacts = t;
  r->cap_contacts = tc;
  r->ncontacts = 0;
  double mu_g = sqrt(r->mat->friction * d->settings.ground_friction); /* CoefficientMixer: sqrt(f1 f2) */
  double e_g = fmax(r->mat->restitution, d->settings.ground_restitution); /* max(e1,e2) */
  memset(r->touch, 0, (size_t)r->nv);
  for (int b = 0; b < r->nb; b++) {
    OPoly A;
    mass_poly(&r->bodies[b], r->half, &A);
    double minx = INFINITY, maxx = -INFINITY, miny = INFINITY;
    for (int i = 0; i < 4; i++) {
      if (A.vx[i] < minx) minx = A.vx[i];
      if (A.vx[i] > maxx) maxx = A.vx[i];
      if (A.vy[i] < miny) miny = A.vy[i];
    }
    for (int seg = 0; seg + 1 < d->terrain_n; seg++) {
      if (d->terrain_xs[seg + 1] < minx || d->terrain_xs[seg] > maxx) continue;
      if (!(d->terrain_xs[seg + 1] > d->terrain_xs[seg])) continue; /* zero-width: no polygon */
      if (miny > fmax(d->terrain_ys[seg], d->terrain_ys[seg + 1])) continue;
      OPoly G;
      ground_poly(d, R->base_y, seg, &G);
      double nx, ny, depth;
      if (!sat_penetration(&A, &G, 0, &nx, &ny, &depth)) continue;
      OContact c;
      memset(&c, 0, sizeof c);
      c.b1 = b;
      c.b2 = -1;
      c.seg = seg;
      c.mu = mu_g;
      c.e = e_g;
      if (!manifold(&A, &G, nx, ny, &c)) continue;
      warm_match(r, &c);
      *push_contact(r) = c;
      r->touch[b / 4] = 1; /* Touch.java:35-48: a body whose userData != this robot */
    }
  }
  if (R->o.self_collision || d->settings.self_collision) {
    /* masses of the same robot collide unless welded (Voxel.java:178-185,474: RobotFilter allows
       every pair, the springs allow collisions, WeldJoint keeps dyn4j's default = not allowed) */
    double mu_s = sqrt(r->mat->friction * r->mat->friction);
    double e_s = r->mat->restitution;
    int canonical = R->o.order == VSR_ORACLE_ORDER_CANONICAL;
    int npairs = canonical ? r->nv * r->nv * 16 : r->nb * r->nb;
    for (int q = 0; q < npairs; q++) {
      int a, b;
      if (canonical) { /* the kernel's enumeration: owner voxel, partner voxel, partner mass, own mass */
        int va = q / (r->nv * 16), rem = q % (r->nv * 16);
        int vb = rem / 16, kb = (rem / 4) % 4, ka = rem % 4;
        a = 4 * va + ka;
        b = 4 * vb + kb;
      } else {
        a = q / r->nb;
        b = q % r->nb;
      }
      if (!(a < b)) continue;
      {
        OBody* A0 = &r->bodies[a];
        OBody* B0 = &r->bodies[b];
        double rr = 2.0 * r->half * 1.4142135623730951;
        if (fabs(A0->x - B0->x) > rr || fabs(A0->y - B0->y) > rr) continue;
        int welded = 0;
        for (int w = 0; w < r->nw && !welded; w++)
          if ((r->welds[w].b1 == a && r->welds[w].b2 == b) || (r->welds[w].b1 == b && r->welds[w].b2 == a))
            welded = 1;
        if (welded) continue;
        OPoly A, B;
        mass_poly(A0, r->half, &A);
        mass_poly(B0, r->half, &B);
        double nx, ny, depth;
        if (!sat_penetration(&A, &B, 1, &nx, &ny, &depth)) continue;
        OContact c;
        memset(&c, 0, sizeof c);
        c.b1 = a;
        c.b2 = b;
        c.seg = -1;
        c.mu = mu_s;
        c.e = e_s;
        if (!manifold(&A, &B, nx, ny, &c)) continue;
        warm_match(r, &c);
        *push_contact(r) = c;
      }
    }
  }
}

/* ------------------------------------------------------------------ dyn4j C.5: solver */
static const OBody STATIC_BODY = {0, 0, 0, 0, 0, 0, 0, 0};

/* SequentialImpulses.initialize (dyn4j) for one constraint, then its warm start.
   Convention: N points from body1 to body2; lambda >= 0 pushes body1 along -N and
   body2 along +N (Box2D's A/B roles). */
static void contact_init(const vsr_settings* s, ORobot* r, OContact* c) {
  OBody* B1 = &r->bodies[c->b1];
  const OBody* B2 = c->b2 >= 0 ? &r->bodies[c->b2] : &STATIC_BODY;
  double c1 = cos(B1->th), s1 = sin(B1->th), c2 = cos(B2->th), s2 = sin(B2->th);
  double nx = c->nx, ny = c->ny, tx = ny, ty = -nx; /* tangent = left-hand orthogonal of N */
  for (int k = 0; k < c->n; k++) {
    OCPoint* p = &c->p[k];
    p->r1x = p->px - B1->x;
    p->r1y = p->py - B1->y;
    p->r2x = p->px - B2->x;
    p->r2y = p->py - B2->y;
    p->l1x = c1 * p->r1x + s1 * p->r1y;
    p->l1y = -s1 * p->r1x + c1 * p->r1y;
    p->l2x = c2 * p->r2x + s2 * p->r2y;
    p->l2y = -s2 * p->r2x + c2 * p->r2y;
    double rn1 = p->r1x * ny - p->r1y * nx, rn2 = p->r2x * ny - p->r2y * nx;
    p->massN = 1.0 / (B1->invM + B2->invM + B1->invI * rn1 * rn1 + B2->invI * rn2 * rn2);
    double rt1 = p->r1x * ty - p->r1y * tx, rt2 = p->r2x * ty - p->r2y * tx;
    p->massT = 1.0 / (B1->invM + B2->invM + B1->invI * rt1 * rt1 + B2->invI * rt2 * rt2);
    /* relative velocity of body2 w.r.t. body1 at the point, along N: < 0 = approaching */
    double rvx = (B2->vx - B2->w * p->r2y) - (B1->vx - B1->w * p->r1y);
    double rvy = (B2->vy + B2->w * p->r2x) - (B1->vy + B1->w * p->r1x);
    double rvn = nx * rvx + ny * rvy;
    p->vb = 0.0;
    if (rvn < -s->restitution_velocity) p->vb += -c->e * rvn;
  }
  c->block = 0;
  if (c->n == 2) {
    double rn1A = c->p[0].r1x * ny - c->p[0].r1y * nx, rn1B = c->p[0].r2x * ny - c->p[0].r2y * nx;
    double rn2A = c->p[1].r1x * ny - c->p[1].r1y * nx, rn2B = c->p[1].r2x * ny - c->p[1].r2y * nx;
    double m = B1->invM + B2->invM;
    c->K00 = m + B1->invI * rn1A * rn1A + B2->invI * rn1B * rn1B;
    c->K01 = m + B1->invI * rn1A * rn2A + B2->invI * rn1B * rn2B;
    c->K11 = m + B1->invI * rn2A * rn2A + B2->invI * rn2B * rn2B;
    double det = c->K00 * c->K11 - c->K01 * c->K01;
    const double maxCondition = 1000.0;
    if (c->K00 * c->K00 < maxCondition * det) {
      c->block = 1;
      double id = 1.0 / det;
      c->iK00 = id * c->K11;
      c->iK01 = -id * c->K01;
      c->iK11 = id * c->K00;
    } else {
      c->n = 1; /* ill conditioned: keep only the first point */
    }
  }
}

static void contact_apply(OBody* B1, OBody* B2m, const OCPoint* p, double Px, double Py) {
  B1->vx -= B1->invM * Px;
  B1->vy -= B1->invM * Py;
  B1->w -= B1->invI * (p->r1x * Py - p->r1y * Px);
  if (B2m) {
    B2m->vx += B2m->invM * Px;
    B2m->vy += B2m->invM * Py;
    B2m->w += B2m->invI * (p->r2x * Py - p->r2y * Px);
  }
}

static void contact_warm(ORobot* r, OContact* c) {
  OBody* B1 = &r->bodies[c->b1];
  OBody* B2 = c->b2 >= 0 ? &r->bodies[c->b2] : NULL;
  double nx = c->nx, ny = c->ny, tx = ny, ty = -nx;
  for (int k = 0; k < c->n; k++) {
    OCPoint* p = &c->p[k];
    contact_apply(B1, B2, p, nx * p->jn + tx * p->jt, ny * p->jn + ty * p->jt);
  }
}

static void rel_vel(const OBody* B1, const OBody* B2, const OCPoint* p, double* rvx, double* rvy) {
  *rvx = (B2->vx - B2->w * p->r2y) - (B1->vx - B1->w * p->r1y);
  *rvy = (B2->vy + B2->w * p->r2x) - (B1->vy + B1->w * p->r1x);
}

/* SequentialImpulses.solveVelocityConstraints (dyn4j) for one constraint: friction of
   every point first, then the normal constraint(s) (block solver for 2 points). */
static void contact_solve(ORobot* r, OContact* c) {
  OBody* B1 = &r->bodies[c->b1];
  OBody* B2m = c->b2 >= 0 ? &r->bodies[c->b2] : NULL;
  const OBody* B2 = B2m ? B2m : &STATIC_BODY;
  double nx = c->nx, ny = c->ny, tx = ny, ty = -nx;
  for (int k = 0; k < c->n; k++) {
    OCPoint* p = &c->p[k];
    double rvx, rvy;
    rel_vel(B1, B2, p, &rvx, &rvy);
    double tn = tx * rvx + ty * rvy;
    double jt = p->massT * (-tn);
    double maxJt = c->mu * p->jn;
    double Jt0 = p->jt;
    p->jt = fmax(-maxJt, fmin(Jt0 + jt, maxJt));
    jt = p->jt - Jt0;
    contact_apply(B1, B2m, p, tx * jt, ty * jt);
  }
  if (c->n == 1 || !c->block) {
    for (int k = 0; k < c->n; k++) {
      OCPoint* p = &c->p[k];
      double rvx, rvy;
      rel_vel(B1, B2, p, &rvx, &rvy);
      double rvn = nx * rvx + ny * rvy;
      double j = -p->massN * (rvn - p->vb);
      double j0 = p->jn;
      p->jn = fmax(j0 + j, 0.0);
      j = p->jn - j0;
      contact_apply(B1, B2m, p, nx * j, ny * j);
    }
  } else {
    OCPoint* p1 = &c->p[0];
    OCPoint* p2 = &c->p[1];
    double ax = p1->jn, ay = p2->jn;
    double rvx, rvy;
    rel_vel(B1, B2, p1, &rvx, &rvy);
    double rvn1 = nx * rvx + ny * rvy;
    rel_vel(B1, B2, p2, &rvx, &rvy);
    double rvn2 = nx * rvx + ny * rvy;
    double bx = rvn1 - p1->vb, by = rvn2 - p2->vb;
    bx -= c->K00 * ax + c->K01 * ay;
    by -= c->K01 * ax + c->K11 * ay;
    double x = 0, y = 0;
    int ok = 0;
    /* case 1: both active */
    x = -(c->iK00 * bx + c->iK01 * by);
    y = -(c->iK01 * bx + c->iK11 * by);
    if (x >= 0.0 && y >= 0.0) ok = 1;
    if (!ok) { /* case 2: x1 active, x2 = 0 */
      x = -p1->massN * bx;
      y = 0.0;
      double v2 = c->K01 * x + by;
      if (x >= 0.0 && v2 >= 0.0) ok = 1;
    }
    if (!ok) { /* case 3 */
      x = 0.0;
      y = -p2->massN * by;
      double v1 = c->K01 * y + bx;
      if (y >= 0.0 && v1 >= 0.0) ok = 1;
    }
    if (!ok) { /* case 4 */
      x = 0.0;
      y = 0.0;
      if (bx >= 0.0 && by >= 0.0) ok = 1;
    }
    if (ok) {
      double dx = x - ax, dy = y - ay;
      contact_apply(B1, B2m, p1, nx * dx, ny * dx);
      contact_apply(B1, B2m, p2, nx * dy, ny * dy);
      p1->jn = x;
      p2->jn = y;
    }
  }
}

/* SequentialImpulses.solvePositionContraints (dyn4j) for one constraint.
   Returns the minimum separation seen. */
static double contact_position(const vsr_settings* s, ORobot* r, OContact* c) {
  OBody* B1 = &r->bodies[c->b1];
  OBody* B2m = c->b2 >= 0 ? &r->bodies[c->b2] : NULL;
  double minSep = 0.0;
  double nx = c->nx, ny = c->ny;
  for (int k = 0; k < c->n; k++) {
    OCPoint* p = &c->p[k];
    const OBody* B2 = B2m ? B2m : &STATIC_BODY;
    double c1 = cos(B1->th), s1 = sin(B1->th), c2 = cos(B2->th), s2 = sin(B2->th);
    double r1x = c1 * p->l1x - s1 * p->l1y, r1y = s1 * p->l1x + c1 * p->l1y;
    double r2x = c2 * p->l2x - s2 * p->l2y, r2y = s2 * p->l2x + c2 * p->l2y;
    double p1x = B1->x + r1x, p1y = B1->y + r1y;
    double p2x = (B2m ? B2->x : p->px - p->r2x) + r2x, p2y = (B2m ? B2->y : p->py - p->r2y) + r2y;
    double sep = (p2x - p1x) * nx + (p2y - p1y) * ny - p->depth;
    if (sep < minSep) minSep = sep;
    double cl = sep + s->linear_tolerance;
    cl = cl < -s->max_linear_correction ? -s->max_linear_correction : (cl > 0.0 ? 0.0 : cl);
    double cp = s->baumgarte * cl;
    double rn1 = r1x * ny - r1y * nx, rn2 = r2x * ny - r2y * nx;
    double K = B1->invM + B2->invM + B1->invI * rn1 * rn1 + B2->invI * rn2 * rn2;
    double jp = (K > EPS_E) ? (-cp / K) : 0.0;
    double Jx = nx * jp, Jy = ny * jp;
    B1->x -= Jx * B1->invM;
    B1->y -= Jy * B1->invM;
    B1->th -= B1->invI * (r1x * Jy - r1y * Jx);
    if (B2m) {
      B2m->x += Jx * B2m->invM;
      B2m->y += Jy * B2m->invM;
      B2m->th += B2m->invI * (r2x * Jy - r2y * Jx);
    }
  }
  return minSep;
}

/* ------------------------------------------------------------------ World.step(1) */

/* ------------------------------------------------------------------ dyn4j C.7: continuous collision detection
   World.solveTOI (dyn4j 4.2.1, recalled) with ContinuousDetectionMode.ALL: after the islands are solved every
   dynamic body that is not a bullet is tested against the STATIC bodies it is not already in contact with
   (body1.isInContact(body2): a contact constraint from the previous detect()); ConservativeAdvancement finds the
   first time in [0, 1] of the step at which the two shapes come within distanceEpsilon (cbrt(ulp) ~ 6.06e-6) when
   the body moves from its pose before the step by (v dt, w dt); the body is put back at that pose
   (transform0.lerp) and TimeOfImpactSolver pushes it linearTolerance INTO the other shape so that the next
   discrete detect() finds the pair.  Velocities are not touched. */
typedef struct { double x, y, th; } OPose;

/* distance between two disjoint convex polygons (closest points p1 on A, p2 on B); returns 0 when they overlap */
static int poly_distance(const OPoly* A, const OPoly* B, double* dist, double* p1x, double* p1y, double* p2x, double* p2y) {
  double nx, ny, depth;
  if (!g_touching_separated && sat_penetration(A, B, 0, &nx, &ny, &depth)) return 0;
  double best = INFINITY;
  for (int pass = 0; pass < 2; pass++) {
    const OPoly* P = pass ? B : A; /* vertices of P against the edges of Q */
    const OPoly* Q = pass ? A : B;
    for (int i = 0; i < P->n; i++)
      for (int j = 0; j < Q->n; j++) {
        int k = (j + 1) % Q->n;
        double ex = Q->vx[k] - Q->vx[j], ey = Q->vy[k] - Q->vy[j];
        /* (the parameter by a multiplication with 1 / |edge|^2 and the comparison on squared distances, one square
           root at the end: the arithmetic of the kernel's quad_distance) */
        double t = ((P->vx[i] - Q->vx[j]) * ex + (P->vy[i] - Q->vy[j]) * ey) * (1.0 / (ex * ex + ey * ey));
        t = t < 0 ? 0 : (t > 1 ? 1 : t);
        double qx = Q->vx[j] + t * ex, qy = Q->vy[j] + t * ey;
        double dx = P->vx[i] - qx, dy = P->vy[i] - qy, dd = dx * dx + dy * dy;
        if (dd < best) {
          best = dd;
          if (pass) { *p1x = qx; *p1y = qy; *p2x = P->vx[i]; *p2y = P->vy[i]; }
          else { *p1x = P->vx[i]; *p1y = P->vy[i]; *p2x = qx; *p2y = qy; }
        }
      }
  }
  /* Touching counts as not separated (Gjk.distance returns false when the closest point of the simplex to the origin
     isZero()).  The margin is not cosmetic: a mass that translates without rotating -- a 2-point ground contact leaves
     it w = 0 -- lands EXACTLY on the surface after its first advance, at a distance of +-1e-16 that the last bit of a
     projection decides; with the margin every implementation reads that landing the same way (overshoot: back up). */
  if (best <= 1e-18) {
    if (!g_touching_separated) return 0;
    *dist = 0.0; /* the other reading (vsr_oracle_opts.touching_separated): a separation of zero, found before the axes are asked */
    return 1;
  }
  if (g_touching_separated && sat_penetration(A, B, 0, &nx, &ny, &depth)) return 0;
  *dist = sqrt(best);
  return 1;
}

static void ccd_pass(ORun* R, ORobot* r, const OPose* pose0, double dt) {
  const vsr_batch_desc* d = R->d;
  const double dist_eps = cbrt(EPS_E);
  const double rmax = r->half * 1.4142135623730951; /* Rectangle.getRadius */
  for (int b = 0; b < r->nb; b++) {
    OBody* B = &r->bodies[b];
    const double dpx = B->vx * dt, dpy = B->vy * dt, da = B->w * dt;
    /* swept AABB of the mass */
    OBody b0 = *B;
    b0.x = pose0[b].x; b0.y = pose0[b].y; b0.th = pose0[b].th;
    OPoly P0, P1;
    mass_poly(&b0, r->half, &P0);
    mass_poly(B, r->half, &P1);
    double minx = INFINITY, maxx = -INFINITY, miny = INFINITY;
    for (int i = 0; i < 4; i++) {
      minx = fmin(minx, fmin(P0.vx[i], P1.vx[i]));
      maxx = fmax(maxx, fmax(P0.vx[i], P1.vx[i]));
      miny = fmin(miny, fmin(P0.vy[i], P1.vy[i]));
    }
    double t2 = 1.0, best_d = 0, bp1x = 0, bp1y = 0, bp2x = 0, bp2y = 0, bnx = 0, bny = 0;
    int found = 0;
    for (int seg = 0; seg + 1 < d->terrain_n; seg++) {
      if (d->terrain_xs[seg + 1] < minx || d->terrain_xs[seg] > maxx) continue;
      if (!(d->terrain_xs[seg + 1] > d->terrain_xs[seg])) continue;
      if (miny > fmax(d->terrain_ys[seg], d->terrain_ys[seg + 1])) continue;
      int in_contact = 0; /* body1.isInContact(body2): constraints of the last detect() */
      for (int i = 0; i < r->ncontacts && !in_contact; i++)
        if (r->contacts[i].b1 == b && r->contacts[i].b2 < 0 && r->contacts[i].seg == seg) in_contact = 1;
      if (in_contact) continue;
      OPoly G;
      ground_poly(d, R->base_y, seg, &G);
      /* ConservativeAdvancement.getTimeOfImpact on [0, t2] */
      OBody lb = b0;
      OPoly A;
      mass_poly(&lb, r->half, &A);
      double dist, p1x, p1y, p2x, p2y;
      if (!poly_distance(&A, &G, &dist, &p1x, &p1y, &p2x, &p2y)) continue; /* not separated at t = 0 */
      double l = 0.0, l0 = 0.0, lnx = 0.0, lny = 0.0; /* (the last normal: the push of a pair found touching) */
      int ok = 1, iter = 0;
      if (dist >= dist_eps) {
        const double amax = rmax * fabs(da);
        if (sqrt(dpx * dpx + dpy * dpy) + amax == 0.0) continue;
        while (dist > dist_eps && iter < 30) {
          double nx = (p2x - p1x) / dist, ny = (p2y - p1y) / dist;
          lnx = nx; lny = ny;
          double drel = dpx * nx + dpy * ny + amax;
          if (drel <= EPS_E) { ok = 0; break; }
          double adv = dist / drel;
          l += adv;
          if (l < 0.0 || l > t2) { ok = 0; break; }
          if (l <= l0) break;
          l0 = l;
          iter++;
          lb.x = b0.x + dpx * l; lb.y = b0.y + dpy * l; lb.th = b0.th + da * l;
          mass_poly(&lb, r->half, &A);
          if (!poly_distance(&A, &G, &dist, &p1x, &p1y, &p2x, &p2y)) { /* overshoot: back up */
            l -= adv;
            lb.x = b0.x + dpx * l; lb.y = b0.y + dpy * l; lb.th = b0.th + da * l;
            mass_poly(&lb, r->half, &A);
            poly_distance(&A, &G, &dist, &p1x, &p1y, &p2x, &p2y);
            break;
          }
        }
      }
      if (!ok) continue;
      if (l < t2) { t2 = l; found = 1; best_d = dist; bp1x = p1x; bp1y = p1y; bp2x = p2x; bp2y = p2y; bnx = lnx; bny = lny; }
    }
    if (!found) continue;
    /* body1.transform0.lerp(transform, t): the pose at the time of impact */
    B->x = pose0[b].x + (B->x - pose0[b].x) * t2;
    B->y = pose0[b].y + (B->y - pose0[b].y) * t2;
    B->th = pose0[b].th + (B->th - pose0[b].th) * t2;
    /* TimeOfImpactSolver.solve: push the pair linearTolerance into each other (static body 2) */
    {
      double nx = bp2x - bp1x, ny = bp2y - bp1y, nl = sqrt(nx * nx + ny * ny);
      if (nl > 0) { nx /= nl; ny /= nl; }
      if (g_touching_separated && best_d == 0.0) { nx = bnx; ny = bny; }
      double C = best_d - d->settings.linear_tolerance;
      C = C < -d->settings.max_linear_correction ? -d->settings.max_linear_correction : (C > 0 ? 0 : C);
      double r1x = bp1x - B->x, r1y = bp1y - B->y;
      double rn1 = r1x * ny - r1y * nx;
      double K = B->invM + B->invI * rn1 * rn1;
      double imp = K > EPS_E ? -C / K : 0.0;
      double Jx = nx * imp, Jy = ny * imp;
      B->x += Jx * B->invM;
      B->y += Jy * B->invM;
      B->th += B->invI * (r1x * Jy - r1y * Jx);
    }
  }
}

/* ConstraintGraph.solve (dyn4j 4.2.1, recalled; Box2D's b2World::Solve): islands are built by a depth-first search
   from the first dynamic body in world order that is not on an island yet.  A body popped from the (LIFO) stack is
   added; its contact constraints are visited first, then its joints, each in the node's list order; a constraint not
   yet on the island is appended to the island's list and its other body, if not seen, is pushed (static bodies are
   added but not expanded).  The island then solves joints and contacts IN THAT DISCOVERY ORDER -- not in world
   insertion order.  All masses of a robot hang together through springs and welds: one island per robot.
   Contact lists per body are taken in detection order (dyn4j: the order the pairs first touched -- a broadphase
   artefact that cannot be restated). */
static void island_order(ORobot* r) {
  int nb = r->nb, nc = r->ncontacts;
  if (r->cap_corder < nc + 1) {
    r->cap_corder = 2 * nc + 16;
    r->corder = (int*)realloc(r->corder, sizeof(int) * (size_t)r->cap_corder);
  }
  unsigned char* on_body = (unsigned char*)calloc((size_t)nb + 1, 1);
  unsigned char* on_joint = (unsigned char*)calloc((size_t)r->nj + 1, 1); /* springs [0, ns), welds ns + w */
  unsigned char* on_con = (unsigned char*)calloc((size_t)nc + 1, 1);
  int* stack = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
  int nj_out = 0, nc_out = 0;
  for (int seed = 0; seed < nb; seed++) {
    if (on_body[seed]) continue;
    int sp = 0;
    stack[sp++] = seed;
    on_body[seed] = 1;
    while (sp > 0) {
      int b = stack[--sp];
      for (int i = 0; i < nc; i++) { /* node.contactConstraints */
        OContact* c = &r->contacts[i];
        if (on_con[i] || (c->b1 != b && c->b2 != b)) continue;
        on_con[i] = 1;
        r->corder[nc_out++] = i;
        int other = c->b1 == b ? c->b2 : c->b1;
        if (other >= 0 && !on_body[other]) { stack[sp++] = other; on_body[other] = 1; }
      }
      for (int q = r->adj_start[b]; q < r->adj_start[b + 1]; q++) { /* node.joints */
        int j = r->adj[q];
        int slot = j >= 0 ? j : r->ns + (~j);
        if (on_joint[slot]) continue;
        on_joint[slot] = 1;
        r->jorder[nj_out++] = j;
        int b1 = j >= 0 ? r->springs[j].b1 : r->welds[~j].b1, b2 = j >= 0 ? r->springs[j].b2 : r->welds[~j].b2;
        int other = b1 == b ? b2 : b1;
        if (!on_body[other]) { stack[sp++] = other; on_body[other] = 1; }
      }
    }
  }
  free(on_body); free(on_joint); free(on_con); free(stack);
}

static int world_step(ORun* R, ORobot* r, double dt) {
  const vsr_batch_desc* d = R->d;
  const vsr_settings* s = &d->settings;
  /* an island at rest is not simulated (ConstraintGraph.solve starts islands from bodies that are not at rest); its
     manifolds do not change, so the detection is skipped with it */
  if (R->o.sleeping && r->asleep) return 0;
  const int dfs = R->o.order == VSR_ORACLE_ORDER_DFS;
  if (dfs) island_order(r);
  OPose* pose0 = NULL;
  if (R->o.ccd || d->settings.continuous_detection) { /* body.transform0 = body.transform, before the islands are solved */
    pose0 = (OPose*)malloc(sizeof(OPose) * (size_t)r->nb);
    for (int b = 0; b < r->nb; b++) { pose0[b].x = r->bodies[b].x; pose0[b].y = r->bodies[b].y; pose0[b].th = r->bodies[b].th; }
  }
#define CON(i) (&r->contacts[dfs ? r->corder[i] : (i)])
  /* Island.solve (dyn4j): integrate velocities */
  for (int b = 0; b < r->nb; b++) integrate_velocity(d, r->mat, &r->bodies[b], dt);
  /* solver.initialize(contacts): all masses/biases first, then all warm starts */
  for (int i = 0; i < r->ncontacts; i++) contact_init(s, r, CON(i));
  for (int i = 0; i < r->ncontacts; i++) contact_warm(r, CON(i));
  /* joints: initializeConstraints (+ warm start) */
  for (int i = 0; i < r->nj; i++) {
    int j = r->jorder[i];
    if (j >= 0) spring_init(d, r, &r->springs[j], dt);
    else weld_init(r, &r->welds[~j]);
  }
  /* velocity iterations: all joints, then all contacts */
  for (int it = 0; it < s->velocity_iterations; it++) {
    for (int i = 0; i < r->nj; i++) {
      int j = r->jorder[i];
      if (j >= 0) {
        spring_solve(r, &r->springs[j]);
        if (r->springs[j].lo_on || r->springs[j].hi_on) spring_limits_solve(r, &r->springs[j], 1.0 / dt);
      } else weld_solve(r, &r->welds[~j]);
    }
    for (int i = 0; i < r->ncontacts; i++) contact_solve(r, CON(i));
  }
  for (int b = 0; b < r->nb; b++) integrate_position(s, &r->bodies[b], dt);
  /* position iterations: contacts, then joints; stop when both report solved */
  int iters = 0;
  for (int it = 0; it < s->position_iterations; it++) {
    iters++;
    double minSep = 0.0;
    for (int i = 0; i < r->ncontacts; i++) {
      double ms = contact_position(s, r, CON(i));
      if (ms < minSep) minSep = ms;
    }
    int contactsSolved = minSep >= -3.0 * s->linear_tolerance;
    int jointsSolved = 1;
    for (int i = 0; i < r->nj; i++) {
      int j = r->jorder[i];
      if (j >= 0) { /* DistanceJoint with frequency > 0: "solved" unless a limit is enabled */
        if (r->springs[j].lo_on || r->springs[j].hi_on || r->springs[j].rigid)
          jointsSolved = spring_limits_position(s, r, &r->springs[j]) && jointsSolved;
        continue;
      }
      int ok = weld_position(s, r, &r->welds[~j]);
      jointsSolved = jointsSolved && ok;
    }
    if (contactsSolved && jointsSolved) break;
  }
#undef CON
  if (R->o.sleeping) {
    /* Island.solve, at-rest detection (dyn4j; Settings defaults: 0.01 u/s, 2 degrees/s, 0.5 s) */
    const double lin2 = 0.01 * 0.01, ang = 2.0 * PI / 180.0, max_rest = 0.5;
    double min_rest = INFINITY;
    for (int b = 0; b < r->nb; b++) {
      OBody* B = &r->bodies[b];
      if (B->vx * B->vx + B->vy * B->vy > lin2 || fabs(B->w) > ang) {
        r->rest_time[b] = 0.0;
        min_rest = 0.0;
      } else {
        r->rest_time[b] += dt;
        if (r->rest_time[b] < min_rest) min_rest = r->rest_time[b];
      }
    }
    if (min_rest >= max_rest) {
      r->asleep = 1; /* Body.setAtRest(true): velocities zeroed */
      for (int b = 0; b < r->nb; b++) r->bodies[b].vx = r->bodies[b].vy = r->bodies[b].w = 0.0;
    }
  }
  if (pose0) {
    ccd_pass(R, r, pose0, dt);
    free(pose0);
  }
  if (R->o.state_fp32) {
    const double ox = d->initial_placement;
    for (int b = 0; b < r->nb; b++) {
      OBody* B = &r->bodies[b];
      B->x = ox + (double)(float)(B->x - ox);
      B->y = (double)(float)B->y;
      B->th = (double)(float)B->th;
      B->vx = (double)(float)B->vx;
      B->vy = (double)(float)B->vy;
      B->w = (double)(float)B->w;
    }
  }
  /* new contacts for Touch and for the next step */
  detect(R, r);
  return iters;
}

/* ------------------------------------------------------------------ voxel getters */
/* outer corner of body k: getIndexedVertex(k, 3-k), Voxel.java:536-542,596-603 */
static void outer_corner(double half, const OBody* b, int k, double* x, double* y) {
  static const double lx[4] = {-1, +1, +1, -1}, ly[4] = {+1, +1, -1, -1};
  double c = cos(b->th), s = sin(b->th), px = lx[k] * half, py = ly[k] * half;
  *x = b->x + c * px - s * py;
  *y = b->y + s * px + c * py;
}

/* Voxel.area (:508-516) with Poly.area (Poly.java:40-49), Voxel.getAreaRatio (:524-526) */
static double voxel_area_ratio(const ORun* R, const ORobot* r, int v) {
  double X[4], Y[4];
  for (int k = 0; k < 4; k++) outer_corner(r->half, &r->bodies[4 * v + k], k, &X[k], &Y[k]);
  double a = 0.0;
  for (int i = 0; i < 4; i++) a = a + X[i] * (Y[(4 + i + 1) % 4] - Y[(4 + i - 1) % 4]);
  a = 0.5 * fabs(a);
  double L = r->mat->side_length;
  return a / L / L;
}

/* VoxelPoly.center = Poly.center = mean of the 4 outer corners (VoxelPoly.java:109-112) */
static void voxel_poly_center(const ORun* R, const ORobot* r, int v, double* cx, double* cy) {
  double sx = 0, sy = 0;
  for (int k = 0; k < 4; k++) {
    double x, y;
    outer_corner(r->half, &r->bodies[4 * v + k], k, &x, &y);
    sx += x;
    sy += y;
  }
  *cx = sx / 4.0;
  *cy = sy / 4.0;
}

/* Voxel.getAngle, :518-522 */
static double voxel_angle(const ORobot* r, int v) {
  const OBody* b = &r->bodies[4 * v];
  double up = atan2(b[1].y - b[0].y, b[1].x - b[0].x);
  double down = atan2(b[2].y - b[3].y, b[2].x - b[3].x);
  return (up + down) / 2.0;
}

/* DoubleRange.normalize + SoftNormalization.sense, SoftNormalization.java:38-48 */
static double soft_norm(double x, double mn, double mx) {
  double cl = fmin(fmax(x, mn), mx);
  double v = (cl - mn) / (mx - mn);
  return tanh(((v * 2.0) - 1.0) * 2.0) / 2.0 + 0.5;
}

/* Voxel.act (:197-203) for all voxels, then sensors (CompositeSensor.act: inner first). */
/* java.util.Random: next(bits), nextDouble, nextGaussian (polar method) for one sensor's generator */
static long long jr_next(unsigned long long* seed, int bits) {
  *seed = (*seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
  return (long long)(*seed >> (48 - bits));
}
static double jr_next_double(unsigned long long* seed) {
  long long a = jr_next(seed, 26), b = jr_next(seed, 27);
  return (double)((a << 27) + b) * (1.0 / (double)(1LL << 53));
}
static double jr_next_gaussian(ORobot* r, int si) {
  if (r->rng_have[si]) {
    r->rng_have[si] = 0;
    return r->rng_spare[si];
  }
  double v1, v2, s;
  do {
    v1 = 2 * jr_next_double(&r->rng_seed[si]) - 1;
    v2 = 2 * jr_next_double(&r->rng_seed[si]) - 1;
    s = v1 * v1 + v2 * v2;
  } while (s >= 1 || s == 0);
  double mul = sqrt(-2 * log(s) / s);
  r->rng_spare[si] = v2 * mul;
  r->rng_have[si] = 1;
  return v1 * mul;
}

/* Lidar.sense (Lidar.java:96-103) for one ray: World.raycast against the ground polygons only (the
   DetectFilter drops the robot's fixtures, :37-46).  A ray that starts above the terrain enters a ground
   polygon through its top edge and the polygons share their sides: the first hit is the nearest crossing
   of the terrain polyline from above; none within the ray length reads the ray length. */
static double lidar_cast(const vsr_batch_desc* d, double ox, double oy, double dx, double dy, double len) {
  double best = len;
  for (int seg = 0; seg + 1 < d->terrain_n; seg++) {
    double x0 = d->terrain_xs[seg], y0 = d->terrain_ys[seg], x1 = d->terrain_xs[seg + 1], y1 = d->terrain_ys[seg + 1];
    if (!(x1 > x0)) continue;
    double ex = x1 - x0, ey = y1 - y0;
    double den = dx * ey - dy * ex;
    if (!(den > 0.0)) continue;
    double ax = x0 - ox, ay = y0 - oy;
    double t = (ax * ey - ay * ex) / den;
    double u = (ax * dy - ay * dx) / den;
    if (t >= 0.0 && u >= 0.0 && u <= 1.0 && t < best) best = t;
  }
  return best;
}

/* java.util.Random.nextInt(bound) */
static int jr_next_int(unsigned long long* seed, int bound) {
  if ((bound & -bound) == bound) return (int)(((long long)bound * jr_next(seed, 31)) >> 31);
  for (;;) {
    long long bits = jr_next(seed, 31);
    long long val = bits % bound;
    if (bits - val + (bound - 1) < (1LL << 31)) return (int)val;
  }
}
static int nth_set_bit(unsigned m, int n) {
  for (int b = 0; b < 32; b++)
    if ((m >> b) & 1u) {
      if (n == 0) return b;
      n--;
    }
  return 0;
}

/* BreakableVoxel.act after super.act (:127-168): snapshot of the readings, trigger counters, one draw per trigger
   threshold (enum order), the break itself (component, malfunction type: two nextInt), the restore.
   updateStructureMalfunctionType (:244-258) runs at a break only -- a restore leaves the springs as they are. */
static void breakable_act(ORun* R, ORobot* r, int tick) {
  const double t = R->ticks[tick];
  for (int v = 0; v < r->nv; v++) {
    OBreak* B = &r->brk[v];
    if (!B->enabled) continue;
    if (B->state[VSR_COMPONENT_SENSORS] == VSR_MALFUNCTION_NONE || !B->have_snapshot) {
      for (int q = r->vox_read_off[v]; q < r->vox_read_off[v + 1]; q++) r->snap[q] = r->readings[q];
      B->have_snapshot = 1;
    }
    B->cnt[VSR_TRIGGER_TIME] = B->cnt[VSR_TRIGGER_TIME] + t - B->lastT;
    B->cnt[VSR_TRIGGER_CONTROL] = B->cnt[VSR_TRIGGER_CONTROL] + (r->ctrl_energy[v] - B->lastCE);
    B->cnt[VSR_TRIGGER_AREA] = B->cnt[VSR_TRIGGER_AREA] + (r->ar_energy[v] - B->lastARE);
    B->lastT = t;
    B->lastCE = r->ctrl_energy[v];
    B->lastARE = r->ar_energy[v];
    int breaking = 0;
    for (int g = 0; g < 3; g++) {
      if (B->thr[g] != B->thr[g]) continue; /* not a key of triggerThresholds */
      double u = jr_next_double(&B->seed);
      if (u < 1.0 - tanh(B->thr[g] / B->cnt[g])) {
        B->cnt[0] = B->cnt[1] = B->cnt[2] = 0.0;
        int ncomp = (B->mal[0] != 0) + (B->mal[1] != 0) + (B->mal[2] != 0);
        if (ncomp > 0) {
          breaking = 1;
          int ci = jr_next_int(&B->seed, ncomp), c = 0;
          for (c = 0; c < 3; c++)
            if (B->mal[c] != 0 && ci-- == 0) break;
          int nt = __builtin_popcount(B->mal[c]);
          int ty = nth_set_bit(B->mal[c], jr_next_int(&B->seed, nt));
          B->state[c] = ty;
          int frozen = B->state[VSR_COMPONENT_STRUCTURE] == VSR_MALFUNCTION_FROZEN;
          if (frozen != B->struct_frozen) {
            B->struct_frozen = frozen;
            int per = r->ns / (r->nv > 0 ? r->nv : 1);
            for (int k = 0; k < per; k++) r->springs[v * per + k].rigid = frozen;
          }
        }
      }
    }
    if (breaking) B->lastBreakT = t;
    if (t - B->lastBreakT > B->restore) B->state[0] = B->state[1] = B->state[2] = VSR_MALFUNCTION_NONE;
  }
}

/* BreakableVoxel.getSensorReadings (:183-193) for every voxel, into r->eff: what a controller reads */
static const double* effective_readings(ORobot* r) {
  if (!r->brk) return r->readings;
  for (int v = 0; v < r->nv; v++) {
    OBreak* B = &r->brk[v];
    for (int q = r->vox_read_off[v]; q < r->vox_read_off[v + 1]; q++) {
      if (!B->enabled) r->eff[q] = r->readings[q];
      else if (B->state[VSR_COMPONENT_SENSORS] == VSR_MALFUNCTION_ZERO) r->eff[q] = 0.0;
      else if (B->state[VSR_COMPONENT_SENSORS] == VSR_MALFUNCTION_RANDOM)
        r->eff[q] = jr_next_double(&B->seed) * (r->rd_max[q] - r->rd_min[q]) + r->rd_min[q];
      else r->eff[q] = r->snap[q]; /* NONE (the fresh snapshot) and FROZEN (the one of the break) */
    }
  }
  return r->eff;
}

static void robot_sense(ORun* R, ORobot* r, int tick) {
  const double* T = R->ticks;
  for (int v = 0; v < r->nv; v++) {
    double ar = voxel_area_ratio(R, r, v);
    r->ar_energy[v] = r->ar_energy[v] + ar * ar;
    r->ctrl_energy[v] = r->ctrl_energy[v] + r->last_force[v] * r->last_force[v];
  }
  for (int si = 0; si < r->n_sens; si++) {
    const vsr_sensor* S = r->sens[si];
    int v = r->sens_voxel[si];
    int dim = sensor_dim(S);
    double raw[VSR_LIDAR_MAX_RAYS] = {0, 0, 0, 0, 0, 0, 0, 0}, dmin = 0, dmax = 1;
    if (S->base == VSR_SENSOR_LIDAR) {
      const OBody* b = &r->bodies[4 * v];
      double cx = (((b[0].x + b[1].x) + b[2].x) + b[3].x) / 4.0, cy = (((b[0].y + b[1].y) + b[2].y) + b[3].y) / 4.0; /* Voxel.center */
      double ang = voxel_angle(r, v);
      for (int k = 0; k < dim; k++)
        raw[k] = lidar_cast(R->d, cx, cy, cos(S->lidar_dirs[k] + ang), sin(S->lidar_dirs[k] + ang), S->max_velocity_norm);
      dmax = S->max_velocity_norm;
    } else if (S->base == VSR_SENSOR_TOUCH) {
      raw[0] = r->touch[v] ? 1.0 : 0.0;
    } else if (S->base == VSR_SENSOR_AREA_RATIO) {
      raw[0] = voxel_area_ratio(R, r, v);
      dmin = 0.5;
      dmax = 1.5;
    } else if (S->base == VSR_SENSOR_ANGLE) { /* Angle.java */
      raw[0] = voxel_angle(r, v);
      dmin = -PI;
      dmax = PI;
    } else if (S->base == VSR_SENSOR_CONSTANT) { /* Constant.java: domain [min(0, v), max(1, v)] */
      raw[0] = S->max_velocity_norm;
      dmin = fmin(0.0, raw[0]);
      dmax = fmax(1.0, raw[0]);
    } else if (S->base == VSR_SENSOR_APPLIED_FORCE) { /* AppliedForce.java */
      raw[0] = r->last_force[v];
      dmin = -1.0;
    } else if (S->base == VSR_SENSOR_MALFUNCTION) { /* Malfunction.java: BreakableVoxel.isBroken (:234-237), else 0 */
      const OBreak* B = r->brk ? &r->brk[v] : NULL;
      raw[0] = (B && B->enabled && (B->state[0] || B->state[1] || B->state[2])) ? 1.0 : 0.0;
    } else if (S->base == VSR_SENSOR_TIME_SIN) { /* TimeFunction.java with t -> sin(2 pi f t) */
      raw[0] = sin(2.0 * PI * S->max_velocity_norm * T[tick]);
      dmin = -1.0;
    } else if (S->base == VSR_SENSOR_CRUMPLING) { /* Crumpling.java */
      const OBody* b = &r->bodies[4 * v];
      double c = 0.0;
      for (int i = 0; i < 4; i++)
        for (int j = i + 1; j < 4; j++) {
          double dx = b[i].x - b[j].x, dy = b[i].y - b[j].y;
          if (sqrt(dx * dx + dy * dy) < r->mat->side_length * 0.2) c = c + 1.0;
        }
      raw[0] = 2.0 * c / 12.0;
    } else if (S->base == VSR_SENSOR_CONTROL_POWER) { /* ControlPower.java: lastT = the previous sense time */
      raw[0] = r->ctrl_energy[v] / (T[tick] - (tick > 0 ? T[tick - 1] : 0.0));
      dmax = 1.0 / S->max_velocity_norm;
    } else { /* Velocity.sense, Velocity.java:57-80 */
      const OBody* b = &r->bodies[4 * v];
      double vx = 0, vy = 0;
      for (int k = 0; k < 4; k++) {
        vx = vx + b[k].vx;
        vy = vy + b[k].vy;
      }
      vx /= 4.0;
      vy /= 4.0;
      double ang = voxel_angle(r, v);
      int c = 0;
      if (S->axes & VSR_AXIS_X) raw[c++] = S->rotated ? vx * cos(ang) + vy * sin(ang) : vx;
      if (S->axes & VSR_AXIS_Y) raw[c++] = S->rotated ? vx * cos(ang + PI / 2.0) + vy * sin(ang + PI / 2.0) : vy;
      dmin = -S->max_velocity_norm;
      dmax = S->max_velocity_norm;
    }
    double val[VSR_LIDAR_MAX_RAYS];
    for (int k = 0; k < VSR_LIDAR_MAX_RAYS; k++) val[k] = raw[k];
    if (S->aggregator != VSR_AGG_NONE) {
      double* H = r->hist[si];
      for (int k = 0; k < dim; k++) H[(size_t)tick * dim + k] = raw[k];
      int f = r->hist_first[si];
      while (T[f] < T[tick] - S->interval) f++;
      r->hist_first[si] = f;
      if (S->aggregator == VSR_AGG_AVERAGE) {
        for (int k = 0; k < dim; k++) {
          double sum = 0.0;
          for (int q = f; q <= tick; q++) sum = sum + H[(size_t)q * dim + k];
          val[k] = sum / (double)(tick - f + 1);
        }
      } else if (S->aggregator == VSR_AGG_DYNAMIC_NORM) { /* DynamicNormalization.aggregate, :40-58 */
        for (int k = 0; k < dim; k++) {
          double mn = INFINITY, mx = -INFINITY;
          for (int q = f; q <= tick; q++) {
            mn = fmin(mn, H[(size_t)q * dim + k]);
            mx = fmax(mx, H[(size_t)q * dim + k]);
          }
          double v = (raw[k] - mn) / (mx - mn); /* 0/0 = NaN while the window holds one distinct value */
          val[k] = v != v ? v : fmin(fmax(v, 0.0), 1.0); /* Math.min / Math.max keep a NaN */
        }
        dmin = 0.0;
        dmax = 1.0;
      } else { /* Trend.aggregate, Trend.java:37-50 */
        double li = T[tick] - T[f];
        for (int k = 0; k < dim; k++)
          val[k] = (li == 0) ? 0.0 : (H[(size_t)tick * dim + k] - H[(size_t)f * dim + k]) / li;
        double ext = dmax - dmin; /* Trend domain: +-extent/interval, Trend.java:31-33 */
        dmin = -ext / S->interval;
        dmax = ext / S->interval;
      }
    }
    for (int k = 0; k < dim; k++) {
      double o = S->soft_norm == VSR_NORM_SOFT ? soft_norm(val[k], dmin, dmax)
                 : S->soft_norm == VSR_NORM_LINEAR ? (fmin(fmax(val[k], dmin), dmax) - dmin) / (dmax - dmin) /* Normalization.java */
                                                   : val[k];
      if (val[k] != val[k]) o = val[k]; /* Java's Math.min / Math.max return NaN for a NaN argument (DynamicNormalization) */
      if (S->noise_sigma > 0) { /* Noisy.sense, Noisy.java:57-62: sigma * extent of the wrapped sensor's domain */
        double ext = S->soft_norm != VSR_NORM_NONE ? 1.0 : dmax - dmin;
        o = o + jr_next_gaussian(r, si) * (ext * S->noise_sigma);
      }
      r->readings[r->sens_off[si] + k] = o;
      if (r->rd_min) { /* Sensor.getDomains of the outermost wrapper (Noisy keeps its sensor's) */
        r->rd_min[r->sens_off[si] + k] = S->soft_norm != VSR_NORM_NONE ? 0.0 : dmin;
        r->rd_max[r->sens_off[si] + k] = S->soft_norm != VSR_NORM_NONE ? 1.0 : dmax;
      }
    }
  }
  if (r->brk) breakable_act(R, r, tick);
}

/* Voxel.applyForce, :224-237 */
static void voxel_apply_force(ORobot* r, int v, double f) {
  if (fabs(f) > 1.0) f = (f > 0) ? 1.0 : -1.0;
  r->last_force[v] = f;
  int per = r->ns / (r->nv > 0 ? r->nv : 1);
  for (int k = 0; k < per; k++) {
    OSpring* s = &r->springs[v * per + k];
    double old = s->rest;
    if (f >= 0) s->rest = s->rrest - (s->rrest - s->rmin) * f;
    else if (f < 0) s->rest = s->rrest + (s->rmax - s->rrest) * -f;
    if (s->rest != old && r->asleep) { /* DistanceJoint.setRestDistance wakes both bodies when the value changes */
      r->asleep = 0;
      for (int b = 0; b < r->nb; b++) r->rest_time[b] = 0.0;
    }
  }
}

static void robot_control(ORun* R, ORobot* r, size_t ridx, int tick, double* sig_out) {
  const vsr_batch_desc* d = R->d;
  const double* P = d->params + d->param_offsets[ridx];
  double t = R->ticks[tick];
  const double* readings = (d->controller_kind == VSR_CTRL_PHASE_SIN || d->controller_kind == VSR_CTRL_TABULATED)
                               ? r->readings : effective_readings(r); /* (open-loop controllers never ask for them) */
  double out_stack[64];
  double* out = r->nv <= 64 ? out_stack : (double*)malloc(sizeof(double) * (size_t)r->nv);
  if (d->controller_kind == VSR_CTRL_PHASE_SIN) {
    for (int v = 0; v < r->nv; v++) out[v] = sin(2.0 * PI * P[0] * t + P[2 + v]) * P[1]; /* PhaseSin.java:61 */
  } else if (d->controller_kind == VSR_CTRL_TABULATED) {
    for (int v = 0; v < r->nv; v++) out[v] = P[(size_t)tick * (size_t)r->nv + v];
  } else if (d->controller_kind >= VSR_CTRL_DISTRIBUTED) {
    /* DistributedSensing.computeControlSignals (:150-186) with one MultiLayerPerceptron per voxel
       (Starter.java:176-190).  Dir order N(0,-1) E(1,0) S(0,1) W(-1,0) (:83-101); a neighbour's
       signals towards this voxel sit at Dir.adjacent(dir).index (:188-207). */
    static const int DX[4] = {0, 1, 0, -1}, DY[4] = {-1, 0, 1, 0}, ADJ[4] = {2, 3, 0, 1};
    const int sg = d->distributed_signals;
    const int nd = d->controller_kind == VSR_CTRL_DISTRIBUTED_ND; /* DistributedSensingNonDirectional.java:57-78 */
    const double* W = P;
    double in[64], outv[64];
    for (int v = 0; v < r->nv; v++) {
      int nr = r->vox_read_off[v + 1] - r->vox_read_off[v];
      for (int q = 0; q < nr; q++) in[q] = readings[r->vox_read_off[v] + q]; /* getSensorReadings first */
      int c = nr;
      for (int dir = 0; dir < 4; dir++) {
        int ax = r->vox_gx[v] + DX[dir], ay = r->vox_gy[v] + DY[dir];
        int a = (ax >= 0 && ax < r->grid_w && ay >= 0 && ay < r->grid_h) ? r->vox_at[ay * r->grid_w + ax] : -1;
        for (int q = 0; q < sg; q++) in[c + q] = a >= 0 ? r->last_sig[(size_t)a * 4 * sg + (nd ? 0 : ADJ[dir] * sg) + q] : 0.0;
        c += sg;
      }
      r->neurons[0] = c;
      r->neurons[r->n_layers - 1] = nd ? 1 + sg : 1 + 4 * sg;
      vsr_oracle_mlp_apply(d->mlp_activation, r->neurons, r->n_layers, W, in, outv);
      for (int i = 1; i < r->n_layers; i++) W += (size_t)r->neurons[i] * (size_t)(r->neurons[i - 1] + 1);
      out[v] = outv[0];
      for (int q = 0; q < (nd ? sg : 4 * sg); q++) r->cur_sig[(size_t)v * 4 * sg + q] = outv[1 + q];
    }
    for (int q = 0; q < r->nv * 4 * sg; q++) r->last_sig[q] = r->cur_sig[q];
  } else {
    vsr_oracle_mlp_apply(d->mlp_activation, r->neurons, r->n_layers, P, readings, out);
  }
  for (int v = 0; v < r->nv; v++) {
    if (r->brk && r->brk[v].enabled) { /* BreakableVoxel.applyForce (:170-180) */
      OBreak* B = &r->brk[v];
      if (B->state[VSR_COMPONENT_ACTUATOR] == VSR_MALFUNCTION_ZERO) out[v] = 0.0;
      else if (B->state[VSR_COMPONENT_ACTUATOR] == VSR_MALFUNCTION_FROZEN) out[v] = r->last_force[v];
      else if (B->state[VSR_COMPONENT_ACTUATOR] == VSR_MALFUNCTION_RANDOM) out[v] = jr_next_double(&B->seed) * 2.0 - 1.0;
    }
    voxel_apply_force(r, v, out[v]);
    if (sig_out) sig_out[v] = r->last_force[v];
  }
  if (out != out_stack) free(out);
}

static void observe(ORun* R, ORobot* r, double* cx, double* cy, double* are, double* ce) {
  double sx = 0, sy = 0, a = 0, c = 0;
  for (int v = 0; v < r->nv; v++) { /* BehaviorUtils.center, :47-56 */
    double x, y;
    voxel_poly_center(R, r, v, &x, &y);
    sx = sx + x;
    sy = sy + y;
    a += r->ar_energy[v];
    c += r->ctrl_energy[v];
  }
  *cx = sx / (double)r->nv;
  *cy = sy / (double)r->nv;
  *are = a;
  *ce = c;
}

static void run_robot(ORun* R, size_t ridx) {
  const vsr_batch_desc* d = R->d;
  size_t row = ridx - R->o.robot_begin;
  ORobot rob;
  robot_build(R, &rob, ridx);
  ORobot* r = &rob;
  vsr_result* res = &R->out[row];
  memset(res, 0, sizeof *res);
  double dt = d->settings.step_frequency;
  int nt = R->nticks;
  if (R->o.max_ticks > 0 && R->o.max_ticks < nt) nt = R->o.max_ticks;
  const vsr_oracle_probe* pb = R->probe;
  /* World is created with updateRequired = true: the first step() runs detect() first */
  detect(R, r);
  for (int k = 0; k < nt; k++) {
    int iters = world_step(R, r, dt);  /* AbstractTask.java:49 */
    {
      /* VSR_ORACLE_DEBUG="robot:tick": dump the manifolds found at the end of that tick */
      const char* dbg = getenv("VSR_ORACLE_DEBUG");
      long dr = -1, dk = -1, de = -1;
      int got = dbg ? sscanf(dbg, "%ld:%ld:%ld", &dr, &dk, &de) : 0;
      if (got == 2) de = dk;
      if (got >= 2 && (size_t)dr == ridx && k >= dk && k <= de)
        for (int i = 0; i < r->ncontacts; i++) {
          OContact* c = &r->contacts[i];
          fprintf(stderr, "[oracle] tick %d body %d seg %d n %d normal (%.9g, %.9g)\n", k, c->b1, c->seg, c->n, c->nx, c->ny);
          for (int q = 0; q < c->n; q++)
            fprintf(stderr, "           p (%.9g, %.9g) depth %.6e id %d jn %.6e jt %.6e\n", c->p[q].px, c->p[q].py,
                    c->p[q].depth, c->p[q].id, c->p[q].jn, c->p[q].jt);
        }
    }
    robot_sense(R, r, k);              /* Robot.act: Voxel.act for every voxel ... */
    double* sig = (pb && pb->signals) ? pb->signals + ((size_t)row * R->nticks + k) * pb->voxel_stride : NULL;
    robot_control(R, r, ridx, k, sig); /* ... then controller.control, Robot.java:74-77 */
    double cx, cy, are, ce;
    observe(R, r, &cx, &cy, &are, &ce); /* Locomotion.java:198-202 */
    if (k == 0) {
      res->first_center_x = cx;
      res->first_center_y = cy;
      res->t_first = R->ticks[k];
      res->area_ratio_energy_first = are;
      res->control_energy_first = ce;
    }
    res->last_center_x = cx;
    res->last_center_y = cy;
    res->t_last = R->ticks[k];
    res->area_ratio_energy_last = are;
    res->control_energy_last = ce;
    res->n_ticks = k + 1;
    if (!isfinite(cx) || !isfinite(cy)) res->status = 1;
    if (pb) {
      if (pb->trajectory) {
        double* o = pb->trajectory + ((size_t)row * R->nticks + k) * pb->body_stride * 6;
        for (int b = 0; b < r->nb && (size_t)b < pb->body_stride; b++) {
          o[6 * b + 0] = r->bodies[b].x;
          o[6 * b + 1] = r->bodies[b].y;
          o[6 * b + 2] = r->bodies[b].th;
          o[6 * b + 3] = r->bodies[b].vx;
          o[6 * b + 4] = r->bodies[b].vy;
          o[6 * b + 5] = r->bodies[b].w;
        }
      }
      if (pb->readings) {
        double* o = pb->readings + ((size_t)row * R->nticks + k) * pb->reading_stride;
        for (int q = 0; q < r->n_readings && (size_t)q < pb->reading_stride; q++) o[q] = r->readings[q];
      }
      if (pb->touch) {
        double* o = pb->touch + ((size_t)row * R->nticks + k) * pb->voxel_stride;
        for (int v = 0; v < r->nv && (size_t)v < pb->voxel_stride; v++) o[v] = r->touch[v] ? 1.0 : 0.0;
      }
      if (pb->n_contacts) pb->n_contacts[(size_t)row * R->nticks + k] = r->ncontacts;
      if (pb->pos_iters) pb->pos_iters[(size_t)row * R->nticks + k] = iters;
    }
  }
  robot_free(r);
}

static void* worker(void* arg) {
  ORun* R = (ORun*)arg;
  size_t end = R->o.robot_end;
  for (;;) {
    pthread_mutex_lock(&R->mu);
    size_t i = R->next++;
    pthread_mutex_unlock(&R->mu);
    if (i >= end) break;
    run_robot(R, i);
  }
  return NULL;
}

static int validate(const vsr_batch_desc* d) {
  if (!d) return fail(VSR_ERR_INVALID_ARGUMENT, "null descriptor");
  if (d->abi_version != VSR_ABI_VERSION) return fail(VSR_ERR_INVALID_ARGUMENT, "abi version mismatch");
  if (d->n_shapes <= 0 || !d->shapes) return fail(VSR_ERR_INVALID_ARGUMENT, "no shapes");
  if (d->terrain_n < 2 || !d->terrain_xs || !d->terrain_ys)
    return fail(VSR_ERR_INVALID_ARGUMENT, "There must be at least 2 points"); /* Ground.java:52-54 */
  for (int i = 1; i < d->terrain_n; i++)
    if (d->terrain_xs[i] < d->terrain_xs[i - 1])
      return fail(VSR_ERR_INVALID_ARGUMENT, "x coordinates must be sorted"); /* Ground.java:55-58 */
  if (!d->param_offsets || !d->params) return fail(VSR_ERR_INVALID_ARGUMENT, "no controller parameters");
  for (int k = 0; k < d->n_shapes; k++)
    if (d->shapes[k].breakable)
      for (int v = 0, nv = shape_nvox(&d->shapes[k]); v < nv; v++) {
        const vsr_breakable* q = &d->shapes[k].breakable[v];
        if (q->enabled && (q->malfunctions[VSR_COMPONENT_STRUCTURE] & ~((1u << VSR_MALFUNCTION_NONE) | (1u << VSR_MALFUNCTION_FROZEN))))
          return fail(VSR_ERR_INVALID_ARGUMENT, "Unsupported structure malfunction type."); /* BreakableVoxel.java:258 */
        if (q->enabled && ((q->malfunctions[0] | q->malfunctions[1] | q->malfunctions[2]) & ~15u))
          return fail(VSR_ERR_INVALID_ARGUMENT, "unknown malfunction type");
      }
  return 0;
}

int vsr_oracle_run(const vsr_batch_desc* d, const vsr_oracle_opts* opts, vsr_result* out,
                   const vsr_oracle_probe* probe) {
  int rc = validate(d);
  if (rc) return rc;
  ORun R;
  memset(&R, 0, sizeof R);
  R.d = d;
  if (opts) R.o = *opts;
  else vsr_oracle_default_opts(&R.o);
  if (R.o.robot_end == 0 || R.o.robot_end > d->n_robots) R.o.robot_end = d->n_robots;
  if (R.o.robot_begin > R.o.robot_end) return fail(VSR_ERR_INVALID_ARGUMENT, "robot range");
  g_dyn4j_labels = R.o.dyn4j_labels != 0;
  g_touching_separated = R.o.touching_separated != 0;
  R.probe = probe;
  R.out = out;
  double dt = d->settings.step_frequency;
  R.nticks = tick_times_from(d->t_start, d->final_t, dt, NULL, 0); /* a stage of a longer task starts its clock at t_start */
  R.ticks = (double*)malloc(sizeof(double) * (size_t)(R.nticks + 1));
  tick_times_from(d->t_start, d->final_t, dt, R.ticks, R.nticks);
  double miny = INFINITY;
  for (int i = 0; i < d->terrain_n; i++)
    if (d->terrain_ys[i] < miny) miny = d->terrain_ys[i];
  R.base_y = miny - 50.0; /* Ground.MIN_Y_THICKNESS, Ground.java:37,61 */
  R.half = d->material.side_length * d->material.mass_side_length_ratio / 2.0;
  R.next = R.o.robot_begin;
  pthread_mutex_init(&R.mu, NULL);
  int nth = R.o.n_threads > 0 ? R.o.n_threads : 1;
  if (nth == 1) {
    worker(&R);
  } else {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nth);
    for (int i = 0; i < nth; i++) pthread_create(&th[i], NULL, worker, &R);
    for (int i = 0; i < nth; i++) pthread_join(th[i], NULL);
    free(th);
  }
  pthread_mutex_destroy(&R.mu);
  free(R.ticks);
  return 0;
}
