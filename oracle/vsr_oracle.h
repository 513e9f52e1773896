/*
 * vsr_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * fp64 CPU restatement of the hot path of 2D-VSR-Sim (see vsr_oracle.c for the
 * file:line map).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product path
 * (2dhmsr_b200/libvsr_b200.so) never links, loads or calls it.
 *
 * PARITY UNPINNED at the dyn4j boundary: the arithmetic of World.step lives in
 * org.dyn4j:dyn4j:4.2.1 (pom.xml:31-35), which is neither vendored nor runnable
 * here (no JVM in this image), and the reference's own tests hold no physics
 * vector.  Everything OUTSIDE dyn4j (voxel/robot assembly, placement, sensors,
 * controllers, Outcome) follows the reference source line by line and is pinned
 * by the golden vectors in tests/golden/.
 */
#ifndef VSR_ORACLE_H
#define VSR_ORACLE_H

#include "../include/vsr_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* constraint orderings inside one velocity / position iteration */
#define VSR_ORACLE_ORDER_CANONICAL 0 /* springs (voxel, index) -> horizontal welds -> vertical
                                        welds -> contacts (body, segment): the order the CUDA
                                        kernel's graph colouring is equivalent to            */
#define VSR_ORACLE_ORDER_INSERTION 1 /* dyn4j world insertion order (Voxel.java:216-221,
                                        Robot.java:81-88): all joints, then contacts         */
#define VSR_ORACLE_ORDER_SHUFFLED 2  /* seeded random permutation of the joints (measures
                                        how much the Gauss-Seidel order matters)             */

#define VSR_ORACLE_ORDER_DFS 3       /* dyn4j's own: island discovery order of ConstraintGraph.solve (depth-first
                                        from the first body, contacts before joints at every node, LIFO stack),
                                        recomputed every step because mass-mass contacts change the walk     */

typedef struct vsr_oracle_opts {
  int32_t order;        /* VSR_ORACLE_ORDER_* */
  uint32_t seed;        /* SHUFFLED */
  int32_t n_threads;    /* one robot per thread at a time; <=0 -> 1 */
  int32_t max_ticks;    /* >0: stop after this many ticks (short-horizon parity) */
  size_t robot_begin;   /* run robots [robot_begin, robot_end) ; end==0 -> n_robots */
  size_t robot_end;
  int32_t self_collision; /* 1: also solve mass-mass contacts inside a robot (Voxel.java:178-185,474) */
  int32_t ccd;          /* 1: force dyn4j's time-of-impact pass for masses about to touch the ground (C.7) even when
                           vsr_settings.continuous_detection is 0 (it runs whenever either is set)                  */
  int32_t state_fp32;   /* 1: after every step the bodies' poses (relative to the placement x, as the kernel keeps
                           them) and velocities are rounded to float: what fp32 STORAGE of the state alone does to a
                           trajectory -- the yardstick of the fp32 short-horizon test                              */
  int32_t sleeping;     /* 1: dyn4j's at-rest detection (C.7): an island whose bodies all stayed below 0.01 u/s and 2 deg/s
                           for 0.5 s is put at rest (velocities zeroed, not simulated) until DistanceJoint.setRestDistance
                           CHANGES a rest distance, i.e. until the controller asks for a different force              */
  int32_t dyn4j_labels; /* 1: contact ids as dyn4j labels them -- the wrap-around edge of a polygon is 0 or n depending on
                           which of its end points wins "farthest vertex"; default 0: one label per edge (the rule the
                           kernel shares: when the SAT axis is the edge's own normal the two tie and dyn4j's label, hence
                           its warm start, follows the rounding)                                                       */
  int32_t touching_separated; /* the other reading of an exactly touching pair in the time-of-impact pass: 1 =
                           separated by a distance of 0 -- the advance stands, the mass is pushed linearTolerance deep in the
                           same tick along the last normal; default 0 = not separated (back out), which the kernel shares */
} vsr_oracle_opts;

/* Optional per-tick probes; any pointer may be NULL.  Row r is robot
   (robot_begin + r).  Strides are given by the caller (max over shapes). */
typedef struct vsr_oracle_probe {
  double* trajectory;   /* [robots][ticks][body_stride][6]  x,y,theta,vx,vy,omega   */
  size_t body_stride;
  double* signals;      /* [robots][ticks][voxel_stride]  force applied at that tick */
  size_t voxel_stride;
  double* readings;     /* [robots][ticks][reading_stride] sensor readings, voxel order */
  size_t reading_stride;
  int32_t* n_contacts;  /* [robots][ticks] number of contact constraints (all kinds)   */
  int32_t* pos_iters;   /* [robots][ticks] position iterations actually run          */
  double* touch;        /* [robots][ticks][voxel_stride] Touch.isTouchingGround (0/1) */
} vsr_oracle_probe;

void vsr_oracle_default_opts(vsr_oracle_opts* o);

/* Run the episode(s).  out[r] for r in [0, robot_end-robot_begin).  Returns 0 or
   a VSR_ERR_* code; message via vsr_oracle_last_error().                        */
int vsr_oracle_run(const vsr_batch_desc* desc, const vsr_oracle_opts* opts, vsr_result* out,
                   const vsr_oracle_probe* probe);

/* MultiLayerPerceptron.apply (MultiLayerPerceptron.java:168-189) on its own:
   neurons[n_layers], flat weights, input[neurons[0]] -> output[neurons[last]].  */
int vsr_oracle_mlp_apply(int activation, const int32_t* neurons, int n_layers,
                         const double* flat_weights, const double* input, double* output);

/* Aggregator window (AggregatorSensor.java:56-66): for every tick k (0-based)
   writes the 0-based index of the oldest tick still kept.                       */
int vsr_oracle_window_first(double final_t, double dt, double interval, int32_t* first, int n);

int vsr_oracle_tick_times(double final_t, double dt, double* t, int n);

const char* vsr_oracle_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
