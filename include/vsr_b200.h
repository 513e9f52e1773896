/*
 * vsr_b200.h -- C ABI of the B200-native batched voxel-soft-robot (VSR) stepper.
 *
 * This is the drop-in boundary for ONE hot path of 2D-VSR-Sim (ericmedvet/2dhmsr):
 *   Locomotion.apply(Robot)            tasks/locomotion/Locomotion.java:168-207
 *     -> AbstractTask.updateWorld      tasks/AbstractTask.java:41-64
 *          -> World.step(1)            (org.dyn4j:dyn4j:4.2.1, pom.xml:31-35)
 *          -> Robot.act(t)             core/objects/Robot.java:73-77
 * evaluated for MANY independent robots at once on one B200.
 *
 * The reference has no FFI seam; a Java maintainer binds these entry points with
 * Panama (java.lang.foreign) or JNI from a batched `Locomotion.apply(List<Robot>)`
 * (see INTEGRATION.md).  Everything here is plain C: pointers, sizes, doubles.
 * The caller owns every host buffer; the library owns all device memory.
 * Handles are opaque and NOT thread-safe (one handle per host thread / GPU).
 * There is no CPU fallback: every entry point that computes needs a CUDA device.
 *
 * All `double` inputs mirror the reference's Java doubles; the device computes in
 * fp32 (VSR_PRECISION_F32, the product path) or fp64 (VSR_PRECISION_F64, parity
 * build of the very same kernel source).
 */
#ifndef VSR_B200_H
#define VSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSR_ABI_VERSION 2

/* ---- error codes (reference convention: unchecked IllegalArgumentException at
 *      validation time, e.g. Voxel.java:132-140, Ground.java:46-56,
 *      CentralizedSensing.java:121-127; the Java side maps <0 to exceptions) ---- */
#define VSR_OK 0
#define VSR_ERR_INVALID_ARGUMENT (-1) /* -> IllegalArgumentException */
#define VSR_ERR_UNSUPPORTED (-2)      /* valid in the reference, not on this path yet */
#define VSR_ERR_CUDA (-3)             /* -> IllegalStateException */
#define VSR_ERR_STATE (-4)            /* call order (e.g. results before run) */
#define VSR_ERR_NO_DEVICE (-5)        /* no CUDA device: there is NO CPU fallback */

/* ---- dyn4j Settings + World gravity (Locomotion.java:50; AbstractTask.java:35).
 *      Defaults = dyn4j 4.2.1 `new Settings()` / `new World<>()`.                ---- */
typedef struct vsr_settings {
  double step_frequency;        /* 1/60  Settings.getStepFrequency(), Locomotion.java:194,197 */
  int32_t velocity_iterations;  /* 10 */
  int32_t position_iterations;  /* 10 */
  double linear_tolerance;      /* 0.005 */
  double angular_tolerance;     /* 2 deg in rad */
  double max_linear_correction; /* 0.2 */
  double baumgarte;             /* 0.2 */
  double restitution_velocity;  /* 1.0 */
  double max_translation;       /* 2.0 */
  double max_rotation;          /* pi/2 */
  double gravity_x;             /* 0 */
  double gravity_y;             /* -9.8 */
  double ground_friction;       /* 0.2  BodyFixture default (Ground.java:72 passes none) */
  double ground_restitution;    /* 0.0 */
  int32_t spring_mass_mode;     /* VSR_SPRING_MASS_* : which mass turns (f, zeta) into (k, d) */
  int32_t self_collision;       /* 1 (default): masses of one robot collide with each other unless welded
                                   (Voxel.java:178-185,474; Robot.java:66-71); 0: ground contacts only */
  int32_t continuous_detection; /* World.solveTOI for the masses against the ground polygons:
                                   1 (default) = dyn4j's ContinuousDetectionMode.ALL, what `new Settings()` gives: after
                                   the position solve a mass about to touch a polygon it is not yet in contact with is
                                   put back to its time of impact and pushed linearTolerance into it; 0 = NONE (and
                                   BULLETS_ONLY: the reference creates no bullets).  The pass lowers the mean 20 s
                                   displacement of a population by 7-13 %. */
  int32_t reserved;
} vsr_settings;

#define VSR_SPRING_MASS_REDUCED 0   /* m1*m2/(m1+m2)  (Box2D-2.4 lineage)            */
#define VSR_SPRING_MASS_EFFECTIVE 1 /* 1/(J M^-1 J^T) (dyn4j 3.x/4.0 formula)         */

/* ---- Voxel material: the 13-argument constructor, Voxel.java:103-118 ---- */
typedef struct vsr_material {
  double side_length;             /* 3.0   Voxel.java:57 */
  double mass_side_length_ratio;  /* 0.35  :58 */
  double spring_f;                /* 8.0   :59 */
  double spring_d;                /* 0.3   :60 */
  double mass_linear_damping;     /* 0.1   :61 */
  double mass_angular_damping;    /* 0.1   :62 */
  double friction;                /* 10.0  :63 */
  double restitution;             /* 0.1   :64 */
  double mass;                    /* 1.0   :65 */
  double area_ratio_passive_min;  /* -inf  :66; a finite bound becomes the lower / upper limit of every spring's   */
  double area_ratio_passive_max;  /* +inf      DistanceJoint (the "rope" of Voxel.java:262-285,331-338)              */
  double area_ratio_active_min;   /* 0.8   :67 */
  double area_ratio_active_max;   /* 1.2 */
  uint32_t spring_scaffoldings;   /* bitmask of VSR_SCAFFOLD_*; :68 = all four */
  uint32_t reserved;
} vsr_material;

#define VSR_SCAFFOLD_SIDE_EXTERNAL 1u
#define VSR_SCAFFOLD_SIDE_INTERNAL 2u
#define VSR_SCAFFOLD_SIDE_CROSS 4u
#define VSR_SCAFFOLD_CENTRAL_CROSS 8u
#define VSR_SCAFFOLD_ALL 15u

/* ---- Sensors: every entry of RobotUtils.PREDEFINED_SENSORS is
 *      [SoftNormalization|Normalization](  [Average|Trend]( base sensor ) )
 *      (RobotUtils.java:38-57).  One vsr_sensor encodes one such composition.     ---- */
#define VSR_SENSOR_TOUCH 0      /* sensors/Touch.java:35-65, 1 reading, domain [0,1]   */
#define VSR_SENSOR_AREA_RATIO 1 /* sensors/AreaRatio.java:21-37, domain [0.5,1.5]      */
#define VSR_SENSOR_VELOCITY 2   /* sensors/Velocity.java:57-80, 1-2 readings           */
#define VSR_SENSOR_ANGLE 3      /* sensors/Angle.java: Voxel.getAngle, domain [-pi,pi]   */
#define VSR_SENSOR_CONSTANT 4   /* sensors/Constant.java (one value = `param`; "px", "py"),
                                   domain [min(0, v), max(1, v)]                        */
#define VSR_SENSOR_APPLIED_FORCE 5 /* sensors/AppliedForce.java, domain [-1,1]          */
#define VSR_SENSOR_MALFUNCTION 6   /* sensors/Malfunction.java on a non-breakable voxel: 0 */
#define VSR_SENSOR_TIME_SIN 7   /* sensors/TimeFunction.java with t -> sin(2 pi f t), f = `param`,
                                   domain [-1,1] (RobotUtils "cpg": f = -1)             */

#define VSR_SENSOR_CRUMPLING 8  /* sensors/Crumpling.java: share of the 6 mass pairs closer than 0.2 side lengths */
#define VSR_SENSOR_CONTROL_POWER 9 /* sensors/ControlPower.java: Voxel.getControlEnergy / (t - previous t),
                                      domain [0, 1 / `param`] (param = controlInterval)                  */

#define VSR_SENSOR_LIDAR 10 /* sensors/Lidar.java: `axes` rays of length `param` from the voxel centre at
                               lidar_dirs[i] + Voxel.getAngle(); a reading is the distance to the terrain
                               (robot parts are filtered out, :37-46) or the ray length                  */
#define VSR_LIDAR_MAX_RAYS 8

#define VSR_NORM_NONE 0
#define VSR_NORM_SOFT 1   /* sensors/SoftNormalization.java:38-48 */
#define VSR_NORM_LINEAR 2 /* sensors/Normalization.java: DoubleRange.normalize of the inner domain */

#define VSR_AGG_NONE 0
#define VSR_AGG_AVERAGE 1 /* sensors/Average.java:33-45 */
#define VSR_AGG_TREND 2   /* sensors/Trend.java:28-50   */
#define VSR_AGG_DYNAMIC_NORM 3 /* sensors/DynamicNormalization.java:26-61: (reading - min) / (max - min) over the kept
                                  window, clamped to [0, 1]; 0/0 = NaN while the window holds one distinct value
                                  (always at the first tick), exactly as the reference computes it; domain [0, 1] */

#define VSR_AXIS_X 1
#define VSR_AXIS_Y 2

typedef struct vsr_sensor {
  int32_t base;             /* VSR_SENSOR_*                                   */
  int32_t axes;             /* VELOCITY: VSR_AXIS_X | VSR_AXIS_Y              */
  int32_t rotated;          /* VELOCITY: project on the voxel's own axes      */
  int32_t aggregator;       /* VSR_AGG_*                                      */
  int32_t soft_norm;        /* VSR_NORM_*: the outermost wrapper              */
  int32_t reserved;
  double max_velocity_norm; /* VELOCITY domain [-m, m]; CONSTANT / TIME_SIN: `param` */
  double interval;          /* aggregator window, seconds (0.25 / 0.5)        */
  double noise_sigma;       /* > 0: wrapped in sensors/Noisy.java (outermost): every reading gets
                               java.util.Random(noise_seed).nextGaussian() * sigma * domain extent */
  int64_t noise_seed;       /* one seed per batch (RobotUtils.java:148 passes 0)                  */
  double lidar_dirs[VSR_LIDAR_MAX_RAYS]; /* LIDAR: rayDirections                                  */
} vsr_sensor;

/* ---- BreakableVoxel (core/objects/BreakableVoxel.java): a voxel whose actuator, sensors or structure may
 *      malfunction.  After Voxel.act the voxel accumulates its trigger counters (time, control energy, area-ratio
 *      energy since the last break); for every entry of triggerThresholds it draws random.nextDouble() and breaks when
 *      the draw is < 1 - tanh(threshold / counter) (:142-158): a component and one of its malfunction types are drawn
 *      with random.nextInt; everything is restored restoreTime after the last break (:163-167).  ACTUATOR: ZERO ->
 *      applyForce(0), FROZEN -> the last applied force again, RANDOM -> U(-1, 1) (:170-180).  SENSORS: ZERO -> zeros,
 *      FROZEN -> the readings of the tick it broke, RANDOM -> uniform in each reading's domain, drawn when a
 *      controller asks for the readings (:183-193).  STRUCTURE: FROZEN -> the voxel's springs become rigid distance
 *      joints (setFrequency(0), :249-257; re-applied only when the voxel breaks again, as there); other structure
 *      types throw in the reference and are rejected here.  The reference iterates its Maps and Sets; this interface
 *      fixes the enum order (CONTROL, AREA, TIME; ACTUATOR, SENSORS, STRUCTURE; NONE, ZERO, FROZEN, RANDOM), which is
 *      what an EnumMap / EnumSet gives (a Map.of with several entries has no defined order in the reference either). */
#define VSR_MALFUNCTION_NONE 0
#define VSR_MALFUNCTION_ZERO 1
#define VSR_MALFUNCTION_FROZEN 2
#define VSR_MALFUNCTION_RANDOM 3
#define VSR_COMPONENT_ACTUATOR 0
#define VSR_COMPONENT_SENSORS 1
#define VSR_COMPONENT_STRUCTURE 2
#define VSR_TRIGGER_CONTROL 0
#define VSR_TRIGGER_AREA 1
#define VSR_TRIGGER_TIME 2
typedef struct vsr_breakable {
  int32_t enabled;           /* 0: a plain Voxel (the other fields are ignored)                                  */
  uint32_t malfunctions[3];  /* per VSR_COMPONENT_*: bit t set = VSR_MALFUNCTION_t is in the component's set;
                                0 = the component is not a key of the `malfunctions` map                        */
  double thresholds[3];      /* per VSR_TRIGGER_*; NaN = the trigger is not a key of `triggerThresholds`         */
  double restore_time;       /* may be +inf                                                                     */
  int64_t random_seed;
} vsr_breakable;

/* ---- Shape class: a Grid<Boolean> body + its per-voxel sensor lists.
 *      Voxels are enumerated in Grid.values() order, i.e. row-major y*w+x with
 *      nulls skipped (Grid.java:180-188,268-270).                              ---- */
typedef struct vsr_shape {
  int32_t w, h;
  const uint8_t* mask;            /* [w*h], index y*w+x, non-zero = voxel present       */
  const int32_t* sensor_offsets;  /* [n_voxels+1] into `sensors`, per voxel in order    */
  const vsr_sensor* sensors;      /* concatenated per-voxel sensor lists                */
  /* MLP controllers only: hidden layer sizes for robots of this shape             */
  int32_t n_hidden;
  int32_t reserved;
  const int32_t* hidden;          /* [n_hidden]                                         */
  const vsr_breakable* breakable; /* [n_voxels] in voxel order, or NULL: no BreakableVoxel in this shape class */
  const struct vsr_material* material; /* the 13 Voxel constructor fields of this class' voxels, or NULL: the batch's
                                          vsr_batch_desc.material.  One material per ROBOT (the rows between voxels
                                          are solved in closed form for equal masses); robots of different materials
                                          go to different shape classes of one batch.                             */
} vsr_shape;

/* ---- Controllers (one kind per batch; parameters per robot) ---- */
#define VSR_CTRL_PHASE_SIN 0 /* controllers/PhaseSin.java:61: sin(2*pi*f*t + phi)*A.
                                params = [f, A, phi_0 .. phi_{nv-1}]                       */
#define VSR_CTRL_TABULATED 1 /* any TimeFunctions (TimeFunctions.java:49-64): the host
                                samples each voxel's lambda at the tick times.
                                params = signal[tick][voxel], n_ticks*nv values            */
#define VSR_CTRL_MLP 2       /* CentralizedSensing.java:85-114 + MultiLayerPerceptron
                                .java:168-189. params = flat weights in
                                MultiLayerPerceptron.flat order (:139-151), bias first     */

#define VSR_CTRL_DISTRIBUTED 3 /* DistributedSensing.java:150-207 with one MultiLayerPerceptron per
                                  voxel (Starter.java:176-190): inputs = the voxel's sensor readings
                                  then `distributed_signals` values from each of the N, E, S, W
                                  neighbours' previous outputs; outputs = control + 4 * signals.
                                  All voxels share the hidden layer sizes (vsr_shape.hidden).
                                  params = the voxels' flat weights (MultiLayerPerceptron.flat
                                  order), concatenated in Grid.values() order                  */

#define VSR_CTRL_DISTRIBUTED_ND 4 /* DistributedSensingNonDirectional.java:57-78: outputs = control +
                                     `distributed_signals` values, the same for all four neighbours   */

#define VSR_ACT_RELU 0 /* MultiLayerPerceptron.java:89-95, enum order */
#define VSR_ACT_SIGMOID 1
#define VSR_ACT_SIN 2
#define VSR_ACT_TANH 3
#define VSR_ACT_SIGN 4
#define VSR_ACT_IDENTITY 5

#define VSR_PRECISION_F32 0
#define VSR_PRECISION_F64 1

#define VSR_FLAG_ORDER_CANONICAL 0 /* graph-coloured Gauss-Seidel   */

/* ---- The batch: N robots, one terrain, one Settings (one Locomotion task) ---- */
typedef struct vsr_batch_desc {
  uint32_t abi_version;  /* VSR_ABI_VERSION */
  uint32_t precision;    /* VSR_PRECISION_*  */
  size_t n_robots;
  /* shape classes */
  int32_t n_shapes;
  int32_t controller_kind; /* VSR_CTRL_* */
  const vsr_shape* shapes;
  const int32_t* robot_shape; /* [n_robots] index into shapes (NULL => all 0) */
  vsr_material material;      /* one material per batch (default Voxel) */
  /* controller parameters, ragged per robot */
  int32_t mlp_activation;      /* VSR_ACT_* */
  int32_t distributed_signals; /* VSR_CTRL_DISTRIBUTED: signals per side (DistributedSensing.java:50) */
  const double* params;        /* concatenated per-robot parameter vectors */
  const size_t* param_offsets; /* [n_robots+1] */
  /* Locomotion task: Locomotion.java:50-59 */
  double final_t;              /* 20: loop `while (t < finalT)` with t += dt, :196 */
  const double* terrain_xs;    /* groundProfile[0] */
  const double* terrain_ys;    /* groundProfile[1] */
  int32_t terrain_n;
  int32_t reserved1;
  double initial_placement;    /* xs[1] + 1 by default, :51 */
  vsr_settings settings;
  /* optional per-tick body dump for the first `trajectory_robots` robots
     (SnapshotListener / Observation path, Locomotion.java:198-202) */
  size_t trajectory_robots;
  /* stages of a longer task (DevoLocomotion.java): the episode's clock starts at t_start (ticks
     t_start + dT, ... while t < final_t; 0 for a plain Locomotion) and every robot may have its
     own initial placement (NULL: initial_placement for all)                                     */
  double t_start;
  const double* robot_placement; /* [n_robots] or NULL */
} vsr_batch_desc;

/* ---- Per-robot result: what Outcome needs from the first and the last
 *      Observation (Outcome.java:42-58,126-142,152-200).                     ---- */
typedef struct vsr_result {
  double first_center_x, first_center_y; /* BehaviorUtils.center at t = dt   */
  double last_center_x, last_center_y;   /* ... at the last tick             */
  double t_first, t_last;                /* observation keys (accumulated t) */
  double area_ratio_energy_first, area_ratio_energy_last; /* sum over voxels  */
  double control_energy_first, control_energy_last;
  int32_t n_ticks;
  int32_t status; /* 0 = ok; else a bit set (VSR_STATUS_*): the record is NOT a valid Outcome */
} vsr_result;

/* vsr_result.status bits.  The reference (dyn4j) has no contact capacity; the device keeps fixed-size
   per-mass / per-voxel constraint lists, and a robot that ever needs more is FLAGGED, never silently
   truncated: callers must treat a non-zero status as a failed episode (the host mirror raises).        */
#define VSR_STATUS_NONFINITE 1        /* a non-finite centre was observed                                 */
#define VSR_STATUS_CONTACT_OVERFLOW 2 /* a mass touched more ground polygons at once than the device keeps */
#define VSR_STATUS_PAIR_OVERFLOW 4    /* a voxel had more mass-pair candidates in one tick than it keeps   */
#define VSR_STATUS_LEVEL_OVERFLOW 8   /* a chain of dependent contact constraints exceeded 32 levels       */

typedef struct vsr_batch vsr_batch;

/* Fill the dyn4j / Voxel defaults. */
void vsr_default_settings(vsr_settings* out);
void vsr_default_material(vsr_material* out);

/* Number of ticks `while (t < finalT) t += dt` runs (Java double accumulation,
   Locomotion.java:195-203) and the tick times themselves. */
int vsr_tick_count(double final_t, double dt);
int vsr_tick_count_from(double t_start, double final_t, double dt);

/* Validate + compile the descriptor (host only; copies everything it needs).
   Replaces: robot.reset() + world assembly, Locomotion.java:172-191.          */
int vsr_batch_create(const vsr_batch_desc* desc, vsr_batch** out);

/* Copy the compiled batch to `device` (cudaMemcpyAsync on `stream`, a cudaStream_t
   passed as void*, NULL = legacy default stream). */
int vsr_batch_upload(vsr_batch* b, int device, void* stream);

/* Enqueue the whole episode (all ticks of every robot) on `stream`.
   Replaces: the `while (t < finalT)` loop, Locomotion.java:196-203.            */
int vsr_batch_run(vsr_batch* b, void* stream);

/* Wait for the run and copy the per-robot records to `out[n]` (n = n_robots).
   Replaces: new Outcome(observations), Locomotion.java:206.                    */
int vsr_batch_results(vsr_batch* b, vsr_result* out, size_t n);

/* Device pointer to the vsr_result[n_robots] array (the buffer a final NCCL
   gather sends from without a staging copy). */
int vsr_batch_results_device(vsr_batch* b, void** dptr, size_t* bytes);

/* Per-tick body states of robot r (< trajectory_robots):
   out[tick][body][6] = x, y, theta, vx, vy, omega (world frame), bodies in
   world-insertion order (voxel order, then NW, NE, SE, SW).  Returns the number
   of doubles written, or <0.                                                   */
long long vsr_batch_trajectory(vsr_batch* b, size_t robot, double* out, size_t capacity);

/* Robot.boundingBox() (Robot.java:116-120) of every robot at the last tick: out[robot][4] = min x, min y,
   max x, max y (world frame).  The next stage of a DevoLocomotion task places the developed robot at
   this min x (DevoLocomotion.java: rebuildWorld).                                                   */
int vsr_batch_final_bbox(vsr_batch* b, double* out, size_t n);

/* What a VoxelPoly carries besides its geometry (Voxel.java:605-616), per tick, for robot r
   (< trajectory_robots): out[tick][voxel][2] = Touch.isTouchingGround (0/1),
   Voxel.getLastAppliedForce, voxels in Grid.values() order.  With vsr_batch_trajectory this is
   all an Outcome.Observation holds (Locomotion.java:198-202).  Returns the number of doubles
   written, or <0.                                                                          */
long long vsr_batch_voxel_trace(vsr_batch* b, size_t robot, double* out, size_t capacity);

/* Voxel.getSensorReadings (Voxel.java:530-536) of every voxel, per tick, for robot r (< trajectory_robots):
   out[tick][n_readings], the voxels' readings concatenated in Grid.values() order -- the input vector of
   CentralizedSensing (CentralizedSensing.java:98-104).  Returns the number of doubles written, or <0.   */
long long vsr_batch_sensor_readings(vsr_batch* b, size_t robot, double* out, size_t capacity);

/* Facts about the compiled batch. */
size_t vsr_batch_voxel_steps(const vsr_batch* b); /* sum over robots of voxels*ticks */
int vsr_batch_n_ticks(const vsr_batch* b);
int vsr_batch_n_bodies(const vsr_batch* b, size_t robot);
/* kernels launched by the last vsr_batch_run (for bench.py's gpu_launches) */
int vsr_batch_last_launches(const vsr_batch* b);
/* host->device bytes copied by the last vsr_batch_upload */
size_t vsr_batch_upload_bytes(const vsr_batch* b);

void vsr_batch_destroy(vsr_batch* b);

/* Thread-local message for the last non-zero return. */
const char* vsr_last_error(void);

/* Library facts. */
int vsr_abi_version(void);
int vsr_device_count(void); /* 0 if no CUDA device / driver */

#ifdef __cplusplus
}
#endif
#endif /* VSR_B200_H */
