/*
 * Panama (java.lang.foreign, JDK 22+) memory layouts of include/vsr_b200.h, field by field.
 *
 * SOURCE ONLY: no JDK exists in the build image, so this file has never been compiled there.
 * tests/test_java_layouts.py parses the structLayout(...) blocks below and checks every field's name, size and
 * offset against the ctypes mirror of the header (2dhmsr_b200/_abi.py), which tests/test_abi_cpu.py in turn checks
 * against the compiled library.  Keep the one-field-per-line form: the test reads it.
 */
package it.units.erallab.hmsrobots.b200;

import java.lang.foreign.MemoryLayout;
import java.lang.foreign.StructLayout;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

public final class VsrLayouts {
  private VsrLayouts() {
  }

  public static final int VSR_ABI_VERSION = 2;
  public static final int VSR_LIDAR_MAX_RAYS = 8;

  /* dyn4j Settings + World gravity (vsr_settings) */
  public static final StructLayout VSR_SETTINGS = MemoryLayout.structLayout(
      JAVA_DOUBLE.withName("step_frequency"),
      JAVA_INT.withName("velocity_iterations"),
      JAVA_INT.withName("position_iterations"),
      JAVA_DOUBLE.withName("linear_tolerance"),
      JAVA_DOUBLE.withName("angular_tolerance"),
      JAVA_DOUBLE.withName("max_linear_correction"),
      JAVA_DOUBLE.withName("baumgarte"),
      JAVA_DOUBLE.withName("restitution_velocity"),
      JAVA_DOUBLE.withName("max_translation"),
      JAVA_DOUBLE.withName("max_rotation"),
      JAVA_DOUBLE.withName("gravity_x"),
      JAVA_DOUBLE.withName("gravity_y"),
      JAVA_DOUBLE.withName("ground_friction"),
      JAVA_DOUBLE.withName("ground_restitution"),
      JAVA_INT.withName("spring_mass_mode"),
      JAVA_INT.withName("self_collision"),
      JAVA_INT.withName("continuous_detection"),
      JAVA_INT.withName("reserved")
  ).withName("vsr_settings");

  /* the 13 Voxel constructor fields, Voxel.java:103-118 (vsr_material) */
  public static final StructLayout VSR_MATERIAL = MemoryLayout.structLayout(
      JAVA_DOUBLE.withName("side_length"),
      JAVA_DOUBLE.withName("mass_side_length_ratio"),
      JAVA_DOUBLE.withName("spring_f"),
      JAVA_DOUBLE.withName("spring_d"),
      JAVA_DOUBLE.withName("mass_linear_damping"),
      JAVA_DOUBLE.withName("mass_angular_damping"),
      JAVA_DOUBLE.withName("friction"),
      JAVA_DOUBLE.withName("restitution"),
      JAVA_DOUBLE.withName("mass"),
      JAVA_DOUBLE.withName("area_ratio_passive_min"),
      JAVA_DOUBLE.withName("area_ratio_passive_max"),
      JAVA_DOUBLE.withName("area_ratio_active_min"),
      JAVA_DOUBLE.withName("area_ratio_active_max"),
      JAVA_INT.withName("spring_scaffoldings"),
      JAVA_INT.withName("reserved")
  ).withName("vsr_material");

  /* one [Noisy]([SoftNormalization|Normalization]([Average|Trend](base))) composition (vsr_sensor) */
  public static final StructLayout VSR_SENSOR = MemoryLayout.structLayout(
      JAVA_INT.withName("base"),
      JAVA_INT.withName("axes"),
      JAVA_INT.withName("rotated"),
      JAVA_INT.withName("aggregator"),
      JAVA_INT.withName("soft_norm"),
      JAVA_INT.withName("reserved"),
      JAVA_DOUBLE.withName("max_velocity_norm"),
      JAVA_DOUBLE.withName("interval"),
      JAVA_DOUBLE.withName("noise_sigma"),
      JAVA_LONG.withName("noise_seed"),
      MemoryLayout.sequenceLayout(8, JAVA_DOUBLE).withName("lidar_dirs")
  ).withName("vsr_sensor");

  /* a Grid<Boolean> body + its per-voxel sensor lists (vsr_shape) */
  /* one BreakableVoxel's parameters (vsr_breakable) */
  public static final StructLayout VSR_BREAKABLE = MemoryLayout.structLayout(
      JAVA_INT.withName("enabled"),
      MemoryLayout.sequenceLayout(3, JAVA_INT).withName("malfunctions"),
      MemoryLayout.sequenceLayout(3, JAVA_DOUBLE).withName("thresholds"),
      JAVA_DOUBLE.withName("restore_time"),
      JAVA_LONG.withName("random_seed")
  ).withName("vsr_breakable");

  public static final StructLayout VSR_SHAPE = MemoryLayout.structLayout(
      JAVA_INT.withName("w"),
      JAVA_INT.withName("h"),
      ADDRESS.withName("mask"),
      ADDRESS.withName("sensor_offsets"),
      ADDRESS.withName("sensors"),
      JAVA_INT.withName("n_hidden"),
      JAVA_INT.withName("reserved"),
      ADDRESS.withName("hidden"),
      ADDRESS.withName("breakable"),
      ADDRESS.withName("material")
  ).withName("vsr_shape");

  /* N robots, one terrain, one Settings (vsr_batch_desc) */
  public static final StructLayout VSR_BATCH_DESC = MemoryLayout.structLayout(
      JAVA_INT.withName("abi_version"),
      JAVA_INT.withName("precision"),
      JAVA_LONG.withName("n_robots"),
      JAVA_INT.withName("n_shapes"),
      JAVA_INT.withName("controller_kind"),
      ADDRESS.withName("shapes"),
      ADDRESS.withName("robot_shape"),
      VSR_MATERIAL.withName("material"),
      JAVA_INT.withName("mlp_activation"),
      JAVA_INT.withName("distributed_signals"),
      ADDRESS.withName("params"),
      ADDRESS.withName("param_offsets"),
      JAVA_DOUBLE.withName("final_t"),
      ADDRESS.withName("terrain_xs"),
      ADDRESS.withName("terrain_ys"),
      JAVA_INT.withName("terrain_n"),
      JAVA_INT.withName("reserved1"),
      JAVA_DOUBLE.withName("initial_placement"),
      VSR_SETTINGS.withName("settings"),
      JAVA_LONG.withName("trajectory_robots"),
      JAVA_DOUBLE.withName("t_start"),
      ADDRESS.withName("robot_placement")
  ).withName("vsr_batch_desc");

  /* what Outcome needs of the first and the last Observation (vsr_result) */
  public static final StructLayout VSR_RESULT = MemoryLayout.structLayout(
      JAVA_DOUBLE.withName("first_center_x"),
      JAVA_DOUBLE.withName("first_center_y"),
      JAVA_DOUBLE.withName("last_center_x"),
      JAVA_DOUBLE.withName("last_center_y"),
      JAVA_DOUBLE.withName("t_first"),
      JAVA_DOUBLE.withName("t_last"),
      JAVA_DOUBLE.withName("area_ratio_energy_first"),
      JAVA_DOUBLE.withName("area_ratio_energy_last"),
      JAVA_DOUBLE.withName("control_energy_first"),
      JAVA_DOUBLE.withName("control_energy_last"),
      JAVA_INT.withName("n_ticks"),
      JAVA_INT.withName("status")
  ).withName("vsr_result");

  /* enums of the header */
  public static final int SENSOR_TOUCH = 0, SENSOR_AREA_RATIO = 1, SENSOR_VELOCITY = 2, SENSOR_ANGLE = 3,
      SENSOR_CONSTANT = 4, SENSOR_APPLIED_FORCE = 5, SENSOR_MALFUNCTION = 6, SENSOR_TIME_SIN = 7,
      SENSOR_CRUMPLING = 8, SENSOR_CONTROL_POWER = 9, SENSOR_LIDAR = 10;
  public static final int NORM_NONE = 0, NORM_SOFT = 1, NORM_LINEAR = 2;
  public static final int AGG_NONE = 0, AGG_AVERAGE = 1, AGG_TREND = 2, AGG_DYNAMIC_NORM = 3;
  public static final int AXIS_X = 1, AXIS_Y = 2;
  public static final int CTRL_PHASE_SIN = 0, CTRL_TABULATED = 1, CTRL_MLP = 2, CTRL_DISTRIBUTED = 3,
      CTRL_DISTRIBUTED_ND = 4;
  public static final int PRECISION_F32 = 0, PRECISION_F64 = 1;
  public static final int SCAFFOLD_SIDE_EXTERNAL = 1, SCAFFOLD_SIDE_INTERNAL = 2, SCAFFOLD_SIDE_CROSS = 4,
      SCAFFOLD_CENTRAL_CROSS = 8;
  public static final int OK = 0, ERR_INVALID_ARGUMENT = -1, ERR_UNSUPPORTED = -2, ERR_CUDA = -3, ERR_STATE = -4,
      ERR_NO_DEVICE = -5;
  public static final int STATUS_NONFINITE = 1, STATUS_CONTACT_OVERFLOW = 2, STATUS_PAIR_OVERFLOW = 4,
      STATUS_LEVEL_OVERFLOW = 8;
}
