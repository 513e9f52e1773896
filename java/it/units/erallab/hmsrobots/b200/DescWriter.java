/*
 * Robot / Voxel / Sensor / Controller objects -> vsr_batch_desc (include/vsr_b200.h) in native memory.
 *
 * The reference keeps most of what the device needs in private final fields without getters (Velocity.rotated,
 * AggregatorSensor.interval, Noisy.sigma, the 13 Voxel constructor fields ...), but every one of them is a
 * @JsonProperty: the reference's own persistence format (util/SerializationUtils.java:25-38: an ObjectMapper that
 * sees annotated fields only) is therefore the schema this writer reads -- no reflection, no new getters.
 * Controllers' numeric payloads come from the public getters (PhaseSin.getPhases, MultiLayerPerceptron.getParams /
 * getNeurons, TimeFunctions.getFunctions, DistributedSensing.getFunctions).
 *
 * SOURCE ONLY: no JDK in the build image, never compiled there.  The same flattening, line for line, is what
 * 2dhmsr_b200/hmsrobots.py (compile_batch, Sensor.encode) does and the GPU parity tests exercise.
 */
package it.units.erallab.hmsrobots.b200;

import com.fasterxml.jackson.annotation.JsonAutoDetect;
import com.fasterxml.jackson.annotation.PropertyAccessor;
import com.fasterxml.jackson.databind.JsonNode;
import com.fasterxml.jackson.databind.ObjectMapper;
import com.fasterxml.jackson.databind.SerializationFeature;
import it.units.erallab.hmsrobots.core.controllers.CentralizedSensing;
import it.units.erallab.hmsrobots.core.controllers.Controller;
import it.units.erallab.hmsrobots.core.controllers.DistributedSensing;
import it.units.erallab.hmsrobots.core.controllers.DistributedSensingNonDirectional;
import it.units.erallab.hmsrobots.core.controllers.MultiLayerPerceptron;
import it.units.erallab.hmsrobots.core.controllers.PhaseSin;
import it.units.erallab.hmsrobots.core.controllers.TimeFunctions;
import it.units.erallab.hmsrobots.core.controllers.TimedRealFunction;
import it.units.erallab.hmsrobots.core.objects.Robot;
import it.units.erallab.hmsrobots.core.objects.Voxel;
import it.units.erallab.hmsrobots.util.Grid;

import java.lang.foreign.Arena;
import java.lang.foreign.MemoryLayout.PathElement;
import java.lang.foreign.MemorySegment;
import java.util.ArrayList;
import java.util.Arrays;
import java.util.HashMap;
import java.util.List;
import java.util.Map;

import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

final class DescWriter {
  private static final ObjectMapper OM = new ObjectMapper();

  static {
    OM.disable(SerializationFeature.FAIL_ON_EMPTY_BEANS);
    OM.setVisibility(PropertyAccessor.FIELD, JsonAutoDetect.Visibility.NONE);
    OM.setVisibility(PropertyAccessor.GETTER, JsonAutoDetect.Visibility.NONE);
  }

  private DescWriter() {
  }

  private static final String S = "it.units.erallab.hmsrobots.core.sensors.";

  /** one flattened [Noisy]([SoftNormalization|Normalization]([Average|Trend](base))) */
  private record FlatSensor(int base, int axes, int rotated, int aggregator, int norm, double param, double interval,
                            double noiseSigma, long noiseSeed, double[] lidarDirs) {
  }

  /** RobotUtils.java:38-60 builds exactly these nestings; anything else is outside the accelerated path */
  private static FlatSensor flatten(JsonNode s) {
    double sigma = 0d;
    long seed = 0L;
    int norm = VsrLayouts.NORM_NONE, agg = VsrLayouts.AGG_NONE;
    double interval = 0d;
    String cls = s.get("@class").asText();
    if (cls.equals(S + "Noisy")) {                       // Noisy.java:40-47 (outermost: RobotUtils.java:172-175)
      sigma = s.get("sigma").asDouble();
      seed = s.get("seed").asLong();
      s = s.get("sensor");
      cls = s.get("@class").asText();
    }
    if (cls.equals(S + "SoftNormalization") || cls.equals(S + "Normalization")) {
      norm = cls.endsWith("SoftNormalization") ? VsrLayouts.NORM_SOFT : VsrLayouts.NORM_LINEAR;
      s = s.get("sensor");
      cls = s.get("@class").asText();
    }
    if (cls.equals(S + "Average") || cls.equals(S + "Trend") || cls.equals(S + "DynamicNormalization")) {   // AggregatorSensor.java:29 interval
      agg = cls.endsWith("Average") ? VsrLayouts.AGG_AVERAGE : (cls.endsWith("Trend") ? VsrLayouts.AGG_TREND : VsrLayouts.AGG_DYNAMIC_NORM);
      interval = s.get("interval").asDouble();
      s = s.get("sensor");
      cls = s.get("@class").asText();
    }
    String name = cls.substring(S.length());
    double[] dirs = new double[VsrLayouts.VSR_LIDAR_MAX_RAYS];
    switch (name) {
      case "Touch":
        return new FlatSensor(VsrLayouts.SENSOR_TOUCH, 0, 0, agg, norm, 0d, interval, sigma, seed, dirs);
      case "AreaRatio":
        return new FlatSensor(VsrLayouts.SENSOR_AREA_RATIO, 0, 0, agg, norm, 0d, interval, sigma, seed, dirs);
      case "Angle":
        return new FlatSensor(VsrLayouts.SENSOR_ANGLE, 0, 0, agg, norm, 0d, interval, sigma, seed, dirs);
      case "AppliedForce":
        return new FlatSensor(VsrLayouts.SENSOR_APPLIED_FORCE, 0, 0, agg, norm, 0d, interval, sigma, seed, dirs);
      case "Malfunction":
        return new FlatSensor(VsrLayouts.SENSOR_MALFUNCTION, 0, 0, agg, norm, 0d, interval, sigma, seed, dirs);
      case "Crumpling":
        return new FlatSensor(VsrLayouts.SENSOR_CRUMPLING, 0, 0, agg, norm, 0d, interval, sigma, seed, dirs);
      case "ControlPower":
        return new FlatSensor(VsrLayouts.SENSOR_CONTROL_POWER, 0, 0, agg, norm, s.get("controlInterval").asDouble(),
            interval, sigma, seed, dirs);
      case "Constant": {                                  // Constant.java:35: one value per instance here ("px", "py")
        JsonNode v = s.get("values");
        if (v.size() != 1) {
          throw new UnsupportedOperationException("Constant sensors with " + v.size() + " values");
        }
        return new FlatSensor(VsrLayouts.SENSOR_CONSTANT, 0, 0, agg, norm, v.get(0).asDouble(), interval, sigma, seed, dirs);
      }
      case "Velocity": {                                  // Velocity.java:28-33
        int axes = 0;
        for (JsonNode a : s.get("axes")) {
          axes |= a.asText().equals("X") ? VsrLayouts.AXIS_X : VsrLayouts.AXIS_Y;
        }
        return new FlatSensor(VsrLayouts.SENSOR_VELOCITY, axes, s.get("rotated").asBoolean() ? 1 : 0, agg, norm,
            s.get("maxVelocityNorm").asDouble(), interval, sigma, seed, dirs);
      }
      case "Lidar": {                                     // Lidar.java:50-58
        JsonNode d = s.get("rayDirections");
        if (d.size() > VsrLayouts.VSR_LIDAR_MAX_RAYS) {
          throw new UnsupportedOperationException("Lidar with more than " + VsrLayouts.VSR_LIDAR_MAX_RAYS + " rays");
        }
        for (int i = 0; i < d.size(); i++) {
          dirs[i] = d.get(i).asDouble();
        }
        return new FlatSensor(VsrLayouts.SENSOR_LIDAR, d.size(), 0, agg, norm, s.get("rayLength").asDouble(), interval,
            sigma, seed, dirs);
      }
      default:
        // TimeFunction holds an opaque lambda (only RobotUtils' "cpg" = sin(2 pi (-1) t) is parametric);
        // DynamicNormalization and anything user-defined is not on the accelerated path
        throw new UnsupportedOperationException("sensor " + cls + " is outside the accelerated path");
    }
  }

  /** a shape class = occupancy + per-voxel flattened sensor lists (+ the MLP hidden layers of its controller) */
  private record ShapeKey(int w, int h, List<Boolean> mask, List<List<FlatSensor>> sensors, List<Integer> hidden,
                          List<List<Double>> breakable, JsonNode material) {
  }

  private static final List<String> COMPONENTS = List.of("ACTUATOR", "SENSORS", "STRUCTURE");   // BreakableVoxel.ComponentType
  private static final List<String> TRIGGERS = List.of("CONTROL", "AREA", "TIME");               // MalfunctionTrigger
  private static final List<String> TYPES = List.of("NONE", "ZERO", "FROZEN", "RANDOM");         // MalfunctionType

  /** vsr_breakable of one voxel as [enabled, mask x 3, threshold x 3, restoreTime, seed]; a plain Voxel is all zeros */
  private static List<Double> breakableOf(JsonNode vj) {
    List<Double> q = new ArrayList<>(java.util.Collections.nCopies(9, 0d));
    if (!vj.has("malfunctions")) {
      return q;
    }
    q.set(0, 1d);
    for (int c = 0; c < 3; c++) {
      int mask = 0;
      JsonNode set = vj.get("malfunctions").get(COMPONENTS.get(c));
      if (set != null) {
        for (JsonNode t : set) {
          mask |= 1 << TYPES.indexOf(t.asText());
        }
      }
      q.set(1 + c, (double) mask);
      JsonNode thr = vj.get("triggerThresholds").get(TRIGGERS.get(c));
      q.set(4 + c, thr == null ? Double.NaN : thr.asDouble());
    }
    q.set(7, Double.parseDouble(vj.get("restoreTime").asText()));
    q.set(8, (double) vj.get("randomSeed").asLong());   // (seeds come from Random.nextInt: exact in a double)
    return q;
  }

  private static int[] hiddenOf(MultiLayerPerceptron mlp) {
    int[] n = mlp.getNeurons();
    return Arrays.copyOfRange(n, 1, n.length - 1);
  }

  /** the tick times of `while (t < finalT) t += dT` (Locomotion.java:195-203, AbstractTask.java:48) */
  static double[] tickTimes(double finalT, double dt) {
    List<Double> ts = new ArrayList<>();
    double t = 0d;
    while (t < finalT) {
      t = t + dt;
      ts.add(t);
    }
    return ts.stream().mapToDouble(d -> d).toArray();
  }

  static MemorySegment write(Arena a, LocomotionSpec task, List<Robot> robots, int precision) {
    if (robots.isEmpty()) {
      throw new IllegalArgumentException("empty batch");
    }
    final double dt = task.settings().getStepFrequency();
    final double[] ticks = tickTimes(task.finalT(), dt);
    // ---- controller kind (one per batch) and per-robot parameter vectors
    Controller c0 = robots.get(0).getController();
    final int kind;
    if (c0 instanceof PhaseSin) {
      kind = VsrLayouts.CTRL_PHASE_SIN;
    } else if (c0 instanceof TimeFunctions) {
      kind = VsrLayouts.CTRL_TABULATED;
    } else if (c0 instanceof CentralizedSensing) {
      kind = VsrLayouts.CTRL_MLP;
    } else if (c0 instanceof DistributedSensingNonDirectional) {
      kind = VsrLayouts.CTRL_DISTRIBUTED_ND;
    } else if (c0 instanceof DistributedSensing) {
      kind = VsrLayouts.CTRL_DISTRIBUTED;
    } else {
      throw new UnsupportedOperationException("controller " + c0.getClass().getName() + " is outside the accelerated path");
    }
    int activation = 0, signals = 0;
    JsonNode material = null;
    List<double[]> params = new ArrayList<>(robots.size());
    Map<ShapeKey, Integer> shapeIndex = new HashMap<>();
    List<ShapeKey> shapes = new ArrayList<>();
    int[] robotShape = new int[robots.size()];
    for (int r = 0; r < robots.size(); r++) {
      Robot robot = robots.get(r);
      Grid<Voxel> body = robot.getVoxels();
      Controller c = robot.getController();
      if (!c.getClass().equals(c0.getClass())) {
        throw new UnsupportedOperationException("one controller kind per batch");
      }
      // ---- voxels in values() order = y * w + x, nulls skipped (Grid.java:180-188,268-270)
      List<Boolean> mask = new ArrayList<>();
      List<List<FlatSensor>> sensors = new ArrayList<>();
      List<int[]> cells = new ArrayList<>();
      List<List<Double>> breakable = new ArrayList<>();
      JsonNode robotMaterial = null;   // one material per robot; robots of different materials -> different shape classes
      for (int y = 0; y < body.getH(); y++) {
        for (int x = 0; x < body.getW(); x++) {
          Voxel v = body.get(x, y);
          mask.add(v != null);
          if (v == null) {
            continue;
          }
          cells.add(new int[]{x, y});
          JsonNode vj = OM.valueToTree(v);
          JsonNode m = vj.deepCopy();
          ((com.fasterxml.jackson.databind.node.ObjectNode) m).remove(List.of("sensors", "malfunctions", "triggerThresholds", "restoreTime", "randomSeed"));
          breakable.add(breakableOf(vj));   // BreakableVoxel.java:34-41 (@JsonProperty fields)
          if (material == null) {
            material = m;                // the batch's default material: the first voxel's
          }
          if (robotMaterial == null) {
            robotMaterial = m;
          } else if (!robotMaterial.equals(m)) {
            throw new UnsupportedOperationException("one voxel material per robot (include/vsr_b200.h: vsr_shape.material)");
          }
          List<FlatSensor> fs = new ArrayList<>();
          for (JsonNode s : vj.get("sensors")) {
            fs.add(flatten(s));
          }
          sensors.add(fs);
        }
      }
      // ---- parameters
      List<Integer> hidden = List.of();
      switch (kind) {
        case VsrLayouts.CTRL_PHASE_SIN -> {             // PhaseSin.java:61: sin(2 pi f t + phi) A
          JsonNode pj = OM.valueToTree(c);
          Grid<Double> ph = ((PhaseSin) c).getPhases();
          double[] p = new double[2 + cells.size()];
          p[0] = pj.get("frequency").asDouble();
          p[1] = pj.get("amplitude").asDouble();
          for (int i = 0; i < cells.size(); i++) {
            p[2 + i] = ph.get(cells.get(i)[0], cells.get(i)[1]);
          }
          params.add(p);
        }
        case VsrLayouts.CTRL_TABULATED -> {             // TimeFunctions.java:49-64: opaque lambdas, sampled at the tick times
          var fs = ((TimeFunctions) c).getFunctions();
          double[] p = new double[ticks.length * cells.size()];
          for (int k = 0; k < ticks.length; k++) {
            for (int i = 0; i < cells.size(); i++) {
              p[k * cells.size() + i] = fs.get(cells.get(i)[0], cells.get(i)[1]).apply(ticks[k]);
            }
          }
          params.add(p);
        }
        case VsrLayouts.CTRL_MLP -> {                   // CentralizedSensing.java:85-114
          TimedRealFunction f = ((CentralizedSensing) c).getFunction();
          if (!(f instanceof MultiLayerPerceptron mlp)) {
            throw new UnsupportedOperationException("CentralizedSensing over " + f.getClass().getName());
          }
          activation = activationOf(mlp);
          hidden = Arrays.stream(hiddenOf(mlp)).boxed().toList();
          params.add(mlp.getParams());                    // MultiLayerPerceptron.flat order, bias first (:139-151)
        }
        default -> {                                    // DistributedSensing.java:150-207, one MLP per voxel
          DistributedSensing ds = (DistributedSensing) c;
          signals = OM.valueToTree(ds).get("signals").asInt();
          List<double[]> parts = new ArrayList<>();
          int total = 0;
          for (int[] cell : cells) {
            TimedRealFunction f = ds.getFunctions().get(cell[0], cell[1]);
            if (f instanceof DistributedSensing.FunctionWrapper) {
              f = OM.convertValue(OM.valueToTree(f).get("inner"), TimedRealFunction.class);
            }
            if (!(f instanceof MultiLayerPerceptron mlp)) {
              throw new UnsupportedOperationException("DistributedSensing over " + f.getClass().getName());
            }
            activation = activationOf(mlp);
            hidden = Arrays.stream(hiddenOf(mlp)).boxed().toList();
            parts.add(mlp.getParams());
            total += mlp.getParams().length;
          }
          double[] p = new double[total];
          int o = 0;
          for (double[] q : parts) {
            System.arraycopy(q, 0, p, o, q.length);
            o += q.length;
          }
          params.add(p);
        }
      }
      ShapeKey key = new ShapeKey(body.getW(), body.getH(), mask, sensors, hidden, breakable, robotMaterial);
      Integer idx = shapeIndex.get(key);
      if (idx == null) {
        idx = shapes.size();
        shapeIndex.put(key, idx);
        shapes.add(key);
      }
      robotShape[r] = idx;
    }

    // ---- native memory
    MemorySegment desc = a.allocate(VsrLayouts.VSR_BATCH_DESC);
    desc.fill((byte) 0);
    MemorySegment shapeArr = a.allocate(VsrLayouts.VSR_SHAPE, shapes.size());
    shapeArr.fill((byte) 0);
    for (int i = 0; i < shapes.size(); i++) {
      ShapeKey k = shapes.get(i);
      MemorySegment sh = shapeArr.asSlice(i * VsrLayouts.VSR_SHAPE.byteSize(), VsrLayouts.VSR_SHAPE.byteSize());
      sh.set(JAVA_INT, off(VsrLayouts.VSR_SHAPE, "w"), k.w());
      sh.set(JAVA_INT, off(VsrLayouts.VSR_SHAPE, "h"), k.h());
      MemorySegment mask = a.allocate(k.mask().size());
      for (int q = 0; q < k.mask().size(); q++) {
        mask.set(JAVA_BYTE, q, (byte) (k.mask().get(q) ? 1 : 0));
      }
      int nSens = k.sensors().stream().mapToInt(List::size).sum();
      MemorySegment offs = a.allocate(JAVA_INT, k.sensors().size() + 1);
      MemorySegment sens = a.allocate(VsrLayouts.VSR_SENSOR, Math.max(1, nSens));
      sens.fill((byte) 0);
      int o = 0;
      for (int v = 0; v < k.sensors().size(); v++) {
        offs.setAtIndex(JAVA_INT, v, o);
        for (FlatSensor f : k.sensors().get(v)) {
          MemorySegment s = sens.asSlice(o * VsrLayouts.VSR_SENSOR.byteSize(), VsrLayouts.VSR_SENSOR.byteSize());
          s.set(JAVA_INT, off(VsrLayouts.VSR_SENSOR, "base"), f.base());
          s.set(JAVA_INT, off(VsrLayouts.VSR_SENSOR, "axes"), f.axes());
          s.set(JAVA_INT, off(VsrLayouts.VSR_SENSOR, "rotated"), f.rotated());
          s.set(JAVA_INT, off(VsrLayouts.VSR_SENSOR, "aggregator"), f.aggregator());
          s.set(JAVA_INT, off(VsrLayouts.VSR_SENSOR, "soft_norm"), f.norm());
          s.set(JAVA_DOUBLE, off(VsrLayouts.VSR_SENSOR, "max_velocity_norm"), f.param());
          s.set(JAVA_DOUBLE, off(VsrLayouts.VSR_SENSOR, "interval"), f.interval());
          s.set(JAVA_DOUBLE, off(VsrLayouts.VSR_SENSOR, "noise_sigma"), f.noiseSigma());
          s.set(JAVA_LONG, off(VsrLayouts.VSR_SENSOR, "noise_seed"), f.noiseSeed());
          for (int q = 0; q < f.lidarDirs().length; q++) {
            s.set(JAVA_DOUBLE, off(VsrLayouts.VSR_SENSOR, "lidar_dirs") + 8L * q, f.lidarDirs()[q]);
          }
          o++;
        }
      }
      offs.setAtIndex(JAVA_INT, k.sensors().size(), o);
      MemorySegment hid = a.allocate(JAVA_INT, Math.max(1, k.hidden().size()));
      for (int q = 0; q < k.hidden().size(); q++) {
        hid.setAtIndex(JAVA_INT, q, k.hidden().get(q));
      }
      sh.set(java.lang.foreign.ValueLayout.ADDRESS, off(VsrLayouts.VSR_SHAPE, "mask"), mask);
      sh.set(java.lang.foreign.ValueLayout.ADDRESS, off(VsrLayouts.VSR_SHAPE, "sensor_offsets"), offs);
      sh.set(java.lang.foreign.ValueLayout.ADDRESS, off(VsrLayouts.VSR_SHAPE, "sensors"), sens);
      sh.set(JAVA_INT, off(VsrLayouts.VSR_SHAPE, "n_hidden"), k.hidden().size());
      sh.set(java.lang.foreign.ValueLayout.ADDRESS, off(VsrLayouts.VSR_SHAPE, "hidden"), hid);
      if (!k.material().equals(material)) {
        MemorySegment mm = a.allocate(VsrLayouts.VSR_MATERIAL);
        writeMaterial(mm, k.material());
        sh.set(java.lang.foreign.ValueLayout.ADDRESS, off(VsrLayouts.VSR_SHAPE, "material"), mm);
      }
      if (k.breakable().stream().anyMatch(q -> q.get(0) != 0d)) {
        MemorySegment br = a.allocate(VsrLayouts.VSR_BREAKABLE, k.breakable().size());
        br.fill((byte) 0);
        for (int v = 0; v < k.breakable().size(); v++) {
          List<Double> q = k.breakable().get(v);
          MemorySegment b = br.asSlice(v * VsrLayouts.VSR_BREAKABLE.byteSize(), VsrLayouts.VSR_BREAKABLE.byteSize());
          b.set(JAVA_INT, off(VsrLayouts.VSR_BREAKABLE, "enabled"), q.get(0).intValue());
          for (int c = 0; c < 3; c++) {
            b.set(JAVA_INT, off(VsrLayouts.VSR_BREAKABLE, "malfunctions") + 4L * c, q.get(1 + c).intValue());
            b.set(JAVA_DOUBLE, off(VsrLayouts.VSR_BREAKABLE, "thresholds") + 8L * c, q.get(4 + c));
          }
          b.set(JAVA_DOUBLE, off(VsrLayouts.VSR_BREAKABLE, "restore_time"), q.get(7));
          b.set(JAVA_LONG, off(VsrLayouts.VSR_BREAKABLE, "random_seed"), q.get(8).longValue());
        }
        sh.set(java.lang.foreign.ValueLayout.ADDRESS, off(VsrLayouts.VSR_SHAPE, "breakable"), br);
      }
    }
    long total = params.stream().mapToLong(p -> p.length).sum();
    MemorySegment pv = a.allocate(JAVA_DOUBLE, Math.max(1, total));
    MemorySegment po = a.allocate(JAVA_LONG, robots.size() + 1L);
    long o = 0;
    for (int r = 0; r < robots.size(); r++) {
      po.setAtIndex(JAVA_LONG, r, o);
      MemorySegment.copy(params.get(r), 0, pv, JAVA_DOUBLE, 8L * o, params.get(r).length);
      o += params.get(r).length;
    }
    po.setAtIndex(JAVA_LONG, robots.size(), o);
    MemorySegment rs = a.allocateFrom(JAVA_INT, robotShape);
    MemorySegment xs = a.allocateFrom(JAVA_DOUBLE, task.groundProfile()[0]);
    MemorySegment ys = a.allocateFrom(JAVA_DOUBLE, task.groundProfile()[1]);

    var D = VsrLayouts.VSR_BATCH_DESC;
    var AD = java.lang.foreign.ValueLayout.ADDRESS;
    desc.set(JAVA_INT, off(D, "abi_version"), VsrLayouts.VSR_ABI_VERSION);
    desc.set(JAVA_INT, off(D, "precision"), precision);
    desc.set(JAVA_LONG, off(D, "n_robots"), robots.size());
    desc.set(JAVA_INT, off(D, "n_shapes"), shapes.size());
    desc.set(JAVA_INT, off(D, "controller_kind"), kind);
    desc.set(AD, off(D, "shapes"), shapeArr);
    desc.set(AD, off(D, "robot_shape"), rs);
    desc.set(JAVA_INT, off(D, "mlp_activation"), activation);
    desc.set(JAVA_INT, off(D, "distributed_signals"), signals);
    desc.set(AD, off(D, "params"), pv);
    desc.set(AD, off(D, "param_offsets"), po);
    desc.set(JAVA_DOUBLE, off(D, "final_t"), task.finalT());
    desc.set(AD, off(D, "terrain_xs"), xs);
    desc.set(AD, off(D, "terrain_ys"), ys);
    desc.set(JAVA_INT, off(D, "terrain_n"), task.groundProfile()[0].length);
    desc.set(JAVA_DOUBLE, off(D, "initial_placement"), task.initialPlacement());
    desc.set(JAVA_LONG, off(D, "trajectory_robots"), 0L);
    desc.set(JAVA_DOUBLE, off(D, "t_start"), 0d);
    desc.set(AD, off(D, "robot_placement"), MemorySegment.NULL);
    writeMaterial(desc.asSlice(off(D, "material"), VsrLayouts.VSR_MATERIAL.byteSize()), material);
    writeSettings(desc.asSlice(off(D, "settings"), VsrLayouts.VSR_SETTINGS.byteSize()), task);
    return desc;
  }

  private static long off(java.lang.foreign.StructLayout l, String field) {
    return l.byteOffset(PathElement.groupElement(field));
  }

  /** MultiLayerPerceptron.ActivationFunction enum order = VSR_ACT_* (MultiLayerPerceptron.java:89-95) */
  private static int activationOf(MultiLayerPerceptron mlp) {
    return MultiLayerPerceptron.ActivationFunction.valueOf(OM.valueToTree(mlp).get("activationFunction").asText()).ordinal();
  }

  /** the 13 Voxel constructor fields as the reference persists them (Voxel.java:69-94) */
  private static void writeMaterial(MemorySegment m, JsonNode v) {
    var M = VsrLayouts.VSR_MATERIAL;
    m.set(JAVA_DOUBLE, off(M, "side_length"), v.get("sideLength").asDouble());
    m.set(JAVA_DOUBLE, off(M, "mass_side_length_ratio"), v.get("massSideLengthRatio").asDouble());
    m.set(JAVA_DOUBLE, off(M, "spring_f"), v.get("springF").asDouble());
    m.set(JAVA_DOUBLE, off(M, "spring_d"), v.get("springD").asDouble());
    m.set(JAVA_DOUBLE, off(M, "mass_linear_damping"), v.get("massLinearDamping").asDouble());
    m.set(JAVA_DOUBLE, off(M, "mass_angular_damping"), v.get("massAngularDamping").asDouble());
    m.set(JAVA_DOUBLE, off(M, "friction"), v.get("friction").asDouble());
    m.set(JAVA_DOUBLE, off(M, "restitution"), v.get("restitution").asDouble());
    m.set(JAVA_DOUBLE, off(M, "mass"), v.get("mass").asDouble());
    // DoubleRange is a record (min, max); Jackson writes +-Infinity as the strings "-Infinity" / "Infinity"
    m.set(JAVA_DOUBLE, off(M, "area_ratio_passive_min"), Double.parseDouble(v.get("areaRatioPassiveRange").get("min").asText()));
    m.set(JAVA_DOUBLE, off(M, "area_ratio_passive_max"), Double.parseDouble(v.get("areaRatioPassiveRange").get("max").asText()));
    m.set(JAVA_DOUBLE, off(M, "area_ratio_active_min"), v.get("areaRatioActiveRange").get("min").asDouble());
    m.set(JAVA_DOUBLE, off(M, "area_ratio_active_max"), v.get("areaRatioActiveRange").get("max").asDouble());
    int sc = 0;
    for (JsonNode s : v.get("springScaffoldings")) {   // Voxel.SpringScaffolding, Voxel.java:162-164
      sc |= switch (s.asText()) {
        case "SIDE_EXTERNAL" -> VsrLayouts.SCAFFOLD_SIDE_EXTERNAL;
        case "SIDE_INTERNAL" -> VsrLayouts.SCAFFOLD_SIDE_INTERNAL;
        case "SIDE_CROSS" -> VsrLayouts.SCAFFOLD_SIDE_CROSS;
        default -> VsrLayouts.SCAFFOLD_CENTRAL_CROSS;
      };
    }
    m.set(JAVA_INT, off(M, "spring_scaffoldings"), sc);
  }

  /** org.dyn4j.dynamics.Settings getters + the World defaults the reference never changes */
  private static void writeSettings(MemorySegment s, LocomotionSpec task) {
    var T = VsrLayouts.VSR_SETTINGS;
    var st = task.settings();
    s.set(JAVA_DOUBLE, off(T, "step_frequency"), st.getStepFrequency());
    s.set(JAVA_INT, off(T, "velocity_iterations"), st.getVelocityConstraintSolverIterations());
    s.set(JAVA_INT, off(T, "position_iterations"), st.getPositionConstraintSolverIterations());
    s.set(JAVA_DOUBLE, off(T, "linear_tolerance"), st.getLinearTolerance());
    s.set(JAVA_DOUBLE, off(T, "angular_tolerance"), st.getAngularTolerance());
    s.set(JAVA_DOUBLE, off(T, "max_linear_correction"), st.getMaximumLinearCorrection());
    s.set(JAVA_DOUBLE, off(T, "baumgarte"), st.getBaumgarte());
    s.set(JAVA_DOUBLE, off(T, "restitution_velocity"), st.getRestitutionVelocity());
    s.set(JAVA_DOUBLE, off(T, "max_translation"), st.getMaximumTranslation());
    s.set(JAVA_DOUBLE, off(T, "max_rotation"), st.getMaximumRotation());
    s.set(JAVA_DOUBLE, off(T, "gravity_x"), 0d);                // World.EARTH_GRAVITY, never changed by Locomotion
    s.set(JAVA_DOUBLE, off(T, "gravity_y"), -9.8);
    s.set(JAVA_DOUBLE, off(T, "ground_friction"), 0.2);         // BodyFixture.DEFAULT_FRICTION (Ground.java:72 passes none)
    s.set(JAVA_DOUBLE, off(T, "ground_restitution"), 0d);       // BodyFixture.DEFAULT_RESTITUTION
    s.set(JAVA_INT, off(T, "spring_mass_mode"), 0);             // VSR_SPRING_MASS_REDUCED
    s.set(JAVA_INT, off(T, "self_collision"), 1);               // Voxel.java:178-185,474
    // World.solveTOI: only ALL (the default of `new Settings()`) tests ordinary dynamic bodies against static ones;
    // BULLETS_ONLY is NONE here, the reference never creates a bullet
    s.set(JAVA_INT, off(T, "continuous_detection"),
        st.getContinuousDetectionMode() == org.dyn4j.dynamics.ContinuousDetectionMode.ALL ? 1 : 0);
    s.set(JAVA_INT, off(T, "reserved"), 0);
  }
}
