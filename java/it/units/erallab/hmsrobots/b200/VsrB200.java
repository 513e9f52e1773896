/*
 * Panama binding of libvsr_b200.so (include/vsr_b200.h) and the batched entry next to Locomotion.apply(Robot)
 * (tasks/locomotion/Locomotion.java:168-207):
 *
 *     List<SummaryOutcome> outcomes = VsrB200.run(locomotion, robots);
 *
 * SOURCE ONLY (no JDK in the build image: never compiled there).  JDK 22+; no CPU fallback: a missing library or
 * GPU is an IllegalStateException, never a silent dyn4j run.
 */
package it.units.erallab.hmsrobots.b200;

import it.units.erallab.hmsrobots.core.objects.Robot;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.util.List;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

public final class VsrB200 {
  private static final Linker L = Linker.nativeLinker();
  private static final SymbolLookup LIB = SymbolLookup.libraryLookup(
      System.getProperty("vsr.b200.library", "libvsr_b200.so"), Arena.global());

  private static MethodHandle h(String name, FunctionDescriptor d) {
    return L.downcallHandle(LIB.find(name).orElseThrow(() -> new IllegalStateException("missing symbol " + name)), d);
  }

  static final MethodHandle ABI_VERSION = h("vsr_abi_version", FunctionDescriptor.of(JAVA_INT));
  static final MethodHandle DEVICE_COUNT = h("vsr_device_count", FunctionDescriptor.of(JAVA_INT));
  static final MethodHandle CREATE = h("vsr_batch_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle UPLOAD = h("vsr_batch_upload", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
  static final MethodHandle RUN = h("vsr_batch_run", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  static final MethodHandle RESULTS = h("vsr_batch_results", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
  static final MethodHandle TRAJECTORY = h("vsr_batch_trajectory", FunctionDescriptor.of(JAVA_LONG, ADDRESS, JAVA_LONG, ADDRESS, JAVA_LONG));
  static final MethodHandle VOXEL_TRACE = h("vsr_batch_voxel_trace", FunctionDescriptor.of(JAVA_LONG, ADDRESS, JAVA_LONG, ADDRESS, JAVA_LONG));
  static final MethodHandle SENSOR_READINGS = h("vsr_batch_sensor_readings", FunctionDescriptor.of(JAVA_LONG, ADDRESS, JAVA_LONG, ADDRESS, JAVA_LONG));
  static final MethodHandle FINAL_BBOX = h("vsr_batch_final_bbox", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
  static final MethodHandle DESTROY = h("vsr_batch_destroy", FunctionDescriptor.ofVoid(ADDRESS));
  static final MethodHandle LAST_ERROR = h("vsr_last_error", FunctionDescriptor.of(ADDRESS));

  private VsrB200() {
  }

  /** error codes -> the reference's unchecked exceptions (IllegalArgumentException at validation, e.g. Voxel.java:132-140) */
  static void check(int rc) throws Throwable {
    if (rc == VsrLayouts.OK) {
      return;
    }
    MemorySegment msg = ((MemorySegment) LAST_ERROR.invoke()).reinterpret(4096);
    String text = msg.getString(0);
    switch (rc) {
      case VsrLayouts.ERR_INVALID_ARGUMENT -> throw new IllegalArgumentException(text);
      case VsrLayouts.ERR_UNSUPPORTED -> throw new UnsupportedOperationException(text);
      default -> throw new IllegalStateException("vsr error " + rc + ": " + text);
    }
  }

  /**
   * The batched form of Locomotion.apply: all robots in one launch per shape class, on CUDA device `device`.
   * The robots are NOT mutated (the reference's apply calls robot.reset(); the device builds its own state).
   */
  public static List<SummaryOutcome> run(LocomotionSpec task, List<Robot> robots, int device, int precision) {
    try (Arena a = Arena.ofConfined()) {
      if ((int) ABI_VERSION.invoke() != VsrLayouts.VSR_ABI_VERSION) {
        throw new IllegalStateException("libvsr_b200.so: ABI version mismatch");
      }
      MemorySegment desc = DescWriter.write(a, task, robots, precision);
      MemorySegment out = a.allocate(ADDRESS);
      check((int) CREATE.invoke(desc, out));
      MemorySegment batch = out.get(ADDRESS, 0);
      try {
        check((int) UPLOAD.invoke(batch, device, MemorySegment.NULL));
        check((int) RUN.invoke(batch, MemorySegment.NULL));
        MemorySegment res = a.allocate(VsrLayouts.VSR_RESULT.byteSize() * robots.size(), 8);
        check((int) RESULTS.invoke(batch, res, (long) robots.size()));
        return OutcomeReader.read(res, robots.size());
      } finally {
        DESTROY.invoke(batch);
      }
    } catch (RuntimeException e) {
      throw e;
    } catch (Throwable t) {
      throw new IllegalStateException(t);
    }
  }

  public static List<SummaryOutcome> run(LocomotionSpec task, List<Robot> robots) {
    return run(task, robots, 0, VsrLayouts.PRECISION_F32);
  }
}
