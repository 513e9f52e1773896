/*
 * The fields of a Locomotion task the device needs.  Locomotion keeps them private (Locomotion.java:47-59: finalT,
 * groundProfile, initialPlacement, settings) without getters, so the maintainer either adds the four getters or builds
 * this record where the Locomotion is built (same constructor arguments).  SOURCE ONLY.
 */
package it.units.erallab.hmsrobots.b200;

import org.dyn4j.dynamics.Settings;

public record LocomotionSpec(double finalT, double[][] groundProfile, double initialPlacement, Settings settings) {
  /** Locomotion(finalT, groundProfile, settings): initialPlacement = xs[1] + 1 (Locomotion.java:40,51) */
  public LocomotionSpec(double finalT, double[][] groundProfile, Settings settings) {
    this(finalT, groundProfile, groundProfile[0][1] + 1d, settings);
  }
}
