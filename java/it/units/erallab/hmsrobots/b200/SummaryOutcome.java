/*
 * What Outcome computes from its first and last Observation (tasks/locomotion/Outcome.java:42-58,126-142,152-200),
 * read from one vsr_result.  SOURCE ONLY.
 */
package it.units.erallab.hmsrobots.b200;

public record SummaryOutcome(
    double firstCenterX, double firstCenterY, double lastCenterX, double lastCenterY, double tFirst, double tLast,
    double areaRatioEnergyFirst, double areaRatioEnergyLast, double controlEnergyFirst, double controlEnergyLast,
    int nTicks, int status
) {
  /** Outcome.getDistance, :152-166 */
  public double getDistance() {
    return lastCenterX - firstCenterX;
  }

  /** Outcome.getTime, :194-196 */
  public double getTime() {
    return tLast - tFirst;
  }

  /** Outcome.getVelocity, :198-200 */
  public double getVelocity() {
    return getDistance() / getTime();
  }

  /** Outcome.getAreaRatioEnergy, :42-49: last minus first of the voxels' accumulated energies */
  public double getAreaRatioEnergy() {
    return areaRatioEnergyLast - areaRatioEnergyFirst;
  }

  /** Outcome.getAreaRatioPower, :51-53 */
  public double getAreaRatioPower() {
    return getAreaRatioEnergy() / getTime();
  }

  /** Outcome.getControlEnergy, :126-133 */
  public double getControlEnergy() {
    return controlEnergyLast - controlEnergyFirst;
  }

  /** Outcome.getControlPower, :135-137 */
  public double getControlPower() {
    return getControlEnergy() / getTime();
  }

  /** Outcome.getCorrectedEfficiency, :139-142 */
  public double getCorrectedEfficiency() {
    return getDistance() / (1d + getControlPower() * getTime());
  }

  /** 0 = a valid episode; see VSR_STATUS_* in include/vsr_b200.h (the device flags, it never truncates silently) */
  public boolean isValid() {
    return status == 0;
  }
}
