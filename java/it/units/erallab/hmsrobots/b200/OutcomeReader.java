/* vsr_result[n] -> SummaryOutcome list.  SOURCE ONLY. */
package it.units.erallab.hmsrobots.b200;

import java.lang.foreign.MemorySegment;
import java.util.ArrayList;
import java.util.List;

import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;

final class OutcomeReader {
  private OutcomeReader() {
  }

  static List<SummaryOutcome> read(MemorySegment res, int n) {
    final long stride = VsrLayouts.VSR_RESULT.byteSize();
    List<SummaryOutcome> out = new ArrayList<>(n);
    for (int i = 0; i < n; i++) {
      long o = i * stride;
      double[] v = new double[10];
      for (int k = 0; k < 10; k++) {
        v[k] = res.get(JAVA_DOUBLE, o + 8L * k);
      }
      int nTicks = res.get(JAVA_INT, o + 80);
      int status = res.get(JAVA_INT, o + 84);
      SummaryOutcome s = new SummaryOutcome(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], nTicks, status);
      if (!s.isValid()) {
        throw new IllegalStateException("robot " + i + ": vsr_result.status = " + status
            + " (1 non-finite state; 2/4/8 contact capacity exceeded, constraints were dropped)");
      }
      out.add(s);
    }
    return out;
  }
}
