/* The drop-in boundary from plain C: BASELINE config 2 in miniature (a few biped-4x3 robots with PhaseSin
 * controllers on flat ground, 20 s) through the six calls a Java binding makes.
 *
 *   gcc -O2 -I include examples/c_abi_example.c -L 2dhmsr_b200 -lvsr_b200 -Wl,-rpath,$PWD/2dhmsr_b200 -lm -o c_abi_example
 *   ./c_abi_example [n_robots]      (needs a GPU: there is no CPU fallback)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vsr_b200.h"

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 4;
  /* biped-4x3 (RobotUtils.java:206-215): 4 x 3 grid without the two bottom centre cells */
  enum { W = 4, Hh = 3 };
  uint8_t mask[W * Hh];
  int nv = 0;
  for (int y = 0; y < Hh; y++)
    for (int x = 0; x < W; x++) {
      mask[y * W + x] = !(y < Hh / 2 && x >= W / 4 && x < W * 3 / 4);
      nv += mask[y * W + x];
    }
  /* uniform-ax+t-0 (RobotUtils.java:38-53): per voxel SoftNormalization(Trend(Velocity x, 0.5 s)), Average(Touch, 0.25 s) */
  vsr_sensor* sensors = (vsr_sensor*)calloc((size_t)nv * 2, sizeof(vsr_sensor));
  int32_t* offs = (int32_t*)calloc((size_t)nv + 1, sizeof(int32_t));
  for (int v = 0; v < nv; v++) {
    vsr_sensor* ax = &sensors[2 * v];
    ax->base = VSR_SENSOR_VELOCITY; ax->axes = VSR_AXIS_X; ax->rotated = 1; ax->max_velocity_norm = 4.0;
    ax->aggregator = VSR_AGG_TREND; ax->interval = 0.5; ax->soft_norm = VSR_NORM_SOFT;
    vsr_sensor* t = &sensors[2 * v + 1];
    t->base = VSR_SENSOR_TOUCH; t->aggregator = VSR_AGG_AVERAGE; t->interval = 0.25;
    offs[v + 1] = 2 * (v + 1);
  }
  vsr_shape shape;
  memset(&shape, 0, sizeof shape);
  shape.w = W; shape.h = Hh; shape.mask = mask; shape.sensor_offsets = offs; shape.sensors = sensors;

  /* PhaseSin (PhaseSin.java:61): params = [frequency, amplitude, phase per voxel] */
  const size_t np = 2 + (size_t)nv;
  double* params = (double*)malloc(sizeof(double) * np * (size_t)n);
  size_t* poffs = (size_t*)malloc(sizeof(size_t) * ((size_t)n + 1));
  for (int r = 0; r < n; r++) {
    poffs[r] = np * (size_t)r;
    params[poffs[r] + 0] = 1.0;
    params[poffs[r] + 1] = 1.0;
    int v = 0;
    for (int y = 0; y < Hh; y++)
      for (int x = 0; x < W; x++)
        if (mask[y * W + x]) params[poffs[r] + 2 + (size_t)(v++)] = 0.1 * (r + 1) + 0.7 * x + 0.3 * y;
  }
  poffs[n] = np * (size_t)n;

  const double xs[4] = {0, 10, 1990, 2000}, ys[4] = {100, 5, 5, 100}; /* Locomotion.createTerrain("flat") */
  vsr_batch_desc d;
  memset(&d, 0, sizeof d);
  d.abi_version = VSR_ABI_VERSION;
  d.precision = VSR_PRECISION_F32;
  d.n_robots = (size_t)n;
  d.n_shapes = 1;
  d.controller_kind = VSR_CTRL_PHASE_SIN;
  d.shapes = &shape;
  vsr_default_material(&d.material);
  d.params = params;
  d.param_offsets = poffs;
  d.final_t = 20.0;
  d.terrain_xs = xs; d.terrain_ys = ys; d.terrain_n = 4;
  d.initial_placement = xs[1] + 1.0;
  vsr_default_settings(&d.settings);

  vsr_batch* b = NULL;
  vsr_result* res = (vsr_result*)calloc((size_t)n, sizeof(vsr_result));
  int rc = vsr_batch_create(&d, &b);
  if (!rc) rc = vsr_batch_upload(b, 0, NULL);
  if (!rc) rc = vsr_batch_run(b, NULL);
  if (!rc) rc = vsr_batch_results(b, res, (size_t)n);
  if (rc) {
    fprintf(stderr, "vsr error %d: %s\n", rc, vsr_last_error());
    return 1;
  }
  printf("{\"n_ticks\": %d, \"distance\": [", res[0].n_ticks);
  for (int r = 0; r < n; r++) printf("%s%.17g", r ? ", " : "", res[r].last_center_x - res[r].first_center_x);
  printf("], \"velocity\": [");
  for (int r = 0; r < n; r++)
    printf("%s%.17g", r ? ", " : "", (res[r].last_center_x - res[r].first_center_x) / (res[r].t_last - res[r].t_first));
  printf("]}\n");
  vsr_batch_destroy(b);
  free(res); free(params); free(poffs); free(sensors); free(offs);
  return 0;
}
