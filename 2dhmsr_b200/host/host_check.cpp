// host_check.cpp -- exercises the C++ host mirror (hmsrobots.hpp).
//   host_check selfcheck     CPU only: prints a JSON document of host-layer facts that the tests
//                            compare with the reference's own vectors and with the Python mirror
//   host_check run           needs a GPU: BASELINE config 1 (Example.java:40-62) + 3 PhaseSin
//                            robots through Locomotion::apply, prints the Outcomes as JSON
#include <cstdio>
#include <iostream>
#include <sstream>

#include "hmsrobots.hpp"

using namespace hmsrobots;

static std::string jarr(const std::vector<double>& v) {
  std::ostringstream o;
  o.precision(17);
  o << "[";
  for (size_t i = 0; i < v.size(); i++) o << (i ? "," : "") << v[i];
  o << "]";
  return o.str();
}
static std::string mask(const Grid<bool>& g) {
  std::string s;
  for (auto& v : g.values()) s += (v && *v) ? '1' : '0';
  return s;
}

static Robot config1Robot() {  // Example.java:41-48
  Grid<bool> shape = RobotUtils::buildShape("biped-4x3");
  Grid<Voxel> body = RobotUtils::buildSensorizingFunction("uniform-ax+t-0")(shape);
  const int w = body.getW();
  auto fns = Grid<TimeFunctions::Fn>::create(w, body.getH(), [w](int x, int) {
    return std::optional<TimeFunctions::Fn>([x, w](double t) { return std::sin(-2 * M_PI * t + M_PI * ((double)x / (double)w)); });
  });
  return Robot(std::make_shared<TimeFunctions>(fns), body);
}
static Robot phaseRobot(double base) {
  Grid<Voxel> body = RobotUtils::buildSensorizingFunction("uniform-ax+t-0")(RobotUtils::buildShape("biped-4x3"));
  auto ph = Grid<double>::create(body.getW(), body.getH(), [base](int x, int y) { return std::optional<double>(base + 0.7 * x + 0.3 * y); });
  return Robot(std::make_shared<PhaseSin>(1.0, 1.0, ph), body);
}

// Starter.java:176-190: DistributedSensing, one MLP(TANH, [2]) per voxel, weights from java.util.Random(seed)
static Robot distributedRobot(long seed, bool directional) {
  Grid<Voxel> body = RobotUtils::buildSensorizingFunction("uniform-ax+t-0")(RobotUtils::buildShape("biped-4x3"));
  auto ds = std::make_shared<DistributedSensing>(body, 1, directional);
  JavaRandom rnd(seed);
  for (int y = 0; y < body.getH(); y++)
    for (int x = 0; x < body.getW(); x++) {
      if (!body.get(x, y)) continue;
      const int nin = ds->nOfInputs(x, y), nout = ds->nOfOutputs(x, y);
      std::vector<double> ws(MultiLayerPerceptron::countWeights({nin, 2, nout}));
      for (auto& w : ws) w = rnd.nextDouble() * 2.0 - 1.0;
      ds->setFunction(x, y, MultiLayerPerceptron(MultiLayerPerceptron::TANH, nin, {2}, nout, ws));
    }
  return Robot(ds, body);
}

template <class F>
static std::string throwsWhat(F f) {
  try {
    f();
  } catch (const Unsupported& e) {
    return std::string("Unsupported:") + e.what();
  } catch (const std::invalid_argument& e) {
    return std::string("invalid_argument:") + e.what();
  } catch (const std::exception& e) {
    return std::string("exception:") + e.what();
  }
  return "none";
}

int main(int argc, char** argv) {
  std::string mode = argc > 1 ? argv[1] : "selfcheck";
  std::cout.precision(17);
  if (mode == "selfcheck") {
    std::cout << "{\n";
    JavaRandom r0(0), r0b(0), r42(42), r0c(0);
    std::cout << " \"seed0_nextDouble\": " << r0.nextDouble() << ",\n \"seed0_nextInt\": " << r0b.nextInt()
              << ",\n \"seed42_nextInt\": " << r42.nextInt() << ",\n \"seed0_nextGaussian\": " << r0c.nextGaussian() << ",\n";
    auto hilly = Locomotion::createTerrain("hilly-1-10-0");
    std::cout << " \"hilly_xs\": " << jarr(hilly.first) << ",\n \"hilly_ys\": " << jarr(hilly.second) << ",\n";
    auto st = Locomotion::createTerrain("steppy-2-8-3");
    std::cout << " \"steppy_xs\": " << jarr(st.first) << ",\n \"steppy_ys\": " << jarr(st.second) << ",\n";
    auto fw = Locomotion::createTerrain("flatWithStart-2");
    std::cout << " \"flatWithStart_ys\": " << jarr(fw.second) << ",\n";
    std::cout << " \"uphill_ys\": " << jarr(Locomotion::createTerrain("uphill-10").second) << ",\n";
    std::cout << " \"shapes\": {";
    const char* names[] = {"biped-4x3", "worm-8x2", "box-10x10", "tripod-5x3", "comb-5x4", "ball-5", "t-5x4", "biped-7x4"};
    for (size_t i = 0; i < 8; i++) std::cout << (i ? ", " : "") << "\"" << names[i] << "\": \"" << mask(RobotUtils::buildShape(names[i])) << "\"";
    std::cout << "},\n";
    // MultiLayerPerceptronTest.java:50-95
    MultiLayerPerceptron mlp(MultiLayerPerceptron::RELU, 1, {2}, 1, {1, 0, 1, 2, 1, -1, 1});
    std::cout << " \"mlp_apply\": " << jarr(mlp.apply({2})) << ",\n";
    auto u = MultiLayerPerceptron::unflat({1, 2, 3, 4, 5, 6, 7, 8, 9}, {2, 2, 1});
    std::cout << " \"mlp_unflat_layer0_neuron1\": " << jarr(u[0][1]) << ",\n \"mlp_flat\": " << jarr(MultiLayerPerceptron::flat(u, {2, 2, 1})) << ",\n";
    std::cout << " \"count_weights\": " << MultiLayerPerceptron::countWeights({32, 21, 21, 16}) << ",\n";
    auto ticks = tickTimes(20, 1.0 / 60.0);
    std::cout << " \"n_ticks\": " << ticks.size() << ",\n \"last_tick\": " << ticks.back() << ",\n";
    // descriptors
    Locomotion loco(20, Locomotion::createTerrain("flat"), Settings());
    auto bd1 = loco.compile({config1Robot()});
    std::cout << " \"cfg1_kind\": " << bd1->desc.controller_kind << ",\n \"cfg1_n_params\": " << bd1->params.size()
              << ",\n \"cfg1_param_5_1\": " << bd1->params[5 * 10 + 1] << ",\n \"cfg1_initial_placement\": " << bd1->desc.initial_placement << ",\n";
    std::vector<double> enc;
    for (auto& s : bd1->sensors[0])
      for (double q : {double(s.base), double(s.axes), double(s.rotated), double(s.aggregator), double(s.soft_norm), s.max_velocity_norm, s.interval}) enc.push_back(q);
    std::cout << " \"cfg1_sensors\": " << jarr(enc) << ",\n";
    auto bd2 = loco.compile({phaseRobot(0.1), phaseRobot(0.2), config1Robot().getController().kind() == VSR_CTRL_TABULATED ? phaseRobot(0.3) : phaseRobot(0.4)});
    std::cout << " \"phase_params\": " << jarr(bd2->params) << ",\n \"phase_n_shapes\": " << bd2->desc.n_shapes << ",\n";
    Grid<Voxel> worm = RobotUtils::buildSensorizingFunction("uniform-t+vxy+a-0")(RobotUtils::buildShape("worm-8x2"));
    std::cout << " \"worm_inputs\": " << CentralizedSensing::nOfInputs(worm) << ",\n";
    // DistributedSensing + the sensors beyond the first round
    auto bd3 = loco.compile({distributedRobot(7, true)});
    auto bd4 = loco.compile({distributedRobot(7, false)});
    std::cout << " \"dist_kind\": " << bd3->desc.controller_kind << ",\n \"dist_signals\": " << bd3->desc.distributed_signals
              << ",\n \"dist_n_params\": " << bd3->params.size() << ",\n \"dist_param_0\": " << bd3->params[0]
              << ",\n \"dist_nd_kind\": " << bd4->desc.controller_kind << ",\n \"dist_nd_n_params\": " << bd4->params.size() << ",\n";
    Grid<Voxel> spined = RobotUtils::buildSensorizingFunction("spinedTouch-t-t-0")(RobotUtils::buildShape("biped-4x3"));
    std::vector<double> nsens;
    for (auto& v : spined.values())
      if (v) nsens.push_back(double(v->getSensors().size()));
    std::cout << " \"spined_sensor_counts\": " << jarr(nsens) << ",\n";
    Grid<Voxel> rich = RobotUtils::buildSensorizingFunction("uniform-r+px+py+cpg+m-0")(RobotUtils::buildShape("worm-4x2"));
    std::vector<double> enc2;
    for (auto& sp : rich.get(3, 1)->getSensors()) {
      vsr_sensor q = sp->encode();
      for (double z : {double(q.base), double(q.soft_norm), q.max_velocity_norm}) enc2.push_back(z);
    }
    std::cout << " \"rich_sensors_3_1\": " << jarr(enc2) << ",\n";
    // error behaviour
    std::cout << " \"err_terrain\": \"" << throwsWhat([] { Locomotion::createTerrain("moon"); }) << "\",\n";
    std::cout << " \"err_shape\": \"" << throwsWhat([] { RobotUtils::buildShape("snake-3"); }) << "\",\n";
    std::cout << " \"err_sensors\": \"" << throwsWhat([] { RobotUtils::buildSensorizingFunction("uniform-zz-0"); }) << "\",\n";
    {
      Grid<Voxel> sighted = RobotUtils::buildSensorizingFunction("uniform-l5-0")(RobotUtils::buildShape("worm-3x2"));
      vsr_sensor q = sighted.get(2, 0)->getSensors()[0]->encode();   // bottom right corner: side E
      std::vector<double> e = {double(q.base), double(q.axes), double(q.soft_norm), q.max_velocity_norm};
      for (int i = 0; i < q.axes; i++) e.push_back(q.lidar_dirs[i]);
      std::cout << " \"lidar_2_0\": " << jarr(e) << ",\n";
    }
    {
      // uniformAll (RobotUtils.java:177-189): the 15 predefined sensors in sorted order = 21 readings per voxel;
      // spinedTouchSighted (:150-165): the last column also carries the 5-ray lidar
      Grid<Voxel> all = RobotUtils::buildSensorizingFunction("uniformAll-0")(RobotUtils::buildShape("box-2x2"));
      std::vector<double> bases;
      int readings = 0;
      for (auto& sp : all.get(1, 1)->getSensors()) {
        bases.push_back(double(sp->encode().base));
        readings += int(sp->getDomains().size());
      }
      std::cout << " \"all_bases_1_1\": " << jarr(bases) << ", \"all_readings\": " << readings << ",\n";
      Grid<Voxel> ss = RobotUtils::buildSensorizingFunction("spinedTouchSighted-t-f-0")(RobotUtils::buildShape("biped-4x3"));
      std::vector<double> cnt;
      for (auto& e : ss.values()) if (e) cnt.push_back(double(e->getSensors().size()));
      std::cout << " \"sighted_sensor_counts\": " << jarr(cnt) << ",\n";
    }
    std::cout << " \"err_weights\": \"" << throwsWhat([] { MultiLayerPerceptron(MultiLayerPerceptron::TANH, 3, {2}, 1, std::vector<double>(5)); }) << "\",\n";
    std::cout << " \"err_dims\": \"" << throwsWhat([&] { CentralizedSensing(worm, MultiLayerPerceptron(MultiLayerPerceptron::TANH, 63, {}, 16, std::vector<double>(16 * 64))); }) << "\",\n";
    std::cout << " \"err_ground\": \"" << throwsWhat([] { Locomotion(1, {{0, 5, 3}, {0, 0, 0}}, Settings()); }) << "\",\n";
    std::cout << " \"err_range\": \"" << throwsWhat([] { DoubleRange(1, 0); }) << "\"\n}\n";
    return 0;
  }
  if (mode == "run") {
    Locomotion loco(20, Locomotion::createTerrain("flat"), Settings());
    Outcome o1 = loco.apply(config1Robot());
    auto outs = loco.apply({phaseRobot(0.1), phaseRobot(0.2), phaseRobot(0.3)});
    std::cout << "{ \"cfg1\": {\"distance\": " << o1.getDistance() << ", \"time\": " << o1.getTime() << ", \"velocity\": " << o1.getVelocity()
              << ", \"control_energy\": " << o1.getControlEnergy() << "},\n  \"phase_distance\": [";
    for (size_t i = 0; i < outs.size(); i++) std::cout << (i ? ", " : "") << outs[i].getDistance();
    Locomotion loco5(5, Locomotion::createTerrain("flat"), Settings());
    auto d1 = loco5.apply({distributedRobot(7, true), distributedRobot(8, true)});
    auto d2 = loco5.apply(std::vector<Robot>{distributedRobot(7, false)});
    // the per-tick Observation map of a 2 s episode
    Locomotion loco2(2, Locomotion::createTerrain("flat"), Settings());
    auto ob = loco2.applyWithObservations({phaseRobot(0.1)})[0];
    int touches = 0;
    for (auto& kv : ob.getObservations())
      for (auto& p : Outcome::polies(kv.second)) touches += p.isTouchingGround() ? 1 : 0;
    auto& lastObs = ob.getObservations().rbegin()->second;
    auto fp = BehaviorUtils::computeFootprint(Outcome::polies(lastObs), 4);
    std::cout << "],\n  \"obs\": {\"n\": " << ob.getObservations().size() << ", \"distance\": " << ob.getDistanceFromObservations()
              << ", \"record_distance\": " << ob.getDistance() << ", \"touches\": " << touches
              << ", \"last_control_energy\": " << Outcome::polies(lastObs)[0].getControlEnergy()
              << ", \"terrain_height\": " << lastObs.terrainHeight << ", \"footprint\": \"" << (fp[0] ? '_' : '.') << (fp[1] ? '_' : '.')
              << (fp[2] ? '_' : '.') << (fp[3] ? '_' : '.') << "\", \"sub\": " << ob.subOutcome(0.5, 1.0).getObservations().size() << "}";
    std::cout << ",\n  \"distributed_distance\": [" << d1[0].getDistance() << ", " << d1[1].getDistance() << ", " << d2[0].getDistance()
              << "] }\n";
    return 0;
  }
  std::fprintf(stderr, "usage: host_check selfcheck|run\n");
  return 2;
}
