// hmsrobots.hpp -- C++17 host-side mirror of the reference's Java API for the Locomotion hot
// path, above the C ABI of include/vsr_b200.h.  Header only.
//
// Same class names, argument meaning and error behaviour as it.units.erallab.hmsrobots
// (file:line cited per item); IllegalArgumentException -> std::invalid_argument,
// "valid in the reference but not on the accelerated path" -> hmsrobots::Unsupported,
// CUDA / call-order failures -> std::runtime_error.  Nothing here steps physics:
// Locomotion::apply flattens the robots into a vsr_batch_desc and calls libvsr_b200.so.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <optional>
#include <regex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/vsr_b200.h"

namespace hmsrobots {

struct Unsupported : std::logic_error {
  using std::logic_error::logic_error;
};

// ---------------------------------------------------------------- java.util.Random
// Bit-exact LCG + polar nextGaussian; needed by Locomotion.createTerrain (Locomotion.java:76-117).
class JavaRandom {
 public:
  explicit JavaRandom(int64_t seed) : seed_((uint64_t(seed) ^ kMult) & kMask) {}
  int32_t nextInt() { return next(32); }
  int32_t nextInt(int32_t bound) {
    if (bound <= 0) throw std::invalid_argument("bound must be positive");
    if ((bound & -bound) == bound) return int32_t((int64_t(bound) * int64_t(next(31))) >> 31);
    for (;;) {
      int32_t bits = next(31), val = bits % bound;
      if (int64_t(bits) - val + (bound - 1) < (int64_t(1) << 31)) return val;
    }
  }
  double nextDouble() { return double((int64_t(next(26)) << 27) + next(27)) * (1.0 / double(int64_t(1) << 53)); }
  double nextGaussian() {
    if (have_) {
      have_ = false;
      return nextNext_;
    }
    double v1, v2, s;
    do {
      v1 = 2 * nextDouble() - 1;
      v2 = 2 * nextDouble() - 1;
      s = v1 * v1 + v2 * v2;
    } while (s >= 1 || s == 0);
    double mul = std::sqrt(-2 * std::log(s) / s);
    nextNext_ = v2 * mul;
    have_ = true;
    return v1 * mul;
  }

 private:
  static constexpr uint64_t kMult = 0x5DEECE66DULL, kMask = (1ULL << 48) - 1;
  int32_t next(int bits) {
    seed_ = (seed_ * kMult + 0xBULL) & kMask;
    return int32_t(int64_t(seed_ >> (48 - bits)));
  }
  uint64_t seed_;
  bool have_ = false;
  double nextNext_ = 0;
};

// ---------------------------------------------------------------- util.Grid
// Storage ts[y*w+x] (Grid.java:180-188); create(w,h,filler) calls the filler x-outer / y-inner
// (:103-111); iteration and values() are row-major (:69-75,268-270).  Null cells = empty optional.
template <class T>
class Grid {
 public:
  Grid(int w, int h) : w_(w), h_(h), ts_(size_t(w) * h) {}
  static Grid create(int w, int h, const std::function<std::optional<T>(int, int)>& filler) {
    Grid g(w, h);
    for (int x = 0; x < w; x++)
      for (int y = 0; y < h; y++) g.ts_[size_t(y) * w + x] = filler(x, y);
    return g;
  }
  static Grid create(int w, int h, const T& v) {
    return create(w, h, [&](int, int) { return std::optional<T>(v); });
  }
  int getW() const { return w_; }
  int getH() const { return h_; }
  const std::optional<T>& get(int x, int y) const {
    static const std::optional<T> none;
    if (x < 0 || x >= w_ || y < 0 || y >= h_) return none;
    return ts_[size_t(y) * w_ + x];
  }
  void set(int x, int y, std::optional<T> v) { ts_[size_t(y) * w_ + x] = std::move(v); }
  const std::vector<std::optional<T>>& values() const { return ts_; }
  size_t count() const {
    size_t n = 0;
    for (auto& v : ts_) n += v.has_value();
    return n;
  }

 private:
  int w_, h_;
  std::vector<std::optional<T>> ts_;
};

struct DoubleRange {  // util/DoubleRange.java:29-58
  double min, max;
  DoubleRange(double mn, double mx) : min(mn), max(mx) {
    if (mx < mn) throw std::invalid_argument("Max has to be lower or equal than min");
  }
  double extent() const { return max - min; }
  double clip(double v) const { return std::min(std::max(v, min), max); }
  double normalize(double v) const { return (clip(v) - min) / (max - min); }
};

// ---------------------------------------------------------------- sensors (descriptions)
struct Sensor {  // core/sensors/Sensor.java:26-34
  virtual ~Sensor() = default;
  virtual std::vector<DoubleRange> getDomains() const = 0;
  virtual vsr_sensor encode() const = 0;
  virtual bool isAggregator() const { return false; }
  virtual bool isSoftNorm() const { return false; }
};
using SensorPtr = std::shared_ptr<const Sensor>;

struct Touch : Sensor {  // sensors/Touch.java
  std::vector<DoubleRange> getDomains() const override { return {DoubleRange(0, 1)}; }
  vsr_sensor encode() const override {
    vsr_sensor s{};
    s.base = VSR_SENSOR_TOUCH;
    return s;
  }
};
struct AreaRatio : Sensor {  // sensors/AreaRatio.java:21-37
  std::vector<DoubleRange> getDomains() const override { return {DoubleRange(0.5, 1.5)}; }
  vsr_sensor encode() const override {
    vsr_sensor s{};
    s.base = VSR_SENSOR_AREA_RATIO;
    return s;
  }
};
struct Velocity : Sensor {  // sensors/Velocity.java:29-80
  enum Axis { X = 1, Y = 2 };
  bool rotated;
  double maxVelocityNorm;
  int axes;
  Velocity(bool rotated_, double maxNorm, int axes_) : rotated(rotated_), maxVelocityNorm(maxNorm), axes(axes_) {}
  std::vector<DoubleRange> getDomains() const override {
    std::vector<DoubleRange> d;
    for (int a : {X, Y})
      if (axes & a) d.emplace_back(-maxVelocityNorm, maxVelocityNorm);
    return d;
  }
  vsr_sensor encode() const override {
    vsr_sensor s{};
    s.base = VSR_SENSOR_VELOCITY;
    s.axes = axes;
    s.rotated = rotated;
    s.max_velocity_norm = maxVelocityNorm;
    return s;
  }
};
struct AggregatorSensor : Sensor {  // sensors/AggregatorSensor.java:31-66
  SensorPtr sensor;
  double interval;
  int kind;
  AggregatorSensor(SensorPtr s, double i, int k) : sensor(std::move(s)), interval(i), kind(k) {
    if (sensor->isAggregator() || sensor->isSoftNorm()) throw Unsupported("nested aggregators are outside the hot path");
  }
  bool isAggregator() const override { return true; }
  vsr_sensor encode() const override {
    vsr_sensor s = sensor->encode();
    s.aggregator = kind;
    s.interval = interval;
    return s;
  }
};
struct Average : AggregatorSensor {  // sensors/Average.java
  Average(SensorPtr s, double i) : AggregatorSensor(std::move(s), i, VSR_AGG_AVERAGE) {}
  std::vector<DoubleRange> getDomains() const override { return sensor->getDomains(); }
};
struct Trend : AggregatorSensor {  // sensors/Trend.java:28-35
  Trend(SensorPtr s, double i) : AggregatorSensor(std::move(s), i, VSR_AGG_TREND) {}
  std::vector<DoubleRange> getDomains() const override {
    std::vector<DoubleRange> d;
    for (auto& r : sensor->getDomains()) d.emplace_back(-r.extent() / interval, r.extent() / interval);
    return d;
  }
};
struct DynamicNormalization : AggregatorSensor {  // sensors/DynamicNormalization.java:26-61 (0/0 = NaN at the first tick, as there)
  DynamicNormalization(SensorPtr s, double i) : AggregatorSensor(std::move(s), i, VSR_AGG_DYNAMIC_NORM) {}
  std::vector<DoubleRange> getDomains() const override {
    return std::vector<DoubleRange>(sensor->getDomains().size(), DoubleRange(0.0, 1.0));
  }
};
struct SoftNormalization : Sensor {  // sensors/SoftNormalization.java
  SensorPtr sensor;
  explicit SoftNormalization(SensorPtr s) : sensor(std::move(s)) {
    if (sensor->isSoftNorm()) throw Unsupported("nested SoftNormalization is outside the hot path");
  }
  bool isSoftNorm() const override { return true; }
  std::vector<DoubleRange> getDomains() const override {
    return std::vector<DoubleRange>(sensor->getDomains().size(), DoubleRange(0, 1));
  }
  vsr_sensor encode() const override {
    vsr_sensor s = sensor->encode();
    s.soft_norm = 1;
    return s;
  }
};

struct Normalization : Sensor {  // sensors/Normalization.java
  SensorPtr sensor;
  explicit Normalization(SensorPtr s) : sensor(std::move(s)) {
    if (sensor->isSoftNorm()) throw Unsupported("nested normalizations are outside the hot path");
  }
  bool isSoftNorm() const override { return true; }
  std::vector<DoubleRange> getDomains() const override {
    return std::vector<DoubleRange>(sensor->getDomains().size(), DoubleRange(0, 1));
  }
  vsr_sensor encode() const override {
    vsr_sensor s = sensor->encode();
    s.soft_norm = VSR_NORM_LINEAR;
    return s;
  }
};
struct Noisy : Sensor {  // sensors/Noisy.java: + Random(seed).nextGaussian() * sigma * extent of the reading's domain
  SensorPtr sensor;
  double sigma;
  long seed;
  Noisy(SensorPtr s, double sigma_, long seed_) : sensor(std::move(s)), sigma(sigma_), seed(seed_) {}
  bool isAggregator() const override { return sensor->isAggregator(); }
  bool isSoftNorm() const override { return sensor->isSoftNorm(); }
  std::vector<DoubleRange> getDomains() const override { return sensor->getDomains(); }
  vsr_sensor encode() const override {
    vsr_sensor s = sensor->encode();
    s.noise_sigma = sigma;
    s.noise_seed = seed;
    return s;
  }
};
struct Lidar : Sensor {  // sensors/Lidar.java
  enum Side { N, E, S, W };
  double rayLength;
  std::vector<double> rayDirections;
  Lidar(double len, std::vector<double> dirs) : rayLength(len), rayDirections(std::move(dirs)) {
    if (rayDirections.empty() || rayDirections.size() > VSR_LIDAR_MAX_RAYS)
      throw Unsupported("Lidar: 1..8 rays per sensor on the accelerated path");
  }
  static std::pair<double, double> sideAngles(Side s) {  // :66-86
    switch (s) {
      case N: return {M_PI / 4.0, M_PI * 3.0 / 4.0};
      case E: return {M_PI * -1.0 / 4.0, M_PI / 4.0};
      case S: return {M_PI * 5.0 / 4.0, M_PI * 7.0 / 4.0};
      default: return {M_PI * 3.0 / 4.0, M_PI * 5.0 / 4.0};
    }
  }
  static std::vector<double> sampleRangeWithRays(int n, double start, double end) {  // :88-94
    if (n == 1) return {(end - start) / 2};
    std::vector<double> out;
    double d = start;
    for (int i = 0; i < n; i++) {  // DoubleStream.iterate: running sum
      out.push_back(d);
      d = d + (end - start) / (n - 1);
    }
    return out;
  }
  static std::shared_ptr<Lidar> ofSide(double len, Side s, int n) {  // :57-65
    auto a = sideAngles(s);
    std::vector<double> dirs;
    for (double d : sampleRangeWithRays(n, a.first, a.second))
      if (std::find(dirs.begin(), dirs.end(), d) == dirs.end()) dirs.push_back(d);
    return std::make_shared<Lidar>(len, dirs);
  }
  std::vector<DoubleRange> getDomains() const override { return std::vector<DoubleRange>(rayDirections.size(), DoubleRange(0, rayLength)); }
  vsr_sensor encode() const override {
    vsr_sensor s{};
    s.base = VSR_SENSOR_LIDAR;
    s.axes = int32_t(rayDirections.size());
    s.max_velocity_norm = rayLength;
    for (size_t i = 0; i < rayDirections.size(); i++) s.lidar_dirs[i] = rayDirections[i];
    return s;
  }
};
struct Angle : Sensor {  // sensors/Angle.java
  std::vector<DoubleRange> getDomains() const override { return {DoubleRange(-M_PI, M_PI)}; }
  vsr_sensor encode() const override {
    vsr_sensor s{};
    s.base = VSR_SENSOR_ANGLE;
    return s;
  }
};
struct Constant : Sensor {  // sensors/Constant.java (one value per sensor on the accelerated path)
  double value;
  explicit Constant(double v) : value(v) {}
  std::vector<DoubleRange> getDomains() const override { return {DoubleRange(std::min(0.0, value), std::max(1.0, value))}; }
  vsr_sensor encode() const override {
    vsr_sensor s{};
    s.base = VSR_SENSOR_CONSTANT;
    s.max_velocity_norm = value;
    return s;
  }
};
struct AppliedForce : Sensor {  // sensors/AppliedForce.java
  std::vector<DoubleRange> getDomains() const override { return {DoubleRange(-1, 1)}; }
  vsr_sensor encode() const override {
    vsr_sensor s{};
    s.base = VSR_SENSOR_APPLIED_FORCE;
    return s;
  }
};
struct Malfunction : Sensor {  // sensors/Malfunction.java (a non-breakable voxel is never broken)
  std::vector<DoubleRange> getDomains() const override { return {DoubleRange(0, 1)}; }
  vsr_sensor encode() const override {
    vsr_sensor s{};
    s.base = VSR_SENSOR_MALFUNCTION;
    return s;
  }
};
struct TimeFunctionSin : Sensor {  // sensors/TimeFunction.java with t -> sin(2 pi f t) on [-1, 1] ("cpg": f = -1)
  double frequency;
  explicit TimeFunctionSin(double f) : frequency(f) {}
  std::vector<DoubleRange> getDomains() const override { return {DoubleRange(-1, 1)}; }
  vsr_sensor encode() const override {
    vsr_sensor s{};
    s.base = VSR_SENSOR_TIME_SIN;
    s.max_velocity_norm = frequency;
    return s;
  }
};

// ---------------------------------------------------------------- Voxel
class Voxel {  // core/objects/Voxel.java:103-160
 public:
  static constexpr double SIDE_LENGTH = 3.0;
  explicit Voxel(std::vector<SensorPtr> sensors) : sensors_(std::move(sensors)) { vsr_default_material_(&material_); }
  Voxel(std::vector<SensorPtr> sensors, const vsr_material& m) : sensors_(std::move(sensors)), material_(m) {
    double lim = (2 * m.mass_side_length_ratio) * (2 * m.mass_side_length_ratio);
    if (m.area_ratio_passive_min > -std::numeric_limits<double>::infinity() && m.area_ratio_passive_min < lim)
      throw std::invalid_argument("Min of areaRatioPassiveRange cannot be lower than the value imposed by massSideLengthRatio");  // :132-140
  }
  const std::vector<SensorPtr>& getSensors() const { return sensors_; }
  const vsr_material& material() const { return material_; }

 private:
  static void vsr_default_material_(vsr_material* m) {  // Voxel.java:57-68
    *m = vsr_material{};
    m->side_length = 3.0; m->mass_side_length_ratio = 0.35; m->spring_f = 8.0; m->spring_d = 0.3;
    m->mass_linear_damping = 0.1; m->mass_angular_damping = 0.1; m->friction = 10.0; m->restitution = 0.1;
    m->mass = 1.0;
    m->area_ratio_passive_min = -std::numeric_limits<double>::infinity();
    m->area_ratio_passive_max = std::numeric_limits<double>::infinity();
    m->area_ratio_active_min = 0.8; m->area_ratio_active_max = 1.2;
    m->spring_scaffoldings = VSR_SCAFFOLD_ALL;
  }
  std::vector<SensorPtr> sensors_;
  vsr_material material_;
};

// ---------------------------------------------------------------- controllers
// Locomotion.java:195-203 + AbstractTask.java:48: t = 0; while (t < finalT) t = t + dT
inline std::vector<double> tickTimes(double finalT, double dT) {
  std::vector<double> t;
  for (double tt = 0.0; tt < finalT;) {
    tt = tt + dT;
    t.push_back(tt);
  }
  return t;
}

struct Controller {  // core/controllers/Controller.java:29 (description only)
  virtual ~Controller() = default;
  virtual int kind() const = 0;
  virtual std::vector<double> params(const Grid<Voxel>& voxels, const std::vector<double>& ticks) const = 0;
};
using ControllerPtr = std::shared_ptr<const Controller>;

class TimeFunctions : public Controller {  // controllers/TimeFunctions.java:49-64
 public:
  using Fn = std::function<double(double)>;
  explicit TimeFunctions(Grid<Fn> functions) : functions_(std::move(functions)) {}
  int kind() const override { return VSR_CTRL_TABULATED; }
  std::vector<double> params(const Grid<Voxel>& voxels, const std::vector<double>& ticks) const override {
    std::vector<std::pair<int, int>> cells;
    for (int y = 0; y < voxels.getH(); y++)
      for (int x = 0; x < voxels.getW(); x++)
        if (voxels.get(x, y)) cells.emplace_back(x, y);
    std::vector<double> out(ticks.size() * cells.size());
    for (size_t j = 0; j < cells.size(); j++) {
      const auto& f = functions_.get(cells[j].first, cells[j].second);
      if (!f) throw std::invalid_argument("no time function for a voxel");
      for (size_t k = 0; k < ticks.size(); k++) out[k * cells.size() + j] = (*f)(ticks[k]);
    }
    return out;
  }

 private:
  Grid<Fn> functions_;
};

class PhaseSin : public Controller {  // controllers/PhaseSin.java:50-66
 public:
  PhaseSin(double frequency, double amplitude, Grid<double> phases)
      : frequency_(frequency), amplitude_(amplitude), phases_(std::move(phases)) {}
  int kind() const override { return VSR_CTRL_PHASE_SIN; }
  std::vector<double> params(const Grid<Voxel>& voxels, const std::vector<double>&) const override {
    std::vector<double> p = {frequency_, amplitude_};
    for (int y = 0; y < voxels.getH(); y++)
      for (int x = 0; x < voxels.getW(); x++)
        if (voxels.get(x, y)) {
          if (!phases_.get(x, y)) throw std::invalid_argument("no phase for a voxel");
          p.push_back(*phases_.get(x, y));
        }
    return p;
  }

 private:
  double frequency_, amplitude_;
  Grid<double> phases_;
};

class MultiLayerPerceptron {  // controllers/MultiLayerPerceptron.java
 public:
  enum ActivationFunction { RELU = 0, SIGMOID, SIN, TANH, SIGN, IDENTITY };  // :89-95
  MultiLayerPerceptron(ActivationFunction a, int nIn, std::vector<int> inner, int nOut, std::vector<double> weights)
      : activation_(a), neurons_(countNeurons(nIn, inner, nOut)), weights_(std::move(weights)) {
    if (int(weights_.size()) != countWeights(neurons_))  // :56-62
      throw std::invalid_argument("Wrong number of weights: " + std::to_string(countWeights(neurons_)) + " expected, " +
                                  std::to_string(weights_.size()) + " found");
  }
  static std::vector<int> countNeurons(int nIn, const std::vector<int>& inner, int nOut) {  // :97-104
    std::vector<int> n = {nIn};
    n.insert(n.end(), inner.begin(), inner.end());
    n.push_back(nOut);
    return n;
  }
  static int countWeights(const std::vector<int>& n) {  // :106-112
    int c = 0;
    for (size_t i = 1; i < n.size(); i++) c += n[i] * (n[i - 1] + 1);
    return c;
  }
  // flat <-> [layer][neuron][bias, w...] (:139-166)
  static std::vector<std::vector<std::vector<double>>> unflat(const std::vector<double>& flat, const std::vector<int>& n) {
    std::vector<std::vector<std::vector<double>>> u(n.size() - 1);
    size_t c = 0;
    for (size_t i = 1; i < n.size(); i++) {
      u[i - 1].assign(n[i], std::vector<double>(n[i - 1] + 1));
      for (int j = 0; j < n[i]; j++)
        for (int k = 0; k < n[i - 1] + 1; k++) u[i - 1][j][k] = flat[c++];
    }
    return u;
  }
  static std::vector<double> flat(const std::vector<std::vector<std::vector<double>>>& u, const std::vector<int>& n) {
    std::vector<double> f;
    for (size_t i = 1; i < n.size(); i++)
      for (int j = 0; j < n[i]; j++)
        for (int k = 0; k < n[i - 1] + 1; k++) f.push_back(u[i - 1][j][k]);
    return f;
  }
  // MultiLayerPerceptron.apply (:168-189), host-side evaluation of the description (the batched
  // episode evaluates the same function on the device): activation on the inputs too, bias first.
  std::vector<double> apply(const std::vector<double>& input) const {
    if (int(input.size()) != neurons_[0])
      throw std::invalid_argument("Expected input length is " + std::to_string(neurons_[0]) + ": found " + std::to_string(input.size()));
    std::vector<double> a(input.size());
    for (size_t k = 0; k < input.size(); k++) a[k] = act(input[k]);
    size_t c = 0;
    for (size_t i = 1; i < neurons_.size(); i++) {
      std::vector<double> b(neurons_[i]);
      for (int j = 0; j < neurons_[i]; j++) {
        double sum = weights_[c++];
        for (int k = 1; k < neurons_[i - 1] + 1; k++) sum = sum + a[k - 1] * weights_[c++];
        b[j] = act(sum);
      }
      a.swap(b);
    }
    return a;
  }
  int getInputDimension() const { return neurons_.front(); }
  int getOutputDimension() const { return neurons_.back(); }
  const std::vector<int>& getNeurons() const { return neurons_; }
  const std::vector<double>& getParams() const { return weights_; }
  ActivationFunction getActivation() const { return activation_; }

 private:
  double act(double x) const {
    switch (activation_) {
      case RELU: return x < 0 ? 0.0 : x;
      case SIGMOID: return 1.0 / (1.0 + std::exp(-x));
      case SIN: return std::sin(x);
      case TANH: return std::tanh(x);
      case SIGN: return x > 0 ? 1.0 : (x < 0 ? -1.0 : x);
      default: return x;
    }
  }
  ActivationFunction activation_;
  std::vector<int> neurons_;
  std::vector<double> weights_;
};

class CentralizedSensing : public Controller {  // controllers/CentralizedSensing.java:68-127
 public:
  CentralizedSensing(const Grid<Voxel>& voxels, MultiLayerPerceptron function)
      : nOfInputs_(nOfInputs(voxels)), nOfOutputs_(nOfOutputs(voxels)), function_(std::move(function)) {
    if (function_.getInputDimension() != nOfInputs_ || function_.getOutputDimension() != nOfOutputs_)  // :121-127
      throw std::invalid_argument("Wrong dimension of input or output in provided function: R^" + std::to_string(nOfInputs_) + "->R^" +
                                  std::to_string(nOfOutputs_) + " expected, R^" + std::to_string(function_.getInputDimension()) +
                                  "->R^" + std::to_string(function_.getOutputDimension()) + " found");
  }
  static int nOfInputs(const Grid<Voxel>& voxels) {
    int n = 0;
    for (auto& v : voxels.values())
      if (v)
        for (auto& s : v->getSensors()) n += int(s->getDomains().size());
    return n;
  }
  static int nOfOutputs(const Grid<Voxel>& voxels) { return int(voxels.count()); }
  int kind() const override { return VSR_CTRL_MLP; }
  std::vector<double> params(const Grid<Voxel>&, const std::vector<double>&) const override { return function_.getParams(); }
  const MultiLayerPerceptron& getFunction() const { return function_; }

 private:
  int nOfInputs_, nOfOutputs_;
  MultiLayerPerceptron function_;
};

// controllers/DistributedSensing.java (and ...NonDirectional): one MultiLayerPerceptron per voxel
// (Starter.java:176-190), fed with the voxel's readings and `signals` values from each of the N, E, S, W
// neighbours' previous outputs; output 0 is the control signal.
class DistributedSensing : public Controller {
 public:
  DistributedSensing(const Grid<Voxel>& voxels, int signals, bool directional = true)
      : signals_(signals), directional_(directional), w_(voxels.getW()), h_(voxels.getH()),
        functions_(size_t(voxels.getW()) * voxels.getH()) {
    for (int y = 0; y < h_; y++)
      for (int x = 0; x < w_; x++) {
        auto v = voxels.get(x, y);
        nIn_.push_back(v ? nOfInputs(*v, signals) : 0);
        nOut_.push_back(v ? nOfOutputs(signals, directional) : 0);
      }
  }
  static int nOfInputs(const Voxel& v, int signals) {  // :141-143
    int n = signals * 4;
    for (auto& s : v.getSensors()) n += int(s->getDomains().size());
    return n;
  }
  static int nOfOutputs(int signals, bool directional = true) { return directional ? 1 + signals * 4 : 1 + signals; }  // :145-147
  int nOfInputs(int x, int y) const { return nIn_[size_t(y) * w_ + x]; }
  int nOfOutputs(int x, int y) const { return nOut_[size_t(y) * w_ + x]; }
  void setFunction(int x, int y, MultiLayerPerceptron f) {
    if (f.getInputDimension() != nOfInputs(x, y) || f.getOutputDimension() != nOfOutputs(x, y))
      throw std::invalid_argument("Wrong dimension of input or output in the function of voxel (" + std::to_string(x) + "," +
                                  std::to_string(y) + ")");
    functions_[size_t(y) * w_ + x] = std::move(f);
  }
  int getSignals() const { return signals_; }
  int kind() const override { return directional_ ? VSR_CTRL_DISTRIBUTED : VSR_CTRL_DISTRIBUTED_ND; }
  const std::optional<MultiLayerPerceptron>& getFunction(int x, int y) const { return functions_[size_t(y) * w_ + x]; }
  std::vector<double> params(const Grid<Voxel>& voxels, const std::vector<double>&) const override {
    std::vector<double> out;
    for (int y = 0; y < h_; y++)
      for (int x = 0; x < w_; x++) {
        if (!voxels.get(x, y)) continue;
        auto& f = functions_[size_t(y) * w_ + x];
        if (!f) throw Unsupported("DistributedSensing: the accelerated path needs a MultiLayerPerceptron for every voxel");
        out.insert(out.end(), f->getParams().begin(), f->getParams().end());
      }
    return out;
  }

 private:
  int signals_;
  bool directional_;
  int w_, h_;
  std::vector<int> nIn_, nOut_;
  std::vector<std::optional<MultiLayerPerceptron>> functions_;
};

class Robot {  // core/objects/Robot.java:57-64
 public:
  Robot(ControllerPtr controller, Grid<Voxel> voxels) : controller_(std::move(controller)), voxels_(std::move(voxels)) {}
  const Controller& getController() const { return *controller_; }
  const Grid<Voxel>& getVoxels() const { return voxels_; }

 private:
  ControllerPtr controller_;
  Grid<Voxel> voxels_;
};

// ---------------------------------------------------------------- RobotUtils
namespace RobotUtils {
inline bool match(const std::string& pattern, const std::string& s, std::smatch& m) {  // util/Utils.java:147-164
  return std::regex_match(s, m, std::regex(pattern));
}
inline Grid<bool> buildShape(const std::string& name) {  // util/RobotUtils.java:198-242
  std::smatch m;
  auto grid = [](int w, int h, const std::function<bool(int, int)>& f) {
    return Grid<bool>::create(w, h, [&](int x, int y) { return std::optional<bool>(f(x, y)); });
  };
  if (match(R"((box|worm)-(\d+)x(\d+))", name, m)) return grid(std::stoi(m[2]), std::stoi(m[3]), [](int, int) { return true; });
  if (match(R"(biped-(\d+)x(\d+))", name, m)) {
    int w = std::stoi(m[1]), h = std::stoi(m[2]);
    return grid(w, h, [=](int x, int y) { return !(y < h / 2 && x >= w / 4 && x < w * 3 / 4); });
  }
  if (match(R"(tripod-(\d+)x(\d+))", name, m)) {
    int w = std::stoi(m[1]), h = std::stoi(m[2]);
    return grid(w, h, [=](int x, int y) { return !(y < h / 2 && x != 0 && x != w - 1 && x != w / 2); });
  }
  if (match(R"(ball-(\d+))", name, m)) {
    int d = std::stoi(m[1]);
    double c = (d - 1) / 2.0;
    return grid(d, d, [=](int x, int y) {
      return std::floor(std::sqrt((x - c) * (x - c) + (y - c) * (y - c)) + 0.5) <= std::floor(d / 2.0);
    });
  }
  if (match(R"(comb-(\d+)x(\d+))", name, m)) {
    int w = std::stoi(m[1]), h = std::stoi(m[2]);
    return grid(w, h, [=](int x, int y) { return y >= h / 2 || x % 2 == 0; });
  }
  if (match(R"(t-(\d+)x(\d+))", name, m)) {
    int w = std::stoi(m[1]), h = std::stoi(m[2]);
    int pad = int(std::floor(std::floor(w / 2.0) / 2.0));
    return grid(w, h, [=](int x, int y) { return y == 0 || (x >= pad && x < h - pad - 1); });
  }
  throw std::invalid_argument("Unknown body name: " + name);
}
inline SensorPtr sensor(const std::string& n, double x = 0, double y = 0) {  // PREDEFINED_SENSORS, :38-57 (all but the lidars)
  auto vel = [](double m, int axes) { return std::make_shared<Velocity>(true, m, axes); };
  if (n == "t") return std::make_shared<Average>(std::make_shared<Touch>(), 0.25);
  if (n == "a") return std::make_shared<SoftNormalization>(std::make_shared<AreaRatio>());
  if (n == "vx") return std::make_shared<SoftNormalization>(std::make_shared<Average>(vel(8, Velocity::X), 0.5));
  if (n == "vy") return std::make_shared<SoftNormalization>(std::make_shared<Average>(vel(8, Velocity::Y), 0.5));
  if (n == "vxy") return std::make_shared<SoftNormalization>(std::make_shared<Average>(vel(8, Velocity::X | Velocity::Y), 0.5));
  if (n == "ax") return std::make_shared<SoftNormalization>(std::make_shared<Trend>(vel(4, Velocity::X), 0.5));
  if (n == "ay") return std::make_shared<SoftNormalization>(std::make_shared<Trend>(vel(4, Velocity::Y), 0.5));
  if (n == "axy") return std::make_shared<SoftNormalization>(std::make_shared<Trend>(vel(4, Velocity::X | Velocity::Y), 0.5));
  if (n == "r") return std::make_shared<SoftNormalization>(std::make_shared<Angle>());
  if (n == "px") return std::make_shared<Constant>(x);
  if (n == "py") return std::make_shared<Constant>(y);
  if (n == "m") return std::make_shared<Malfunction>();
  if (n == "cpg") return std::make_shared<Normalization>(std::make_shared<TimeFunctionSin>(-1.0));
  if (n == "l5" || n == "l1") {  // lidarSide, :244-255
    Lidar::Side side = (y > x && y > (1 - x)) ? Lidar::N : (y >= x && y <= (1 - x)) ? Lidar::W : (y < x && y < (1 - x)) ? Lidar::S : Lidar::E;
    return std::make_shared<Normalization>(Lidar::ofSide(10.0, side, n == "l5" ? 5 : 1));
  }
  throw std::invalid_argument("Unknown sensor name: " + n);
}
inline std::function<Grid<Voxel>(const Grid<bool>&)> buildSensorizingFunction(const std::string& name) {  // :124-196
  std::smatch m;
  const std::string keys = "a|ax|axy|ay|cpg|l1|l5|m|px|py|r|t|vx|vxy|vy";
  if (match("uniform-((" + keys + ")(\\+(" + keys + "))*)-(\\d+(\\.\\d+)?)", name, m)) {
    const double sigma = std::stod(m[5]);
    std::vector<std::string> names;
    std::string list = m[1], cur;
    for (char c : list + "+") {
      if (c == '+') {
        names.push_back(cur);
        cur.clear();
      } else cur += c;
    }
    return [names, sigma](const Grid<bool>& body) {
      return Grid<Voxel>::create(body.getW(), body.getH(), [&](int x, int y) -> std::optional<Voxel> {
        if (!body.get(x, y) || !*body.get(x, y)) return std::nullopt;
        std::vector<SensorPtr> ss;
        // RobotUtils.sensor (:249-263): x / (w - 1), y / (h - 1)
        const double nx = double(x) / (double(body.getW()) - 1.0), ny = double(y) / (double(body.getH()) - 1.0);
        for (auto& n : names) {
          SensorPtr sp = sensor(n, nx, ny);
          ss.push_back(sigma == 0 ? sp : std::make_shared<Noisy>(sp, sigma, 0));  // :172-175
        }
        return Voxel(ss);
      });
    };
  }
  if (name == "empty")
    return [](const Grid<bool>& body) {
      return Grid<Voxel>::create(body.getW(), body.getH(), [&](int x, int y) -> std::optional<Voxel> {
        if (!body.get(x, y) || !*body.get(x, y)) return std::nullopt;
        return Voxel({});
      });
    };
  if (match(R"(spinedTouch(Sighted)?-([tf])-([tf])-(\d+(\.\d+)?))", name, m)) {  // :134-165
    const double sigma = std::stod(m[4]);
    const bool sighted = m[1].matched, cpg = m[2] == "t", malfunction = m[3] == "t";
    return [sighted, cpg, malfunction, sigma](const Grid<bool>& body) {
      return Grid<Voxel>::create(body.getW(), body.getH(), [&](int x, int y) -> std::optional<Voxel> {
        if (!body.get(x, y) || !*body.get(x, y)) return std::nullopt;
        std::vector<SensorPtr> ss;
        ss.push_back(sensor("a"));
        if (malfunction) ss.push_back(sensor("m"));
        if (y == 0) ss.push_back(sensor("t"));
        if (y == body.getH() - 1) ss.push_back(sensor("vxy"));
        if (x == body.getW() - 1 && y == body.getH() - 1 && cpg) ss.push_back(sensor("cpg"));
        if (sighted && x == body.getW() - 1)  // :163; RobotUtils.sensor normalises the position (:257-263)
          ss.push_back(sensor("l5", double(x) / (double(body.getW()) - 1.0), double(y) / (double(body.getH()) - 1.0)));
        if (sigma != 0)
          for (auto& sp : ss) sp = std::make_shared<Noisy>(sp, sigma, 0);
        return Voxel(ss);
      });
    };
  }
  if (match(R"(uniformAll-(\d+(\.\d+)?))", name, m)) {  // :177-189: every predefined sensor in TreeMap (sorted) order
    std::string all = keys;
    std::replace(all.begin(), all.end(), '|', '+');
    return buildSensorizingFunction("uniform-" + all + "-" + std::string(m[1]));
  }
  throw std::invalid_argument("Unknown sensorizing function name: " + name);
}
}  // namespace RobotUtils

// ---------------------------------------------------------------- Settings / Outcome
class Settings {  // org.dyn4j.dynamics.Settings as a value holder (Locomotion.java:50)
 public:
  Settings() {
    s_ = vsr_settings{};
    s_.step_frequency = 1.0 / 60.0; s_.velocity_iterations = 10; s_.position_iterations = 10;
    s_.linear_tolerance = 0.005; s_.angular_tolerance = 2.0 * M_PI / 180.0; s_.max_linear_correction = 0.2;
    s_.baumgarte = 0.2; s_.restitution_velocity = 1.0; s_.max_translation = 2.0; s_.max_rotation = 0.5 * M_PI;
    s_.gravity_x = 0.0; s_.gravity_y = -9.8; s_.ground_friction = 0.2; s_.ground_restitution = 0.0;
    s_.spring_mass_mode = VSR_SPRING_MASS_REDUCED;
    s_.self_collision = 1;
    s_.continuous_detection = 1;  // ContinuousDetectionMode.ALL, dyn4j's default
  }
  double getStepFrequency() const { return s_.step_frequency; }
  vsr_settings& raw() { return s_; }
  const vsr_settings& raw() const { return s_; }

 private:
  vsr_settings s_;
};

// ---------------------------------------------------------------- per-tick view (Observation path)
struct Point2 {
  double x, y;
};
struct BoundingBox {  // util/BoundingBox.java
  Point2 min, max;
};
class VoxelPoly {  // core/snapshots/VoxelPoly.java: the 4 outer vertexes of a voxel and its scalar state
 public:
  VoxelPoly(std::array<Point2, 4> v, double angle, Point2 vel, bool touching, double areaRatio, double areaRatioEnergy,
            double lastAppliedForce, double controlEnergy)
      : v_(v), angle_(angle), vel_(vel), touching_(touching), areaRatio_(areaRatio), areaRatioEnergy_(areaRatioEnergy),
        lastAppliedForce_(lastAppliedForce), controlEnergy_(controlEnergy) {}
  const std::array<Point2, 4>& vertexes() const { return v_; }
  BoundingBox boundingBox() const {  // Poly.java:51-64
    BoundingBox b{{v_[0].x, v_[0].y}, {v_[0].x, v_[0].y}};
    for (auto& p : v_) {
      b.min.x = std::min(b.min.x, p.x); b.min.y = std::min(b.min.y, p.y);
      b.max.x = std::max(b.max.x, p.x); b.max.y = std::max(b.max.y, p.y);
    }
    return b;
  }
  double area() const {  // Poly.java:40-49
    double a = 0;
    for (int i = 0; i < 4; i++) a = a + v_[i].x * (v_[(4 + i + 1) % 4].y - v_[(4 + i - 1) % 4].y);
    return 0.5 * std::fabs(a);
  }
  Point2 center() const { return {(v_[0].x + v_[1].x + v_[2].x + v_[3].x) / 4.0, (v_[0].y + v_[1].y + v_[2].y + v_[3].y) / 4.0}; }
  double getAngle() const { return angle_; }
  Point2 getLinearVelocity() const { return vel_; }
  bool isTouchingGround() const { return touching_; }
  double getAreaRatio() const { return areaRatio_; }
  double getAreaRatioEnergy() const { return areaRatioEnergy_; }
  double getLastAppliedForce() const { return lastAppliedForce_; }
  double getControlEnergy() const { return controlEnergy_; }

 private:
  std::array<Point2, 4> v_;
  double angle_;
  Point2 vel_;
  bool touching_;
  double areaRatio_, areaRatioEnergy_, lastAppliedForce_, controlEnergy_;
};
struct Observation {  // Outcome.java:38-39
  Grid<VoxelPoly> voxelPolies;
  double terrainHeight, computationTime;
};
namespace BehaviorUtils {  // behavior/BehaviorUtils.java (the geometric part)
inline Point2 center(const std::vector<VoxelPoly>& shapes) {  // :47-56
  double x = 0, y = 0;
  for (auto& s : shapes) {
    Point2 c = s.center();
    x = x + c.x;
    y = y + c.y;
  }
  return {x / double(shapes.size()), y / double(shapes.size())};
}
inline std::vector<bool> computeFootprint(const std::vector<VoxelPoly>& polies, int n) {  // :68-91
  if (polies.empty()) throw std::invalid_argument("Empty robot");
  double mn = INFINITY, mx = -INFINITY;
  for (auto& p : polies) {
    mn = std::min(mn, p.boundingBox().min.x);
    mx = std::max(mx, p.boundingBox().max.x);
  }
  std::vector<bool> mask(size_t(n), false);
  for (auto& p : polies) {
    if (!p.isTouchingGround()) continue;
    int lo = int(std::floor((p.boundingBox().min.x - mn) / (mx - mn) * double(n - 1) + 0.5));
    int hi = int(std::floor((p.boundingBox().max.x - mn) / (mx - mn) * double(n - 1) + 0.5));
    for (int x = lo; x <= std::min(hi, n - 1); x++) mask[size_t(x)] = true;
  }
  return mask;
}
}  // namespace BehaviorUtils

class Outcome {  // tasks/locomotion/Outcome.java: the record of the first + last Observation, optionally all of them
 public:
  explicit Outcome(const vsr_result& r) : r_(r) {}
  Outcome(const vsr_result& r, std::map<double, Observation> obs) : r_(r), observations_(std::move(obs)) {}
  // the per-tick part (needs Locomotion::apply(robots, true))
  const std::map<double, Observation>& getObservations() const {  // :190-192
    if (observations_.empty()) throw Unsupported("this Outcome holds the first/last observation only");
    return observations_;
  }
  static std::vector<VoxelPoly> polies(const Observation& o) {
    std::vector<VoxelPoly> out;
    for (auto& v : o.voxelPolies.values())
      if (v) out.push_back(*v);
    return out;
  }
  double getDistanceFromObservations() const {  // :152-166 on the map
    auto& m = getObservations();
    return BehaviorUtils::center(polies(m.rbegin()->second)).x - BehaviorUtils::center(polies(m.begin()->second)).x;
  }
  Outcome subOutcome(double startT, double endT) const {  // :202-204: subMap(startT, endT)
    std::map<double, Observation> sub;
    for (auto& kv : getObservations())
      if (kv.first >= startT && kv.first < endT) sub.emplace(kv.first, kv.second);
    return Outcome(r_, std::move(sub));
  }
  double getDistance() const { return r_.last_center_x - r_.first_center_x; }          // :152-166
  double getTime() const { return r_.t_last - r_.t_first; }                              // :194-196
  double getVelocity() const { return getDistance() / getTime(); }                       // :198-200
  double getAreaRatioEnergy() const { return r_.area_ratio_energy_last - r_.area_ratio_energy_first; }  // :42-58
  double getControlEnergy() const { return r_.control_energy_last - r_.control_energy_first; }          // :126-142
  double getControlPower() const { return getControlEnergy() / getTime(); }
  double getAreaRatioPower() const { return getAreaRatioEnergy() / getTime(); }
  double getCorrectedEfficiency() const { return getDistance() / (1.0 + getControlPower() * getTime()); }  // :148-150
  const vsr_result& record() const { return r_; }

 private:
  vsr_result r_;
  std::map<double, Observation> observations_;
};

// ---------------------------------------------------------------- flattening
struct BatchDesc {  // owns the flat arrays behind a vsr_batch_desc
  vsr_batch_desc desc{};
  std::vector<vsr_shape> shapes;
  std::vector<std::vector<uint8_t>> masks;
  std::vector<std::vector<int32_t>> sensorOffsets, hidden;
  std::vector<std::vector<vsr_sensor>> sensors;
  std::vector<int32_t> robotShape;
  std::vector<double> params, xs, ys;
  std::vector<size_t> paramOffsets;
  std::vector<int> nVoxels;
  int nTicks = 0;
};

inline bool sameSensor(const vsr_sensor& a, const vsr_sensor& b) {
  return a.base == b.base && a.axes == b.axes && a.rotated == b.rotated && a.aggregator == b.aggregator &&
         a.soft_norm == b.soft_norm && a.max_velocity_norm == b.max_velocity_norm && a.interval == b.interval &&
         a.noise_sigma == b.noise_sigma && a.noise_seed == b.noise_seed &&
         std::memcmp(a.lidar_dirs, b.lidar_dirs, sizeof a.lidar_dirs) == 0;
}

class Locomotion {  // tasks/locomotion/Locomotion.java
 public:
  static constexpr double TERRAIN_BORDER_HEIGHT = 100, TERRAIN_BORDER_WIDTH = 10;
  static constexpr int TERRAIN_LENGTH = 2000;
  using Profile = std::pair<std::vector<double>, std::vector<double>>;

  Locomotion(double finalT, Profile groundProfile, Settings settings)
      : Locomotion(finalT, groundProfile, std::numeric_limits<double>::quiet_NaN(), settings) {}
  Locomotion(double finalT, Profile groundProfile, double initialPlacement, Settings settings)
      : finalT_(finalT), profile_(std::move(groundProfile)), settings_(settings) {
    auto& xs = profile_.first;
    if (xs.size() != profile_.second.size()) throw std::invalid_argument("xs[] and ys[] must have the same length");  // Ground.java:49-51
    if (xs.size() < 2) throw std::invalid_argument("There must be at least 2 points");                                // :52-54
    for (size_t i = 1; i < xs.size(); i++)
      if (xs[i] < xs[i - 1]) throw std::invalid_argument("x coordinates must be sorted");                            // :55-58
    initialPlacement_ = std::isnan(initialPlacement) ? xs[1] + 1.0 : initialPlacement;                               // Locomotion.java:50-52
  }

  static Profile createTerrain(const std::string& name) {  // :61-146
    const double L = TERRAIN_LENGTH, BW = TERRAIN_BORDER_WIDTH, BH = TERRAIN_BORDER_HEIGHT;
    const std::string num = R"([0-9]+(\.[0-9]+)?)";
    std::smatch m;
    if (name == "flat") return {{0, BW, L - BW, L}, {BH, 5, 5, BH}};
    if (RobotUtils::match("flatWithStart-([0-9]+)", name, m)) {
      JavaRandom rnd(std::stoll(m[1]));
      int n = rnd.nextInt(10) + 10;
      for (int i = 0; i < n; i++) rnd.nextDouble();
      double angle = M_PI / 18.0 * (rnd.nextDouble() * 2.0 - 1.0);
      double start = Voxel::SIDE_LENGTH * 8.0;
      return {{0, BW, BW + start, L - BW, L}, {BH, 5 + start * std::sin(angle), 5, 5, BH}};
    }
    bool hilly = RobotUtils::match("hilly-(" + num + ")-(" + num + ")-([0-9]+)", name, m);
    bool steppy = !hilly && RobotUtils::match("steppy-(" + num + ")-(" + num + ")-([0-9]+)", name, m);
    if (hilly || steppy) {
      double h = std::stod(m[1]), w = std::stod(m[3]);
      JavaRandom rnd(std::stoll(m[5]));
      std::vector<double> xs = {0, BW}, ys = {BH, 0};
      while (xs.back() < L - BW) {
        xs.push_back(xs.back() + std::max(1.0, (rnd.nextGaussian() * 0.25 + 1) * w));
        if (steppy) {
          xs.push_back(xs.back() + 0.5);
          ys.push_back(ys.back());
        }
        ys.push_back(ys.back() + rnd.nextGaussian() * h);
      }
      xs.push_back(xs.back() + BW);
      ys.push_back(BH);
      return {xs, ys};
    }
    if (RobotUtils::match("downhill-(" + num + ")", name, m)) {
      double dY = (L - 2 * BW) * std::sin(std::stod(m[1]) / 180 * M_PI);
      return {{0, BW, L - BW, L}, {BH + dY, 5 + dY, 5, BH}};
    }
    if (RobotUtils::match("uphill-(" + num + ")", name, m)) {
      double dY = (L - 2 * BW) * std::sin(std::stod(m[1]) / 180 * M_PI);
      return {{0, BW, L - BW, L}, {BH, 5, 5 + dY, BH + dY}};
    }
    throw std::invalid_argument("Unknown terrain name: " + name);
  }

  // Flatten robots + this task into a vsr_batch_desc (what the Java side would hand over).
  std::unique_ptr<BatchDesc> compile(const std::vector<Robot>& robots, uint32_t precision = VSR_PRECISION_F32,
                                     size_t trajectoryRobots = 0) const {
    if (robots.empty()) throw std::invalid_argument("empty batch");
    auto b = std::make_unique<BatchDesc>();
    vsr_batch_desc& d = b->desc;
    d.abi_version = VSR_ABI_VERSION;
    d.precision = precision;
    d.n_robots = robots.size();
    d.controller_kind = robots[0].getController().kind();
    d.settings = settings_.raw();
    d.final_t = finalT_;
    b->xs = profile_.first;
    b->ys = profile_.second;
    d.initial_placement = initialPlacement_;
    d.trajectory_robots = trajectoryRobots;
    std::vector<double> ticks = tickTimes(finalT_, d.settings.step_frequency);
    b->nTicks = int(ticks.size());
    b->paramOffsets.push_back(0);
    bool haveMaterial = false;
    for (const Robot& r : robots) {
      if (r.getController().kind() != d.controller_kind) throw std::invalid_argument("one controller kind per batch");
      const Grid<Voxel>& vox = r.getVoxels();
      std::vector<uint8_t> mask;
      std::vector<int32_t> offs = {0}, hid;
      std::vector<vsr_sensor> sens;
      for (auto& v : vox.values()) {
        mask.push_back(v ? 1 : 0);
        if (!v) continue;
        for (auto& s : v->getSensors()) sens.push_back(s->encode());
        offs.push_back(int32_t(sens.size()));
        if (!haveMaterial) {
          d.material = v->material();
          haveMaterial = true;
        } else if (std::memcmp(&d.material, &v->material(), sizeof(vsr_material)) != 0) {
          throw Unsupported("one voxel material per batch on the accelerated path");
        }
      }
      if (d.controller_kind == VSR_CTRL_MLP) {
        auto& f = static_cast<const CentralizedSensing&>(r.getController()).getFunction();
        hid.assign(f.getNeurons().begin() + 1, f.getNeurons().end() - 1);
        d.mlp_activation = int(f.getActivation());
      } else if (d.controller_kind >= VSR_CTRL_DISTRIBUTED) {
        auto& ds = static_cast<const DistributedSensing&>(r.getController());
        bool first = true;
        for (int y = 0; y < vox.getH(); y++)
          for (int x = 0; x < vox.getW(); x++) {
            if (!vox.get(x, y)) continue;
            auto& f = ds.getFunction(x, y);
            if (!f) throw Unsupported("DistributedSensing: the accelerated path needs a MultiLayerPerceptron for every voxel");
            std::vector<int32_t> h2(f->getNeurons().begin() + 1, f->getNeurons().end() - 1);
            if (first) hid = h2;
            else if (hid != h2 || d.mlp_activation != int(f->getActivation()))
              throw Unsupported("DistributedSensing: one activation and one set of hidden layer sizes per robot");
            d.mlp_activation = int(f->getActivation());
            first = false;
          }
        d.distributed_signals = ds.getSignals();
      }
      // find or add the shape class
      int cls = -1;
      for (size_t k = 0; k < b->masks.size() && cls < 0; k++) {
        if (b->shapes[k].w != vox.getW() || b->shapes[k].h != vox.getH() || b->masks[k] != mask || b->sensorOffsets[k] != offs ||
            b->hidden[k] != hid || b->sensors[k].size() != sens.size())
          continue;
        bool same = true;
        for (size_t q = 0; q < sens.size() && same; q++) same = sameSensor(sens[q], b->sensors[k][q]);
        if (same) cls = int(k);
      }
      if (cls < 0) {
        cls = int(b->masks.size());
        vsr_shape sh{};
        sh.w = vox.getW();
        sh.h = vox.getH();
        b->shapes.push_back(sh);
        b->masks.push_back(mask);
        b->sensorOffsets.push_back(offs);
        b->sensors.push_back(sens);
        b->hidden.push_back(hid);
        b->nVoxels.push_back(int(offs.size()) - 1);
      }
      b->robotShape.push_back(cls);
      std::vector<double> p = r.getController().params(vox, ticks);
      b->params.insert(b->params.end(), p.begin(), p.end());
      b->paramOffsets.push_back(b->params.size());
    }
    if (!haveMaterial) throw std::invalid_argument("robot without voxels");
    for (size_t k = 0; k < b->shapes.size(); k++) {  // pointers only once the vectors stopped growing
      b->shapes[k].mask = b->masks[k].data();
      b->shapes[k].sensor_offsets = b->sensorOffsets[k].data();
      b->shapes[k].sensors = b->sensors[k].data();
      b->shapes[k].n_hidden = int32_t(b->hidden[k].size());
      b->shapes[k].hidden = b->hidden[k].data();
    }
    d.shapes = b->shapes.data();
    d.n_shapes = int32_t(b->shapes.size());
    d.robot_shape = b->robotShape.data();
    d.params = b->params.data();
    d.param_offsets = b->paramOffsets.data();
    d.terrain_xs = b->xs.data();
    d.terrain_ys = b->ys.data();
    d.terrain_n = int32_t(b->xs.size());
    return b;
  }

  // The batch entry the reference lacks: List<Outcome> apply(List<Robot>).  One kernel launch.
  std::vector<Outcome> apply(const std::vector<Robot>& robots, int device = 0) const {
    auto bd = compile(robots);
    vsr_batch* h = nullptr;
    check(vsr_batch_create(&bd->desc, &h));
    struct Guard {
      vsr_batch* h;
      ~Guard() { vsr_batch_destroy(h); }
    } guard{h};
    check(vsr_batch_upload(h, device, nullptr));
    check(vsr_batch_run(h, nullptr));
    std::vector<vsr_result> res(robots.size());
    check(vsr_batch_results(h, res.data(), res.size()));
    std::vector<Outcome> out;
    for (auto& r : res) out.emplace_back(r);
    return out;
  }
  Outcome apply(const Robot& robot, int device = 0) const { return apply(std::vector<Robot>{robot}, device)[0]; }  // :168-207

  // ... with the observations map of Locomotion.java:193-203 built from the device's per-tick dumps
  // (vsr_batch_trajectory + vsr_batch_voxel_trace); the scalar getters follow Voxel.java:508-556
  std::vector<Outcome> applyWithObservations(const std::vector<Robot>& robots, int device = 0) const {
    auto bd = compile(robots, VSR_PRECISION_F32, robots.size());
    vsr_batch* h = nullptr;
    check(vsr_batch_create(&bd->desc, &h));
    struct Guard {
      vsr_batch* h;
      ~Guard() { vsr_batch_destroy(h); }
    } guard{h};
    check(vsr_batch_upload(h, device, nullptr));
    check(vsr_batch_run(h, nullptr));
    std::vector<vsr_result> res(robots.size());
    check(vsr_batch_results(h, res.data(), res.size()));
    const std::vector<double> tt = tickTimes(finalT_, settings_.getStepFrequency());
    const size_t nt = tt.size();
    std::vector<Outcome> out;
    for (size_t r = 0; r < robots.size(); r++) {
      const Grid<Voxel>& vox = robots[r].getVoxels();
      const size_t nv = vox.count();
      std::vector<double> traj(nt * nv * 4 * 6), trace(nt * nv * 2);
      if (vsr_batch_trajectory(h, r, traj.data(), traj.size()) < 0 || vsr_batch_voxel_trace(h, r, trace.data(), trace.size()) < 0)
        throw std::runtime_error(vsr_last_error());
      const vsr_material& m = vox.values()[size_t(std::find_if(vox.values().begin(), vox.values().end(), [](auto& v) { return bool(v); }) - vox.values().begin())]->material();
      const double hh = m.side_length * m.mass_side_length_ratio / 2.0;
      static const double lx[4] = {-1, 1, 1, -1}, ly[4] = {1, 1, -1, -1};
      std::vector<double> are(nv, 0.0), ce(nv, 0.0), prevF(nv, 0.0);
      std::map<double, Observation> obs;
      for (size_t k = 0; k < nt; k++) {
        double cxs = 0;
        size_t j = 0;
        Grid<VoxelPoly> polies = Grid<VoxelPoly>::create(vox.getW(), vox.getH(), [&](int x, int y) -> std::optional<VoxelPoly> {
          (void)x; (void)y;
          return std::nullopt;
        });
        for (int y = 0; y < vox.getH(); y++)
          for (int x = 0; x < vox.getW(); x++) {
            if (!vox.get(x, y)) continue;
            const double* b = &traj[((k * nv + j) * 4) * 6];
            std::array<Point2, 4> vs;
            double mvx = 0, mvy = 0, bcx = 0;
            for (int q = 0; q < 4; q++) {
              const double c = std::cos(b[6 * q + 2]), s2 = std::sin(b[6 * q + 2]);
              vs[size_t(q)] = {b[6 * q] + c * lx[q] * hh - s2 * ly[q] * hh, b[6 * q + 1] + s2 * lx[q] * hh + c * ly[q] * hh};
              mvx = mvx + b[6 * q + 3];
              mvy = mvy + b[6 * q + 4];
              bcx += b[6 * q];
            }
            cxs += bcx / 4.0;
            double area = 0;
            for (int i = 0; i < 4; i++) area = area + vs[size_t(i)].x * (vs[size_t((i + 1) % 4)].y - vs[size_t((i + 3) % 4)].y);
            const double ratio = 0.5 * std::fabs(area) / m.side_length / m.side_length;
            const double ang = (std::atan2(b[6 + 1] - b[1], b[6] - b[0]) + std::atan2(b[12 + 1] - b[18 + 1], b[12] - b[18])) / 2.0;
            are[j] += ratio * ratio;               // Voxel.act (Voxel.java:197-203) runs before the controller:
            ce[j] += prevF[j] * prevF[j];          // the energy of tick k holds the force applied at tick k - 1
            const double f = trace[(k * nv + j) * 2 + 1];
            polies.set(x, y, VoxelPoly(vs, ang, {mvx / 4.0, mvy / 4.0}, trace[(k * nv + j) * 2] != 0.0, ratio, are[j], f, ce[j]));
            prevF[j] = f;
            j++;
          }
        obs.emplace(tt[k], Observation{polies, groundYAt(cxs / double(nv)), 0.0});
      }
      out.emplace_back(res[r], std::move(obs));
    }
    return out;
  }
  double groundYAt(double x) const {  // Ground.yAt (Ground.java:104-111)
    for (size_t i = 1; i < profile_.first.size(); i++)
      if (profile_.first[i - 1] <= x && x <= profile_.first[i])
        return (x - profile_.first[i - 1]) * (profile_.second[i] - profile_.second[i - 1]) / (profile_.first[i] - profile_.first[i - 1]) +
               profile_.second[i - 1];
    return std::nan("");
  }

 private:
  static void check(int rc) {
    if (rc == VSR_OK) return;
    std::string msg = vsr_last_error();
    if (rc == VSR_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    if (rc == VSR_ERR_UNSUPPORTED) throw Unsupported(msg);
    throw std::runtime_error("vsr error " + std::to_string(rc) + ": " + msg);
  }
  double finalT_;
  Profile profile_;
  double initialPlacement_;
  Settings settings_;
};

}  // namespace hmsrobots
