"""B200-native batched stepper for 2D-VSR-Sim's Locomotion hot path.

The package name starts with a digit, so import it with
``importlib.import_module("2dhmsr_b200")``.  The compute path is the CUDA library
``libvsr_b200.so`` behind the C ABI of ``include/vsr_b200.h``; there is no CPU
fallback.
"""
from . import _abi as abi  # noqa: F401
from .hmsrobots import (  # noqa: F401
    Angle, AppliedForce, Constant, ControlPower, Crumpling, Lidar, Malfunction, Noisy, Normalization, TimeFunction,
    AreaRatio, Average, BatchDesc, BreakableVoxel, CentralizedSensing, Controller, DeviceBatch, DistributedSensing, DistributedSensingNonDirectional,
    DoubleRange, DynamicNormalization, Grid,
    JavaRandom, Locomotion, MultiLayerPerceptron, Outcome, PhaseSin, RESULT_DTYPE, Robot, RobotShape, RobotUtils,
    Sensor, Settings, Snapshot, SoftNormalization, TimeFunctions, Touch, Trend, Velocity, Voxel, VsrError,
    compile_batch, tick_times,
)
from .behavior import (  # noqa: F401
    BehaviorUtils, BoundingBox, Footprint, Gait, Observation, VoxelPoly, build_observations, ground_y_at,
)
from .devolocomotion import (  # noqa: F401
    DevoLocomotion, DevoOutcome, DevoStageOutcome, DistanceBasedDevoLocomotion, TimeBasedDevoLocomotion,
)
