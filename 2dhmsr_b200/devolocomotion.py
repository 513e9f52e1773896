"""tasks/devolocomotion: locomotion tasks in which the robot is re-developed during the episode.

A *solution* is a ``UnaryOperator<Robot>``: called with ``None`` it returns the first robot, called with
the current robot it returns the next one.  The new robot is put where the old one stood (its bounding
box's min x, one unit above the ground) and the task clock keeps running (DevoLocomotion.java,
TimeBasedDevoLocomotion.java, DistanceBasedDevoLocomotion.java).

On the device a stage is one episode of the batched stepper with ``t_start`` = the clock at the stage's
beginning and a per-robot placement; the robots of all solutions of a stage run in one launch.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _abi as A
from .hmsrobots import DeviceBatch, Locomotion, Outcome, Robot, Settings, compile_batch, tick_times


class DevoStageOutcome:
    """DevoOutcome.DevoStageOutcome: the stage's robot and its velocity, time, distance."""

    def __init__(self, robot: Robot, outcome: Outcome):
        self.robot = robot
        self.velocity, self.time, self.distance = outcome.get_velocity(), outcome.get_time(), outcome.get_distance()
        self.outcome = outcome


class DevoOutcome:
    """tasks/devolocomotion/DevoOutcome.java"""

    def __init__(self, outcomes: Optional[List[DevoStageOutcome]] = None):
        self.outcome_list: List[DevoStageOutcome] = list(outcomes) if outcomes else []

    def add_devo_stage_outcome(self, o: DevoStageOutcome) -> None:
        self.outcome_list.append(o)

    def get_devo_outcomes(self) -> List[DevoStageOutcome]:
        return self.outcome_list

    def get_distances(self) -> List[float]:
        return [o.distance for o in self.outcome_list]

    def get_robots(self) -> List[Robot]:
        return [o.robot for o in self.outcome_list]

    def get_times(self) -> List[float]:
        return [o.time for o in self.outcome_list]

    def get_velocities(self) -> List[float]:
        return [o.velocity for o in self.outcome_list]


class DevoLocomotion:
    """tasks/devolocomotion/DevoLocomotion.java"""

    def __init__(self, max_t: float, ground_profile, settings: Settings, initial_placement: Optional[float] = None):
        self.max_t = float(max_t)
        self.ground_profile = [list(map(float, ground_profile[0])), list(map(float, ground_profile[1]))]
        self.initial_placement = (self.ground_profile[0][1] + Locomotion.INITIAL_PLACEMENT_X_GAP
                                  if initial_placement is None else float(initial_placement))
        self.settings = settings

    def _stage(self, robots: Sequence[Robot], placements: Sequence[float], t_start: float, t_end: float,
               device: int, precision: int, trajectories: bool = False):
        """One stage for a set of robots: ticks t_start + dT ... while t < t_end."""
        bd = compile_batch(robots, t_end, self.ground_profile, self.initial_placement, self.settings, precision,
                           len(robots) if trajectories else 0, t_start, list(placements))
        with DeviceBatch(bd) as db:
            db.upload(device).run()
            res, box = db.results(), db.final_bbox()
            traj = [db.trajectory(i) for i in range(len(robots))] if trajectories else None
            trace = [db.voxel_trace(i) for i in range(len(robots))] if trajectories else None
        bad = np.flatnonzero(res["status"] != 0)
        if bad.size:  # truncated physics or a non-finite state must not become a silent Outcome
            raise RuntimeError(f"stage at t={t_start}: status {res['status'][bad[0]]} for robot {int(bad[0])} "
                               "(1 = non-finite state, 2 = contact capacity exceeded; see include/vsr_b200.h)")
        return res, box, traj, trace


class TimeBasedDevoLocomotion(DevoLocomotion):
    """tasks/devolocomotion/TimeBasedDevoLocomotion.java: development at the times of a schedule."""

    def __init__(self, development_schedule: Sequence[float], max_t: float, ground_profile, settings: Settings,
                 initial_placement: Optional[float] = None):
        super().__init__(max_t, ground_profile, settings, initial_placement)
        self.development_schedule = [float(t) for t in development_schedule]

    @staticmethod
    def fixed_interval(interval: float, max_t: float, ground_profile, settings: Settings) -> "TimeBasedDevoLocomotion":
        n = int(math.floor(max_t / interval))
        sched, d = [], interval
        for _ in range(n):  # DoubleStream.iterate(interval, d -> d + interval).limit(floor(maxT / interval)).sorted()
            sched.append(d)
            d = d + interval
        return TimeBasedDevoLocomotion(sorted(sched), max_t, ground_profile, settings)

    @staticmethod
    def uniformly_distributed(n_stages: int, max_t: float, ground_profile, settings: Settings) -> "TimeBasedDevoLocomotion":
        return TimeBasedDevoLocomotion.fixed_interval(max_t / n_stages, max_t, ground_profile, settings)

    def apply(self, solutions, device: int = 0, precision: int = A.VSR_PRECISION_F32):
        """``DevoOutcome apply(UnaryOperator<Robot>)`` (:82-137) for one solution or a list of them (all the
        solutions' robots of a stage run in one launch)."""
        single = callable(solutions)
        sols: List[Callable] = [solutions] if single else list(solutions)
        dt = self.settings.get_step_frequency()
        robots = [s(None) for s in sols]
        places = [self.initial_placement] * len(sols)
        outs = [DevoOutcome() for _ in sols]
        schedule = list(self.development_schedule)
        t = 0.0
        while t < self.max_t:
            stage_final = schedule.pop(0) if schedule else self.max_t
            # the stage ends with the first update that brings t to stageFinalT or beyond, or at maxT
            t_end = min(stage_final, self.max_t)
            ticks = tick_times(t_end, dt, t)
            if len(ticks) == 0:
                continue
            res, box, _, _ = self._stage(robots, places, t, t_end, device, precision)
            t = float(ticks[-1])
            for i in range(len(sols)):
                outs[i].add_devo_stage_outcome(DevoStageOutcome(robots[i], Outcome(res[i])))
            if t < self.max_t:
                places = [float(box[i, 0]) for i in range(len(sols))]   # robot.boundingBox().min().x()
                robots = [s(r) for s, r in zip(sols, robots)]
        return outs[0] if single else outs


class DistanceBasedDevoLocomotion(DevoLocomotion):
    """tasks/devolocomotion/DistanceBasedDevoLocomotion.java: development whenever the robot has advanced by
    ``stage_min_distance``; the task ends when a stage lasts longer than ``stage_max_t``.  A stage's end
    depends on the motion, so every solution runs its own stages (with the per-tick dump)."""

    def __init__(self, stage_min_distance: float, stage_max_t: float, max_t: float, ground_profile, settings: Settings,
                 initial_placement: Optional[float] = None):
        super().__init__(max_t, ground_profile, settings, initial_placement)
        self.stage_min_distance, self.stage_max_t = float(stage_min_distance), float(stage_max_t)

    def apply(self, solution: Callable, device: int = 0, precision: int = A.VSR_PRECISION_F32) -> DevoOutcome:
        from .behavior import build_observations
        dt = self.settings.get_step_frequency()
        robot = solution(None)
        place = self.initial_placement
        out = DevoOutcome()
        t = 0.0
        stage_x = place                       # :109 robot.boundingBox().min().x() of the first robot
        while t < self.max_t:
            stage_t = t
            # run until the stage's time limit (one update past it ends the task) or maxT
            t_end = min(self.max_t, stage_t + self.stage_max_t + dt)
            ticks = tick_times(t_end, dt, t)
            res, box, traj, trace = self._stage([robot], [place], t, t_end, device, precision, trajectories=True)
            mat = next(v for v in robot.get_voxels().values() if v is not None).material
            obs = build_observations(robot.get_voxels(), traj[0], trace[0], ticks, self.ground_profile, mat.side_length,
                                     mat.mass_side_length_ratio * mat.side_length)
            keys = list(obs)
            minx = [min(p.bounding_box().min[0] for p in o.voxel_polies.values() if p is not None) for o in obs.values()]
            cut, develop = len(keys), False
            for k, tk in enumerate(keys):
                if tk - stage_t > self.stage_max_t:   # :121-123: the task ends here
                    cut = k + 1
                    break
                if minx[k] - stage_x > self.stage_min_distance:
                    cut, develop = k + 1, True
                    break
            stage_obs = {tk: obs[tk] for tk in keys[:cut]}
            t = keys[cut - 1]
            if develop:
                out.add_devo_stage_outcome(DevoStageOutcome(robot, Outcome(observations=stage_obs)))
                place = minx[cut - 1]
                robot = solution(robot)
                # :145 stageX = robot.center().x() of the developed robot, placed with its bounding box's min x at
                # `place`: Robot.center = the mean of the voxel centres (Robot.java:123-125)
                cells = [x for (x, _), v in robot.get_voxels() if v is not None]
                side = next(v for v in robot.get_voxels().values() if v is not None).material.side_length
                stage_x = place + side * (float(np.mean(cells)) + 0.5 - min(cells))
                continue
            if stage_obs:   # :147-153: what is left when the loop ends
                out.add_devo_stage_outcome(DevoStageOutcome(robot, Outcome(observations=stage_obs)))
            break
        return out
