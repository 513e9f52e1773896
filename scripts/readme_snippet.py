import importlib, math
H = importlib.import_module("2dhmsr_b200")                 # the package name starts with a digit
body = H.RobotUtils.build_sensorizing_function("uniform-ax+t-0")(H.RobotUtils.build_shape("biped-4x3"))
ctrl = H.PhaseSin(1.0, 1.0, H.Grid.create(body.w, body.h, lambda x, y: math.pi * x / body.w))
loco = H.Locomotion(20, H.Locomotion.create_terrain("flat"), H.Settings())
outcomes = loco.apply([H.Robot(ctrl, body)] * 4096)        # one kernel launch for the whole batch
print(outcomes[0].get_velocity())
o = loco.apply(H.Robot(ctrl, body), observations=True)
print(len(o.get_observations()), o.get_main_gait(0.5, 4.0, 8), list(o.get_center_y_position_spectrum(0, 5, 5).values())[:2])
