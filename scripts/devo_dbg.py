import sys, importlib, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
H = importlib.import_module("2dhmsr_b200")
from oracle import oracle
import helpers
F64=1
bodies = [helpers.body_of("biped-4x3"), helpers.body_of("worm-5x2"), helpers.body_of("biped-5x3")]
def solution(seed):
    rng = np.random.default_rng(seed)
    def develop(prev):
        b = bodies[0] if prev is None else bodies[(bodies.index(prev.get_voxels()) + 1) % 3]
        ph = rng.uniform(0, 2 * np.pi, size=(b.w, b.h))
        return H.Robot(H.PhaseSin(1.0, 1.0, H.Grid.create(b.w, b.h, lambda x, y: float(ph[x, y]))), b)
    return develop
terrain = H.Locomotion.create_terrain("hilly-1-10-0")
for nsol in (1, 2):
    dev = [solution(s) for s in (1, 2)[:nsol]]
    robots = [d(None) for d in dev]; places=[11.0]*nsol; t=0.0
    for stage, t_end in enumerate([1.5, 3.0, 4.5]):
        bd = H.compile_batch(robots, t_end, terrain, None, H.Settings(), F64, nsol, t, places)
        with H.DeviceBatch(bd) as db:
            db.upload(0).run(); res, box = db.results(), db.final_bbox()
            tr = [db.trajectory(i) for i in range(nsol)]
        ores, pr = oracle.run(bd, trajectory=True)
        for i in range(nsol):
            nb = 4*bd.n_voxels[bd.robot_shape[i]] if hasattr(bd,'robot_shape') else tr[i].shape[1]
            d = np.abs(tr[i][..., :3] - pr["trajectory"][i][:, :tr[i].shape[1], :3]).max()
            print("nsol",nsol,"stage",stage,"robot",i,"place",places[i],"t",t,"dist gpu %.6f orc %.6f traj diff %.2e bbox %.6f" % (
                res[i]["last_center_x"]-res[i]["first_center_x"], ores[i]["last_center_x"]-ores[i]["first_center_x"], d, box[i,0]))
        t = float(res[0]["t_last"]); places=[float(box[i,0]) for i in range(nsol)]
        robots = [d(r) for d, r in zip(dev, robots)]
print("status of last stage", res["status"], ores["status"])
i=1
d = np.abs(tr[i][..., :3] - pr["trajectory"][i][:, :tr[i].shape[1], :3]).max(axis=(1,2))
k = int(np.argmax(d > 1e-9)); print("first bad tick", k, d[max(0,k-2):k+4])
_, prs = oracle.run(bd, stats=True)
print("oracle contacts around", prs["n_contacts"][i][max(0,k-3):k+3], "max", prs["n_contacts"][i].max())
a = tr[1][0]; o = pr["trajectory"][1][0, :a.shape[0]]
dd = np.abs(a[:, :3] - o[:, :3]); b = np.unravel_index(np.argmax(dd), dd.shape); print("worst body/comp", b)
for bb in (0, 1, 2, 3, b[0]):
    print(bb, np.round(a[bb], 5), np.round(o[bb], 5))
