import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from roboticsplayroompybullet_b200.envs import make
from roboticsplayroompybullet_b200.model import load_model, CompiledModel
import bench
N = 8192
for iters in [50, 25, 1]:
    m = load_model('UR5PlayAbsRPY1Obj-v0')
    d = dict(m.d); d.update(m.meta); d['solver_iters'] = iters
    env = make('UR5PlayAbsRPY1Obj-v0', num_envs=N, model=CompiledModel(d))
    acts = torch.as_tensor(bench.synth_actions(np.random.default_rng(0), N, 8, 'UR5PlayAbsRPY1Obj-v0')).cuda()
    env.reset_device(); torch.cuda.synchronize()
    env.enable_kernel_timing(True)
    ts = []
    for s in range(8):
        env.step_device(acts[s]); ts.append(env.last_kernel_ms()[1])
    print('solver_iters', iters, 'step kernel ms', np.round(ts[3:], 2), 'overflow', env.overflow_count())
    env.close()
