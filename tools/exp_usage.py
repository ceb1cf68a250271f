import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from roboticsplayroompybullet_b200.envs import make
import bench
N = 16384
env = make('UR5PlayAbsRPY1Obj-v0', num_envs=N)
acts = torch.as_tensor(bench.synth_actions(np.random.default_rng(0), N, 40, 'UR5PlayAbsRPY1Obj-v0')).cuda()
env.reset_device(); torch.cuda.synchronize()
agg = []
for s in range(40):
    env.step_device(acts[s])
    if s >= 10: agg.append(env.debug_usage())
u = np.concatenate(agg)
for i, name in enumerate(['A floats', 'contacts', 'pool floats', 'units']):
    x = u[:, i]
    print(name, 'mean %.0f' % x.mean(), 'pct50/80/90/95/99/max', [int(np.percentile(x, p)) for p in (50, 80, 90, 95, 99, 100)])
for cap in [1400, 1536, 1920, 2560, 3072, 4096, 5888]:
    print('A <=', cap, 'fraction of env-steps %.3f' % (u[:, 0] <= cap).mean())
for cap in [512, 768, 1024, 1408]:
    print('pool <=', cap, '%.3f' % (u[:, 2] <= cap).mean())
print('overflow', env.overflow_count())
