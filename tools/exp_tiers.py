import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from roboticsplayroompybullet_b200.envs import make
import bench
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
env = make('UR5PlayAbsRPY1Obj-v0', num_envs=N)
acts = torch.as_tensor(bench.synth_actions(np.random.default_rng(0), N, 12, 'UR5PlayAbsRPY1Obj-v0')).cuda()
env.reset_device(); torch.cuda.synchronize()
env.enable_kernel_timing(True)
for s in range(12):
    env.step_device(acts[s])
    a, b = env.last_tier_ms()
    u = env.debug_usage()
    print(s, 'small %.2f ms large %.2f ms' % (a, b), 'mean A', u[:, 0].mean())
