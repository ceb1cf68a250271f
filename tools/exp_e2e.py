"""Where the end-to-end step time goes: device step, D2H, host copies (numpy API)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from roboticsplayroompybullet_b200.envs import make
import bench
n = 65536
env = make('UR5PlayAbsRPY1Obj-v0', num_envs=n)
obs0 = env.reset()
acts = bench.synth_actions(np.random.default_rng(1), n, 40, 'UR5PlayAbsRPY1Obj-v0', block_xyz=obs0['achieved_goal'][:, :3], ee_xyz=obs0['obs_quat'][:, :3])
ad = torch.as_tensor(acts).cuda()
for s in range(10): env.step(acts[s])
torch.cuda.synchronize(); t = time.perf_counter()
for s in range(10, 25): o, r, d, i = env.step(acts[s])
torch.cuda.synchronize(); t_np = (time.perf_counter() - t) / 15
t = time.perf_counter()
for s in range(25, 40): env.step_device(ad[s])
torch.cuda.synchronize(); t_dev = (time.perf_counter() - t) / 15
t = time.perf_counter()
for s in range(15): c = env._host_copy()
t_copy = (time.perf_counter() - t) / 15
print('numpy step %.2f ms  device step %.2f ms  difference %.2f ms  (one single-thread copy of the block: %.2f ms, %d MB)' % (1e3 * t_np, 1e3 * t_dev, 1e3 * (t_np - t_dev), 1e3 * t_copy, env.out_floats * 4 >> 20))
