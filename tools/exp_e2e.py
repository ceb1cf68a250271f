"""Where the end-to-end step time goes: two handles with the same seed and actions, one stepped through the device API,
one through the numpy API (host buffers), same steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from roboticsplayroompybullet_b200.envs import make
import bench
n = 65536
envs = [make('UR5PlayAbsRPY1Obj-v0', num_envs=n, seed=3) for _ in range(2)]
obs0 = envs[0].reset(); envs[1].reset()
T = 60
acts = bench.synth_actions(np.random.default_rng(1), n, T, 'UR5PlayAbsRPY1Obj-v0', block_xyz=obs0['achieved_goal'][:, :3], ee_xyz=obs0['obs_quat'][:, :3])
ad = torch.as_tensor(acts).cuda()
td, tn, th = [], [], []
for s in range(T):
    torch.cuda.synchronize(); t = time.perf_counter()
    envs[0].step_device(ad[s]); torch.cuda.synchronize()
    td.append(time.perf_counter() - t)
    t = time.perf_counter()
    o, r, d, i = envs[1].step(acts[s])
    tn.append(time.perf_counter() - t)
td, tn = np.array(td[20:]) * 1e3, np.array(tn[20:]) * 1e3
print('device step %.2f ms   numpy step %.2f ms   overhead %.2f ms (median %.2f)' % (td.mean(), tn.mean(), (tn - td).mean(), np.median(tn - td)))
# pieces of the host path, timed alone
e = envs[1]
t = time.perf_counter()
for s in range(20): e._h_action.numpy()[...] = acts[s]
print('action -> pinned: %.2f ms' % ((time.perf_counter() - t) / 20 * 1e3))
t = time.perf_counter()
for s in range(20): b = np.empty(e.out_floats, np.float32)
print('np.empty: %.3f ms' % ((time.perf_counter() - t) / 20 * 1e3))
pin = torch.empty(e.out_floats, dtype=torch.float32).pin_memory()
dev = torch.empty(e.out_floats, dtype=torch.float32, device='cuda')
torch.cuda.synchronize(); t = time.perf_counter()
for s in range(20): pin.copy_(dev, non_blocking=True); torch.cuda.synchronize()
print('D2H of the block into pinned: %.2f ms (%.1f GB/s)' % ((time.perf_counter() - t) / 20 * 1e3, e.out_floats * 4 / ((time.perf_counter() - t) / 20) / 1e9))
