"""Per-env-substep record statistics on the GPU (O.dbg: contacts, stream q, joint rows) and a dump of the
heaviest envs' states for offline analysis in the CPU emulator (gpurun_out/heavy_states.npz)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from roboticsplayroompybullet_b200.envs import make
import bench
N = 65536
env = make('UR5PlayAbsRPY1Obj-v0', num_envs=N)
obs0 = env.reset_device(); torch.cuda.synchronize()
T = 120
acts = torch.as_tensor(bench.synth_actions(np.random.default_rng(1234), N, T, 'UR5PlayAbsRPY1Obj-v0', block_xyz=obs0['achieved_goal'][:, :3].cpu().numpy(), ee_xyz=obs0['obs_quat'][:, :3].cpu().numpy())).cuda()
agg = []
for s in range(T):
    st = env.get_state() if s == T - 1 else None
    env.step_device(acts[s])
    u = env.debug_usage()
    if s >= 100: agg.append(u)
    if s == T - 1:
        worst = np.argsort(-u[:, 2])[:64]
        np.savez('gpurun_out/heavy_states.npz', state=st[worst], action=acts[s].cpu().numpy()[worst], usage=u[worst])
u = np.concatenate(agg)
ovf = u[:, 3] >> 8
u[:, 3] &= 0xff
print('capacity flags per env-substep: pairs>64 %.5f  contacts>64 %.5f  slide-slide %.5f' % tuple(((ovf >> b) & 1).mean() for b in range(3)))
for i, name in [(1, 'contacts'), (2, 'stream q'), (3, 'joint rows')]:
    x = u[:, i]
    print(name, 'mean %.1f' % x.mean(), 'pct50/90/99/99.9/max', [int(np.percentile(x, p)) for p in (50, 90, 99, 99.9, 100)])
print('stream q > 200: %.3f  > 400: %.3f  > 600: %.3f  > 900: %.3f' % tuple((u[:, 2] > x).mean() for x in (200, 400, 600, 900)))
cls, reg0 = u[:, 0] & 0x7f, u[:, 0] >> 8
print('arm-island class histogram (0 = joint-row kernel):', np.bincount(cls, minlength=5) / len(cls))
h = reg0[cls > 0]
print('region-0 q of heavy envs: pct10/50/90/99/99.9/max', [int(np.percentile(h, p)) for p in (10, 50, 90, 99, 99.9, 100)])
print('heavy envs by region-0 q <=104/144/216/320/416/864/more:', [float(((h > a) & (h <= b)).mean()) for a, b in
      ((0, 104), (104, 144), (144, 216), (216, 320), (320, 416), (416, 864), (864, 100000))])
print('overflow', env.overflow_count())
# a random sample of heavy envs (state + action of the last step) for offline work in the CPU emulator
last = agg[-1]
hv = np.nonzero((last[:, 0] & 0x7f) > 0)[0]
sel = np.random.default_rng(0).choice(hv, min(512, len(hv)), replace=False)
np.savez_compressed('gpurun_out/heavy_sample.npz', state=st[sel], action=acts[T - 1].cpu().numpy()[sel], usage=last[sel])
