"""Per-env-substep record statistics on the GPU (O.dbg: contacts, stream q, joint rows) and a dump of the
heaviest envs' states for offline analysis in the CPU emulator (gpurun_out/heavy_states.npz)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from roboticsplayroompybullet_b200.envs import make
import bench
N = 65536
env = make('UR5PlayAbsRPY1Obj-v0', num_envs=N)
acts = torch.as_tensor(bench.synth_actions(np.random.default_rng(1234), N, 24, 'UR5PlayAbsRPY1Obj-v0')).cuda()
env.reset_device(); torch.cuda.synchronize()
agg = []
for s in range(24):
    st = env.get_state()
    env.step_device(acts[s])
    u = env.debug_usage()
    if s >= 4: agg.append(u)
    if s == 23:
        worst = np.argsort(-u[:, 2])[:64]
        np.savez('gpurun_out/heavy_states.npz', state=st[worst], action=acts[s].cpu().numpy()[worst], usage=u[worst])
u = np.concatenate(agg)
for i, name in [(1, 'contacts'), (2, 'stream q'), (3, 'joint rows')]:
    x = u[:, i]
    print(name, 'mean %.1f' % x.mean(), 'pct50/90/99/99.9/max', [int(np.percentile(x, p)) for p in (50, 90, 99, 99.9, 100)])
print('heavy fraction (q > 160)', (u[:, 2] > 160).mean(), ' q > 400:', (u[:, 2] > 400).mean(), ' q > 800:', (u[:, 2] > 800).mean())
print('overflow', env.overflow_count())
