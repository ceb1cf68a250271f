"""Per-round cost of a full-batch reset (PRB_TRACE_RESET=1 prints the rounds on stderr) and of masked resets of a few envs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['PRB_TRACE_RESET'] = '1'
import numpy as np, torch
from roboticsplayroompybullet_b200.envs import make
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = make('UR5PlayAbsRPY1Obj-v0', num_envs=n)
for rep in range(2):
    torch.cuda.synchronize(); t = time.time(); env.reset(); torch.cuda.synchronize()
    print('full reset %d envs: %.1f ms, rounds %d' % (n, 1e3 * (time.time() - t), env.reset_rounds()), flush=True)
mask = np.zeros(n, np.uint8); mask[::64] = 1
torch.cuda.synchronize(); t = time.time(); env.reset(mask=mask); torch.cuda.synchronize()
print('masked reset of %d envs: %.1f ms, rounds %d' % (mask.sum(), 1e3 * (time.time() - t), env.reset_rounds()), flush=True)
