"""Row statistics of the record stream (CPU emulator): joint rows, contacts by body-pair type."""
import sys, ctypes, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from roboticsplayroompybullet_b200.model import load_model
from emu_lib import EmuSim, lib
import bench
N, T = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 20
m = load_model('UR5PlayAbsRPY1Obj-v0')
sim = EmuSim(m, N, seed=1234)
t0 = time.time(); sim.reset(); print('reset', time.time() - t0)
acts = bench.synth_actions(np.random.default_rng(1234), N, T, 'UR5PlayAbsRPY1Obj-v0')
L = lib(); L.emu_sbuf.restype = ctypes.POINTER(ctypes.c_float)
SBQ = L.emu_sbuf_q()
rows = []
QST = 19
for s_ in range(T):
    sim.step(acts[s_])
    buf = np.ctypeslib.as_array(L.emu_sbuf(), shape=(SBQ, 32, 4))            # [q][lane][4]
    for e in range(N):
        h = [buf[i, e].view(np.int32) for i in range(3)]
        rows.append((int(h[0][0] & 0xff), int((h[0][0] >> 8) & 0xff), int(h[1][0] & 0xff), int(h[2][0] & 0xff),
                     int((h[1][0] >> 16) & 3), int((h[1][0] >> 18) & 3), int(h[0][3])))
r = np.array(rows)
for i, nm in enumerate(['joint rows', 'slot-0 contacts', 'slot-1 contacts', 'slot-2 contacts']):
    x = r[:, i]
    print(nm, 'mean %.1f' % x.mean(), 'pct 50/90/99/max', [int(np.percentile(x, p)) for p in (50, 90, 99, 100)])
print('free body 0 merged into the arm island: %.3f   free body 1: %.3f' % ((r[:, 4] == 0).mean(), (r[:, 5] == 0).mean()))
print('arm island needs the general solver: %.3f   region-0 q pct 50/90/99/max' % (r[:, 1] > 0).mean(), [int(np.percentile(r[:, 6], p)) for p in (50, 90, 99, 100)])
print('step time', (time.time() - t0) / T)
