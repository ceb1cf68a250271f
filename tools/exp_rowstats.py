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
for s in range(T):
    sim.step(acts[s])
    buf = np.ctypeslib.as_array(L.emu_sbuf(), shape=(SBQ * 32 // 2, 32, 2, 4))   # [q>>1][lane][q&1][4]
    for e in range(N):
        q = lambda i: buf[i >> 1, e, i & 1]
        hdr = q(0).view(np.int32)
        njr, nc, ns = int(hdr[0]), int(hdr[1]), int(hdr[2])
        types = {}
        for c in range(nc):
            pk = int(q(134 + c).view(np.int32)[0])
            nA, nB = (pk >> 5) & 15, (pk >> 14) & 15
            k = tuple(sorted((nA, nB)))
            types[k] = types.get(k, 0) + 1
        rows.append((njr, nc, ns, types))
njr = np.array([r[0] for r in rows]); nc = np.array([r[1] for r in rows]); ns = np.array([r[2] for r in rows])
for nm, x in (('jrows', njr), ('contacts', nc), ('spin', ns)):
    print(nm, 'mean %.1f' % x.mean(), 'pct 50/90/99/max', [int(np.percentile(x, p)) for p in (50, 90, 99, 100)])
tot = {}
for r in rows:
    for k, v in r[3].items(): tot[k] = tot.get(k, 0) + v
print('contact types (nA,nB) per env-substep:', {k: round(v / len(rows), 2) for k, v in sorted(tot.items())})
print('step time', (time.time() - t0) / T)
