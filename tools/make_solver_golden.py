"""Regression fixture for the solver kernels: heavy (arm island in contact) states sampled on the GPU from the scripted
bench workload (tools/exp_usage.py -> gpurun_out/heavy_sample.npz), stepped once by the kernels of the checkout this
script runs in, under the CPU SIMT emulator.  Written from the round-1 tree (four-lanes-per-env arm-island solver) with
round 2's tie rule of the collision code patched in (box_box TIE, manifold-reduction ties: prb_kernels.cuh), so that both
solvers see the same contacts;
tests/test_cpu_emu_parity.py::test_emu_solver_regression_bit_identical then holds every later solver to those bits.

    python tools/make_solver_golden.py <repo root> <heavy_sample.npz> <out.npz>
"""
import sys

R = sys.argv[1]
sys.path.insert(0, R)
sys.path.insert(0, R + '/tests')
import numpy as np  # noqa: E402
from roboticsplayroompybullet_b200.model import load_model  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from emu_lib import EmuSim  # noqa: E402

m = load_model('UR5PlayAbsRPY1Obj-v0')
d = np.load(sys.argv[2])
if 'usage' in d.files:
    n = 48
    # spread over the island sizes: sort by the q count of the env's record stream and take every k-th
    order = np.argsort(d['usage'][:, 2])
    sel = order[np.linspace(0, len(order) - 1, n).astype(int)]
else:                      # an earlier fixture: keep its input states, regenerate the outputs (after a tie-rule change)
    n = len(d['state'])
    sel = np.arange(n)
sim = EmuSim(m, n, seed=1)
sd = Oracle(m).state_dim
sim.state[:, :sd] = d['state'][sel]
out = sim.step(d['action'][sel])
np.savez_compressed(sys.argv[3], state=d['state'][sel], action=d['action'][sel], state_after=sim.state[:, :sd].copy(),
                    obs_quat=out['obs_quat'], reward=out['reward'])
print('wrote', sys.argv[3], 'envs', n)
