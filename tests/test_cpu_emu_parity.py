"""The product's kernel source (csrc/prb_kernels.cuh) compiled for the CPU SIMT emulator
(tests/emu) against the fp64 oracle: same tolerances as the GPU parity tests, tiny sizes.
This exercises the warp-collective code paths (shuffles, ballots, scans) without a GPU; it is a
test harness, not a product path."""
import numpy as np
import pytest

from roboticsplayroompybullet_b200.model import load_model
from oracle.oracle import Oracle
from emu_lib import EmuSim, box_box as emu_box_box
from oracle.oracle import box_box as orc_box_box


def _sync(sim, o, i=0):
    o.state[:] = o.state.astype(np.float32).astype(np.float64)
    sim.state[i, :o.state_dim] = o.state.astype(np.float32)


@pytest.mark.parametrize('env_id', ['UR5Reach-v0', 'pandaPick-v0'])
def test_emu_step_matches_oracle_arm_envs(env_id):
    m = load_model(env_id)
    sim, o = EmuSim(m, 1, seed=4), Oracle(m, seed=4)
    de, do = sim.reset(), o.reset()
    assert np.abs(de['desired_goal'][0] - do['desired_goal']).max() < 1e-6
    rng = np.random.default_rng(0)
    for _ in range(3):
        a = np.concatenate([rng.uniform(-0.15, 0.15, 3), rng.uniform(-0.3, 0.3, 3), rng.uniform(-1, 1, 1)])
        _sync(sim, o)
        de, do = sim.step(a[None]), o.step(a)
        nd = m['nd']
        assert np.abs(sim.state[0, :nd] - o.state[:nd]).max() < 2e-6                  # joint positions
        assert np.abs(de['target_poses'][0] - do['target_poses']).max() < 2e-6        # IK + clipping
        assert np.abs(de['obs_quat'][0][:3] - do['obs_quat'][:3]).max() < 1e-5
        assert de['reward'][0, 0] == do['reward'][0]


def test_emu_play_step_with_contacts():
    m = load_model('UR5PlayAbsRPY1Obj-v0')
    sim, o = EmuSim(m, 1, seed=3), Oracle(m, seed=3)
    o.reset()
    blk = o.state[60:63].copy()
    for a in [[blk[0], blk[1], 0.15, 0, 0, 0, -1], [blk[0], blk[1], 0.10, 0.1, 0, 0, 1]]:
        _sync(sim, o)
        de, do = sim.step(np.array([a])), o.step(a)
        sd = o.state_dim
        d = np.abs(sim.state[0, :sd - 1] - o.state[:sd - 1])
        assert d[:12].max() < 2e-6 and d[60:67].max() < 1e-5 and d[73:80].max() < 1e-5
        for k in ['obs_quat', 'achieved_goal', 'full_positional_state', 'observation']:
            assert np.abs(de[k][0] - do[k]).max() < 2e-5, k
        assert o.L.orc_last_contacts() >= 8            # block on table + drawer on its blockers


def test_emu_box_box_matches_oracle():
    rng = np.random.default_rng(5)
    from scipy.spatial.transform import Rotation
    n_hit = 0
    for _ in range(60):
        R1 = Rotation.random(random_state=rng.integers(1 << 30)).as_matrix()
        R2 = Rotation.random(random_state=rng.integers(1 << 30)).as_matrix()
        h1, h2 = rng.uniform(0.02, 0.1, 3), rng.uniform(0.02, 0.1, 3)
        p2 = rng.uniform(-0.08, 0.08, 3)
        a = orc_box_box([0, 0, 0], R1.reshape(-1), h1, p2, R2.reshape(-1), h2)
        b = emu_box_box([0, 0, 0], R1.reshape(-1), h1, p2, R2.reshape(-1), h2)
        assert len(a) == len(b)
        if len(a):
            n_hit += 1
            assert np.abs(a - b).max() < 5e-5
    assert n_hit > 20
