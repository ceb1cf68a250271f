"""CPU tests of the host side: schema/header sync, the C-ABI library's exports and loud failure
without a GPU, sharding + statistics gather over gloo (world_size 2), bench helpers."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_in_sync_with_schema():
    from roboticsplayroompybullet_b200.model import c_header
    assert open(os.path.join(ROOT, 'include', 'prb_model.h')).read() == c_header()


def test_struct_layout_matches_c():
    """ctypes mirror of prb_model == what a C compiler lays out (compile a sizeof/offsetof probe)."""
    from roboticsplayroompybullet_b200.model import PrbModelStruct, SCHEMA
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "prb_model.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(prb_model), offsetof(prb_model, arm_parent), offsetof(prb_model, params));return 0;}\n'
    exe = os.path.join('/tmp', 'prb_layout_probe')
    open(exe + '.c', 'w').write(src)
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), '-o', exe, exe + '.c'])
    sz, o1, o2 = map(int, subprocess.check_output([exe]).split())
    assert sz == ctypes.sizeof(PrbModelStruct)
    assert o1 == PrbModelStruct.arm_parent.offset and o2 == PrbModelStruct.params.offset


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from roboticsplayroompybullet_b200 import lib
    L = lib.load()
    import re
    hdr = open(os.path.join(ROOT, 'include', 'prb.h')).read()
    declared = set(re.findall(r'\b(prb_[a-z_]+)\s*\(', hdr))
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s
    assert b'sm_100a' in L.prb_version()


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from roboticsplayroompybullet_b200.envs import make
    from roboticsplayroompybullet_b200.lib import PrbError
    with pytest.raises(PrbError) as e:
        make('UR5Reach-v0', num_envs=4)
    assert 'no CUDA device' in str(e.value) or 'CUDA' in str(e.value)
    with pytest.raises(NotImplementedError):
        make('pandaPlay-v0', num_envs=1)            # two-object world: not compiled (envList.py:28-33)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing in the package may import, include, link or load it."""
    import re
    pkg = os.path.join(ROOT, 'roboticsplayroompybullet_b200')
    bad = re.compile(r'(from\s+oracle|import\s+oracle|#include\s*[<"][^>"]*oracle|libprb_oracle|oracle/|orc_[a-z_]+\()')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f)).read()
                assert not bad.search(txt), (dp, f, bad.search(txt).group(0))


def test_shard_ranges():
    from roboticsplayroompybullet_b200.dist import shard_range
    for total, world in [(65536, 8), (10, 3), (7, 7), (5, 8)]:
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from roboticsplayroompybullet_b200.dist import gather_stats, max_over_ranks, shard_range
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_range(1000, rank, world)
    st = gather_stats({'env_steps': hi - lo, 'successes': rank + 1, 'reward_sum': -float(hi - lo), 'resets': 0})
    mx = max_over_ranks(0.5 + rank)
    q.put((rank, st, mx))
    dist.destroy_process_group()


def test_stats_gather_gloo_world2():
    import torch.multiprocessing as tmp
    ctx = tmp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, st, mx in res:
        assert st['env_steps'] == 1000 and st['successes'] == 3 and st['reward_sum'] == -1000 and mx == 1.5


def test_synthetic_actions_shape_and_rate():
    """Scripted teleop-shaped stream: bounded end-effector speed, gripper toggles,
    every sub-task reached; the optional Random-stream tail jumps into the +-6 clip box."""
    sys.path.insert(0, ROOT)
    import bench
    blk = np.random.default_rng(1).uniform([-0.18, 0.0, 0.0], [0.18, 0.3, 0.0], (64, 3))
    a = bench.synth_actions(np.random.default_rng(0), 64, 400, 'UR5PlayAbsRPY1Obj-v0', block_xyz=blk, ee_xyz=np.tile([0.0, 0.2, 0.25], (64, 1)))
    assert a.shape == (400, 64, 7) and a.dtype == np.float32
    step = np.linalg.norm(np.diff(a[..., :3], axis=0), axis=-1)
    assert step.max() <= 0.0151                # teleop-shaped: <= 0.015 m per 25 Hz step
    assert np.abs(np.diff(a[..., 5], axis=0)).max() <= 0.1001
    assert set(np.unique(a[..., 6])) == {-1.0, 1.0}            # gripper opens and closes
    assert np.abs(a[..., :3]).max() < 0.6
    # grasp sub-task reaches the block: some step targets a point within 2 cm above a block
    d = np.linalg.norm(a[..., :2] - blk[None, :, :2], axis=-1)
    assert ((d < 0.01) & (a[..., 2] < 0.02)).any()
    j = bench.synth_actions(np.random.default_rng(0), 64, 200, 'UR5PlayAbsRPY1Obj-v0', jump_frac=0.05)
    outside = np.abs(j[..., :3]).max(-1) > 0.6
    assert 0.01 < outside.mean() < 0.12        # ~5% jumps into the +-6 clip box


def test_bench_reference_arm_cli():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys, the
    oracle port on all host cores for a bounded sample; under torchrun only rank 0 works, the other ranks exit 0 silently."""
    import json
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1'],
                         env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ''
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--env', 'UR5Reach-v0', '--preroll', '4',
                          '--steps', '1', '--warmup', '1'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'env-steps/s' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['gpu_launches'] == 0 and line['dtype'] == 'f64'
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] == (os.cpu_count() or 1) and cb['value'] == line['value'] and 'procs x' in cb['sample']
    assert line['e2e'] == {'value': line['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
