"""Compile the reference assets (URDFs, meshes, programmatic scenes) into the
struct-of-arrays models shipped under roboticsplayroompybullet_b200/assets/.

Run in a container where the reference checkout is available:
    python tools/compile_models.py /root/reference/roboticsPlayroomPybullet/envs
The GPU box has no /root/reference; it only ever loads the committed .npz files.
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from roboticsplayroompybullet_b200.compiler.compile import compile_env
from roboticsplayroompybullet_b200.model import asset_path, ENV_KINDS

if __name__ == '__main__':
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/roboticsPlayroomPybullet/envs'
    for env_id in ENV_KINDS:
        m = compile_env(env_id, ref)
        os.makedirs(os.path.dirname(asset_path(env_id)), exist_ok=True)
        m.save(asset_path(env_id))
        print(env_id, 'nd', m['nd'], 'n_col', m['n_col'], 'n_pair', m['n_pair'], 'n_free', m['n_free'],
              'n_slide', m['n_slide'])
