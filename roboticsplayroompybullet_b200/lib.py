"""ctypes binding of libprb_b200.so (include/prb.h).  The library is the product's only compute
path: if it is missing or no CUDA device is present, loading/creation FAILS LOUDLY -- there is no
CPU or PyTorch fallback."""
import ctypes
import os

from .model import PrbModelStruct

_HERE = os.path.dirname(os.path.abspath(__file__))
# PRB_LIB selects another build of the same sources (A/B experiments with compile-time knobs, tools/build_variants.sh)
LIB_PATH = os.environ.get('PRB_LIB') or os.path.join(_HERE, 'libprb_b200.so')
_LIB = None


class PrbConfig(ctypes.Structure):
    _fields_ = [('num_envs', ctypes.c_int32), ('env_offset', ctypes.c_int32), ('device', ctypes.c_int32),
                ('reserved', ctypes.c_int32), ('seed', ctypes.c_uint64)]


_PTRS = ['state', 'out_base', 'obs_quat', 'achieved_goal', 'desired_goal', 'controllable_achieved_goal',
         'full_positional_state', 'joints', 'velocity', 'observation', 'gripper_proprioception', 'reward',
         'is_success', 'target_poses']
_INTS = ['num_envs', 'state_dim', 'state_stride', 'obs_dim', 'goal_dim', 'fps_dim', 'observation_dim', 'n_ik']


class PrbBuffers(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_void_p) for n in _PTRS] + [('out_floats', ctypes.c_int64)] +
                [(n, ctypes.c_int32) for n in _INTS])


# every symbol include/prb.h declares (a CPU test checks the built library exports them all)
SYMBOLS = ['prb_create', 'prb_destroy', 'prb_reset', 'prb_reset_rounds', 'prb_reset_to', 'prb_set_goal', 'prb_step', 'prb_observe', 'prb_substeps',
           'prb_get_buffers', 'prb_compute_reward', 'prb_get_state', 'prb_set_state', 'prb_step_host',
           'prb_enable_kernel_timing', 'prb_last_kernel_ms', 'prb_last_tier_ms', 'prb_launch_count', 'prb_overflow_count', 'prb_debug_usage', 'prb_kernel_info', 'prb_last_error', 'prb_version']


class PrbError(RuntimeError):
    pass


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise PrbError('%s not built: run `python -c "import __graft_entry__ as g; g.build()"` '
                       '(nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    L.prb_create.argtypes = [ctypes.POINTER(PrbModelStruct), ctypes.POINTER(PrbConfig), ctypes.POINTER(vp)]
    L.prb_destroy.argtypes = [vp]
    L.prb_reset.argtypes = [vp, vp, vp]
    L.prb_reset_rounds.argtypes = [vp]
    L.prb_reset_to.argtypes = [vp, vp, vp, i32, vp]
    L.prb_set_goal.argtypes = [vp, vp, vp, vp]
    L.prb_step.argtypes = [vp, vp, vp]
    L.prb_observe.argtypes = [vp, vp]
    L.prb_substeps.argtypes = [vp, i32, vp]
    L.prb_get_buffers.argtypes = [vp, ctypes.POINTER(PrbBuffers)]
    L.prb_compute_reward.argtypes = [vp, vp, vp, i64, vp, vp]
    L.prb_get_state.argtypes = [vp, vp]
    L.prb_set_state.argtypes = [vp, vp]
    L.prb_step_host.argtypes = [vp, vp, vp, vp]
    L.prb_enable_kernel_timing.argtypes = [vp, i32]
    L.prb_last_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    L.prb_last_tier_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    L.prb_launch_count.argtypes = [vp]
    L.prb_launch_count.restype = i64
    L.prb_overflow_count.argtypes = [vp]
    L.prb_debug_usage.argtypes = [vp, vp]
    L.prb_overflow_count.restype = i64
    L.prb_kernel_info.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.prb_last_error.argtypes = [vp]
    L.prb_last_error.restype = ctypes.c_char_p
    L.prb_version.restype = ctypes.c_char_p
    _LIB = L
    return L


def check(L, handle, rc):
    if rc != 0:
        raise PrbError('libprb_b200: status %d: %s' % (rc, L.prb_last_error(handle).decode()))
