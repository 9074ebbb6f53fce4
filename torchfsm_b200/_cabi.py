"""ctypes binding of the C ABI declared in ``include/fsm_b200.h`` (libfsm_b200.so).

The product path needs the CUDA library: importing works without it (so the host logic can
be unit-tested), but any call that needs it raises ``RuntimeError`` — there is no CPU or
torch fallback. ``use_library`` exists for the CPU test-suite, which loads the host-emulator
build of the very same sources (tests/emu) to check kernel logic without a GPU.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FSM_B200_LIB: kernel-tuning hook, points at another CUDA build of the same ABI (tools/ only)
DEFAULT_LIB = os.environ.get("FSM_B200_LIB") or os.path.join(_HERE, "libfsm_b200.so")

FSM_F32, FSM_F64 = 0, 1
PROG_LINEAR, PROG_CONVECTION, PROG_KS, PROG_NS2D_VORT, PROG_NS3D = range(5)
INTEGRATOR_IDS = {"ETDRK0": 0, "ETDRK1": 1, "ETDRK2": 2, "SETDRK1": 3, "SETDRK2": 4,
                  "SETDRK3": 5, "SETDRK4": 6, "RK4": 7}


class FsmDesc(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_int32), ("dtype", ctypes.c_int32), ("ndim", ctypes.c_int32),
        ("n", ctypes.c_int32 * 3), ("batch", ctypes.c_int32), ("channels", ctypes.c_int32),
        ("program", ctypes.c_int32), ("integrator", ctypes.c_int32), ("kmax", ctypes.c_int32 * 3),
        ("ks_remove_mean", ctypes.c_int32), ("tab_channels", ctypes.c_int32), ("chunk", ctypes.c_int32),
        ("dt", ctypes.c_double), ("nl_coef", ctypes.c_double), ("ks_ext_sum", ctypes.c_double),
        ("dynamic_force", ctypes.c_int32), ("tab_complex", ctypes.c_int32),
        ("dk", ctypes.c_void_p * 3), ("dkraw", ctypes.c_void_p * 3),
        ("tab_exp", ctypes.c_void_p), ("tab_half_exp", ctypes.c_void_p), ("tab_coef", ctypes.c_void_p * 6),
        ("tab_lin", ctypes.c_void_p), ("source_hat", ctypes.c_void_p),
        ("slab_rank", ctypes.c_int32), ("slab_nranks", ctypes.c_int32),
        ("lanes", ctypes.c_int32), ("tab_batched", ctypes.c_int32),
        ("force_hat", ctypes.c_void_p), ("nl_coef_b", ctypes.c_void_p),
    ]


class FsmMapTerm(ctypes.Structure):
    """fsm_map_term of include/fsm_b200.h: one monomial-symbol term of a point-wise spectral map."""
    _fields_ = [("out_channel", ctypes.c_int32), ("in_channel", ctypes.c_int32), ("power", ctypes.c_int32 * 3),
                ("inv_laplacian", ctypes.c_int32), ("coef", ctypes.c_double)]


EXPORTS = ["fsm_plan_create", "fsm_plan_destroy", "fsm_workspace_bytes", "fsm_step", "fsm_rhs", "fsm_r2c",
           "fsm_c2r", "fsm_spectral_map", "fsm_stage_input", "fsm_stage_combine", "fsm_stage_run", "fsm_mask_state", "fsm_sym_outer", "fsm_lincomb", "fsm_half_to_full", "fsm_full_to_half", "fsm_plan_info", "fsm_plan_traffic", "fsm_stage_kinds", "fsm_slab_phase", "fsm_slab_info", "fsm_slab_peers", "fsm_ks_log", "fsm_profile_enable", "fsm_profile_read", "fsm_last_error",
           "fsm_abi_version", "fsm_backend"]

_lib = None
_lib_path = None


def _declare(lib):
    vp, sz, i32, i64p = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)
    lib.fsm_plan_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(FsmDesc)]
    lib.fsm_plan_create.restype = i32
    lib.fsm_plan_destroy.argtypes = [vp]
    lib.fsm_plan_destroy.restype = None
    lib.fsm_workspace_bytes.argtypes = [vp]
    lib.fsm_workspace_bytes.restype = sz
    lib.fsm_step.argtypes = [vp, vp, vp, sz, i32, vp]
    lib.fsm_step.restype = i32
    lib.fsm_rhs.argtypes = [vp, vp, vp, vp, sz, vp]
    lib.fsm_rhs.restype = i32
    lib.fsm_r2c.argtypes = [vp, vp, vp, vp, sz, vp]
    lib.fsm_r2c.restype = i32
    lib.fsm_c2r.argtypes = [vp, vp, vp, vp, sz, vp]
    lib.fsm_c2r.restype = i32
    lib.fsm_spectral_map.argtypes = [vp, vp, i32, vp, i32, ctypes.POINTER(FsmMapTerm), i32, i32, vp]
    lib.fsm_spectral_map.restype = i32
    lib.fsm_stage_input.argtypes = [vp, i32, i64p]
    lib.fsm_stage_input.restype = i32
    lib.fsm_stage_combine.argtypes = [vp, i32, vp, vp, vp, vp, sz, vp]
    lib.fsm_stage_combine.restype = i32
    lib.fsm_stage_run.argtypes = [vp, i32, vp, vp, vp, vp, sz, vp]
    lib.fsm_stage_run.restype = i32
    lib.fsm_mask_state.argtypes = [vp, vp, i32, vp]
    lib.fsm_mask_state.restype = i32
    lib.fsm_sym_outer.argtypes = [vp, vp, vp, i32, vp]
    lib.fsm_sym_outer.restype = i32
    lib.fsm_lincomb.argtypes = [vp, vp, vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_double), i32, ctypes.c_int64, vp]
    lib.fsm_lincomb.restype = i32
    lib.fsm_half_to_full.argtypes = [vp, vp, vp, vp]
    lib.fsm_half_to_full.restype = i32
    lib.fsm_full_to_half.argtypes = [vp, vp, vp, vp]
    lib.fsm_full_to_half.restype = i32
    lib.fsm_plan_info.argtypes = [vp, i64p, i64p, i64p, ctypes.POINTER(ctypes.c_int32)]
    lib.fsm_plan_info.restype = i32
    lib.fsm_plan_traffic.argtypes = [vp, i64p]
    lib.fsm_plan_traffic.restype = i32
    lib.fsm_stage_kinds.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), i32]
    lib.fsm_stage_kinds.restype = i32
    lib.fsm_slab_phase.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, sz, vp, vp, vp]
    lib.fsm_slab_phase.restype = i32
    lib.fsm_slab_info.argtypes = [vp, i32, i64p, i64p, ctypes.POINTER(ctypes.c_int32)]
    lib.fsm_slab_info.restype = i32
    lib.fsm_slab_peers.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_void_p), i32]
    lib.fsm_slab_peers.restype = i32
    lib.fsm_ks_log.argtypes = [vp, vp, ctypes.c_int64]
    lib.fsm_ks_log.restype = i32
    lib.fsm_profile_enable.argtypes = [vp, i32]
    lib.fsm_profile_enable.restype = i32
    lib.fsm_profile_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), i64p, i64p]
    lib.fsm_profile_read.restype = i32
    lib.fsm_last_error.argtypes = []
    lib.fsm_last_error.restype = ctypes.c_char_p
    lib.fsm_abi_version.restype = i32
    lib.fsm_backend.restype = i32
    return lib


def use_library(path):
    """Load a specific build of the C ABI (tests only: the host-emulator build)."""
    global _lib, _lib_path
    _lib = _declare(ctypes.CDLL(path))
    _lib_path = path
    return _lib


def lib():
    """The loaded C-ABI library; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(DEFAULT_LIB):
            raise RuntimeError(
                f"torchfsm_b200: the CUDA library {DEFAULT_LIB} is missing. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C torchfsm_b200/csrc`). "
                "There is no CPU/torch fallback for this path.")
        use_library(DEFAULT_LIB)
    return _lib


def is_emulator():
    return lib().fsm_backend() == 1


def library_path():
    return _lib_path


def check(code, what):
    if code != 0:
        msg = lib().fsm_last_error().decode("utf-8", "replace")
        if code == -38:  # -ENOSYS
            raise NotImplementedError(f"torchfsm_b200 {what}: {msg}")
        raise RuntimeError(f"torchfsm_b200 {what} failed ({code}): {msg}")
