"""ctypes binding of libvmgym.so — the C ABI declared in include/vmgym.h.

The library is plain CUDA C++ (no torch types in any signature); PyTorch is only used by the callers for
device memory and streams.  `build()` compiles it in-tree with nvcc for sm_100a; there is NO fallback: if the
library is missing or fails to load, importing the env raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
REPO_ROOT = os.path.dirname(PKG_ROOT)
CSRC = os.path.join(PKG_ROOT, "csrc")
LIB_PATH = os.path.join(_HERE, "libvmgym.so")
SOURCES = ["vmgym_env.cu", "vmgym_policy.cu", "vmgym_gemm.cu", "vmgym_optim.cu", "vmgym_train.cu"]
# -fmad=false: the env kernels reproduce numpy's fp64/fp32 arithmetic exactly, so no FMA contraction anywhere
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]

OK, EINVAL, ECUDA, EARCH, EUNSUPPORTED = 0, -1, -2, -3, -4
REWARD_IDS = {"wr": 1, "ut": 2, "kl": 3}
AGENT_NONE, AGENT_FIRSTFIT, AGENT_BESTFIT = 0, 1, 2
TIE_IDS = {"stable": 0, "numpy_introsort": 1}
U8, I16, I64 = 1, 2, 3
TRACE_PRESAMPLED, TRACE_PHILOX = 0, 1


class Config(C.Structure):
    _fields_ = [("pms", C.c_int32), ("vms", C.c_int32), ("allow_null_action", C.c_int32),
                ("reward_function", C.c_int32), ("cap_target_util", C.c_int32), ("step_limit", C.c_int32),
                ("beta", C.c_double)]


class Layout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "record_bytes", "pms_padded", "vms_padded", "place_bytes", "off_cpu", "off_memory", "off_remaining",
        "off_placement", "off_cpu_code", "off_mem_code", "off_scalars", "obs_dim", "action_dim",
        "smem_bytes_per_env", "off_capacity")]


class Trace(C.Structure):
    _fields_ = [("mode", C.c_int32), ("reserved", C.c_int32),
                ("d_arrivals", C.c_void_p), ("arrivals_len", C.c_int64),
                ("d_admissions", C.c_void_p), ("admissions_len", C.c_int64),
                ("d_arrival_cdf", C.c_void_p), ("arrival_cdf_len", C.c_int32), ("arrival_kmin", C.c_int32),
                ("d_service_cdf", C.c_void_p), ("service_cdf_len", C.c_int32), ("service_kmin", C.c_int32),
                ("size_lo_code", C.c_int32), ("size_hi_code", C.c_int32), ("d_service_bracket", C.c_void_p)]


class Outputs(C.Structure):
    _fields_ = [("d_obs", C.c_void_p), ("d_reward", C.c_void_p), ("d_terminated", C.c_void_p),
                ("d_valid", C.c_void_p), ("d_action", C.c_void_p), ("d_stats", C.c_void_p),
                ("d_vm_slots", C.c_void_p), ("d_vm_hist", C.c_void_p), ("d_vm_totals", C.c_void_p), ("d_obs_mirror", C.c_void_p),
                ("obs_persistent", C.c_int32), ("reserved0", C.c_int32),
                ("d_next_action", C.c_void_p), ("next_agent", C.c_int32), ("next_tiebreak", C.c_int32)]


VMSTAT_BINS = 1024
STATS = 16          # doubles per env in vmgym_outputs.d_stats (VMGYM_STATS)


# offsets inside struct vmgym_env_scalars (include/vmgym.h)
SCALARS_I32 = ["timestep", "total_requests", "served_requests", "dropped_requests", "suspend_actions", "place_actions",
               "arrival_pos", "admission_pos", "status", "slot_counts"]   # slot_counts = n_waiting | n_empty << 16
SCALARS_BYTES = 80

EXPORTS = ["vmgym_last_error", "vmgym_abi_version", "vmgym_get_layout", "vmgym_reset", "vmgym_step",
           "vmgym_agent_step", "vmgym_agent_step_rotation", "vmgym_agent_act", "vmgym_observe", "vmgym_invalid_action_mask", "vmgym_set_tuning",
           "vmgym_policy_heads", "vmgym_policy_heads_backward", "vmgym_gae", "vmgym_drlvmp_choice", "vmgym_drlvmp_iter", "vmgym_linear_bf16", "vmgym_policy_fused", "vmgym_segtree_update",
           "vmgym_segtree_retrieve", "vmgym_vmstats_finalize", "vmgym_per_sample", "vmgym_c51_project", "vmgym_adamw_step", "vmgym_tc_gemm", "vmgym_cast_pad_bf16", "vmgym_cast_split_bf16",
           "vmgym_value_head", "vmgym_value_head_backward", "vmgym_ppo_loss", "vmgym_policy_fused_grad", "vmgym_policy_fused_eval", "vmgym_policy_fused_rows", "vmgym_obs_mirror_update"]


class VmgymError(RuntimeError):
    pass


HEADERS = ["vmgym_device.cuh", "vmgym_sort.cuh", "vmgym_env_kernels.cuh", "vmgym_sample.cuh", "vmgym_tc.cuh"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into vmgym/libvmgym.so (sm_100a, -lineinfo).  nvcc cross-compiles without a GPU.  Each source is
    compiled to its own object (in parallel, only when it or a header is newer), then linked."""
    from concurrent.futures import ThreadPoolExecutor
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(REPO_ROOT, "include", "vmgym.h")]
    obj_dir = os.path.join(CSRC, "_build")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + os.environ.get("VMGYM_NVCC_EXTRA", "").split()
    newest_hdr = max(os.path.getmtime(h) for h in hdrs)
    jobs, objs = [], []
    for src_name in SOURCES:
        src = os.path.join(CSRC, src_name)
        obj = os.path.join(obj_dir, src_name[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_hdr):
            cmd = [nvcc] + flags + (["-Xptxas=-v"] if verbose else []) + ["-I", os.path.join(REPO_ROOT, "include"), "-c", src, "-o", obj]
            jobs.append(cmd)
    if not jobs and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(o) for o in objs):
        return LIB_PATH
    if jobs:
        with ThreadPoolExecutor(len(jobs)) as ex:
            list(ex.map(subprocess.check_call, jobs))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH] + objs)
    return LIB_PATH


_lib = None


def lib():
    """Load libvmgym.so (building it first if the sources are newer and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    try:
        build()
    except (OSError, subprocess.CalledProcessError) as e:
        if not os.path.exists(LIB_PATH):
            raise VmgymError(f"libvmgym.so is not built and nvcc failed ({e}); there is no CPU fallback") from e
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.vmgym_last_error.restype = C.c_char_p
    L.vmgym_abi_version.restype = i32
    L.vmgym_get_layout.argtypes = [C.POINTER(Config), C.POINTER(Layout)]
    L.vmgym_reset.argtypes = [C.POINTER(Config), vp, i64, vp, vp, i32, vp, vp]
    L.vmgym_step.argtypes = [C.POINTER(Config), vp, i64, C.POINTER(Trace), vp, i32, C.POINTER(Outputs), vp]
    L.vmgym_agent_step.argtypes = [C.POINTER(Config), vp, i64, C.POINTER(Trace), i32, i32, i32, C.POINTER(Outputs), vp]
    L.vmgym_agent_step_rotation.argtypes = [C.POINTER(Config), vp, i64, i32, i32, i32, C.POINTER(Trace), i32, i32, i32,
                                            C.POINTER(Outputs), vp]
    L.vmgym_agent_act.argtypes = [C.POINTER(Config), i32, i32, vp, i64, vp, i32, vp]
    L.vmgym_observe.argtypes = [C.POINTER(Config), vp, i64, vp, vp]
    L.vmgym_obs_mirror_update.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, i64, vp]
    L.vmgym_invalid_action_mask.argtypes = [C.POINTER(Config), vp, i64, vp, vp]
    L.vmgym_per_sample.restype = i32
    L.vmgym_per_sample.argtypes = [vp, vp, i64, i64, i32, vp, C.c_double, vp, vp, vp]
    L.vmgym_c51_project.restype = i32
    L.vmgym_c51_project.argtypes = [vp, vp, vp, vp, C.c_float, C.c_float, C.c_float, i32, i64, vp, vp]
    L.vmgym_vmstats_finalize.restype = i32
    L.vmgym_vmstats_finalize.argtypes = [C.POINTER(Config), vp, i64, vp, vp, vp, vp, vp, vp]
    L.vmgym_set_tuning.argtypes = [i32, i32]
    f32, u64 = C.c_float, C.c_uint64
    L.vmgym_policy_heads.argtypes = [C.POINTER(Config), vp, vp, i32, vp, i64, vp, i32, f32, u64, u64, vp, vp, vp, vp, vp]
    L.vmgym_policy_heads_backward.argtypes = [C.POINTER(Config), vp, i32, vp, i64, vp, i32, vp, vp, vp, vp]
    L.vmgym_gae.argtypes = [vp, vp, vp, vp, C.c_int32, i64, f32, f32, vp, vp, vp]
    L.vmgym_drlvmp_choice.argtypes = [C.POINTER(Config), vp, vp, vp, i64, vp, vp]
    L.vmgym_drlvmp_iter.argtypes = [C.POINTER(Config), i32, i32, i32, vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, i64, vp]
    L.vmgym_linear_bf16.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, vp]
    L.vmgym_segtree_update.argtypes = [vp, vp, i64, vp, vp, C.c_int32, vp]
    L.vmgym_segtree_retrieve.argtypes = [vp, i64, vp, C.c_int32, vp, vp]
    L.vmgym_policy_fused.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, u64, u64, vp, vp, vp, vp]
    L.vmgym_adamw_step.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, f32, f32, vp, vp, vp, vp, vp]
    L.vmgym_tc_gemm.argtypes = [vp, i32, i64, vp, i32, i64, i64, i64, i64, vp, i32, vp, i64, vp, i64, i32, vp, i64, vp, vp]
    L.vmgym_cast_pad_bf16.argtypes = [vp, i64, i64, i64, vp, i64, vp]
    L.vmgym_cast_split_bf16.argtypes = [vp, i64, i64, i64, vp, i64, i32, vp]
    L.vmgym_value_head.argtypes = [vp, i64, i32, vp, vp, vp, vp]
    L.vmgym_value_head_backward.argtypes = [vp, i64, i32, vp, vp, vp, vp, vp, vp]
    L.vmgym_ppo_loss.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, f32, f32, f32, i32, f32, vp, vp, vp, vp]
    L.vmgym_policy_fused_grad.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, vp, f32, vp, vp, vp, vp, i64, vp]
    L.vmgym_policy_fused_rows.argtypes = [i64, i64]
    L.vmgym_policy_fused_eval.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, vp, vp, vp, vp, vp]
    for name in EXPORTS:
        getattr(L, name)
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        raise VmgymError(f"{what} failed ({rc}): {lib().vmgym_last_error().decode()}")
