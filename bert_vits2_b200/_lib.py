"""Build + ctypes binding of libbv2.so (C ABI in include/bv2.h).  No CPU fallback: importing works without
a GPU (so CPU-only hosts can inspect the ABI), but every compute entry point needs an sm_100 device."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libbv2.so")
TUNING_LIB_PATH = os.path.join(HERE, "libbv2_tuning.so")  # development build (-DBV2_TUNING: the BV2_* environment knobs of the probes are live)
SOURCES = [os.path.join(HERE, "csrc", "engine.cu")]
HEADERS = sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith(".cuh")) + [
    os.path.join(ROOT, "include", "bv2.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]

MAX_UPS, MAX_RK, MAX_DIL = 8, 4, 4


class Bv2Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_vocab", "num_tones", "num_languages", "bert_dim", "inter_channels", "hidden_channels", "filter_channels",
        "n_heads", "n_layers", "kernel_size", "window_size", "gin_channels", "n_speakers", "n_flow_layer",
        "n_layers_trans_flow", "use_transformer_flow", "flow_kernel_size", "wn_layers", "upsample_initial_channel", "n_ups")] + [
        ("upsample_rates", C.c_int32 * MAX_UPS), ("upsample_kernel_sizes", C.c_int32 * MAX_UPS),
        ("n_resblock_kernels", C.c_int32), ("n_dilations", C.c_int32),
        ("resblock_kernel_sizes", C.c_int32 * MAX_RK), ("resblock_dilation_sizes", (C.c_int32 * MAX_DIL) * MAX_RK),
        ("sdp_filter", C.c_int32), ("sdp_kernel", C.c_int32), ("sdp_n_flows", C.c_int32), ("sdp_dds_layers", C.c_int32),
        ("sdp_num_bins", C.c_int32), ("sdp_tail_bound", C.c_float), ("dp_filter", C.c_int32), ("dp_kernel", C.c_int32),
        ("cond_layer_idx", C.c_int32), ("generator_precision", C.c_int32), ("n_flows", C.c_int32)]


#: every symbol include/bv2.h declares -> (restype, argtypes)
P, I64P, F32P = C.c_void_p, C.c_void_p, C.c_void_p
SYMBOLS = {
    "bv2_version": (C.c_char_p, []),
    "bv2_create": (C.c_int, [C.POINTER(P), C.POINTER(Bv2Config), C.c_int]),
    "bv2_set_weight": (C.c_int, [P, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_int]),
    "bv2_finalize": (C.c_int, [P]),
    "bv2_save_packed": (C.c_int, [P, C.c_char_p]),
    "bv2_load_packed": (C.c_int, [P, C.c_char_p]),
    "bv2_infer_begin": (C.c_int, [P, C.c_int, C.c_int, I64P, I64P, I64P, I64P, I64P, F32P, F32P, F32P, F32P, C.c_float,
                                  C.c_float, C.c_float, F32P, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "bv2_infer_finish": (C.c_int, [P, F32P, C.c_int64, C.c_float, C.c_int32, F32P, F32P, F32P, F32P, F32P, F32P, F32P, C.c_void_p]),
    "bv2_infer_finish_pcm16": (C.c_int, [P, F32P, C.c_int64, C.c_float, C.c_int32, C.c_void_p, F32P, F32P, F32P, F32P, F32P, F32P, C.c_void_p]),
    "bv2_wave_to_pcm16": (C.c_int, [P, C.c_int, C.c_int64, F32P, I64P, C.c_void_p, C.c_void_p]),
    "bv2_attn_path": (C.c_int, [P, F32P, C.c_void_p]),
    "bv2_reserve": (C.c_int, [P, C.c_int, C.c_int, C.c_int]),
    "bv2_text_encoder": (C.c_int, [P, C.c_int, C.c_int, I64P, I64P, I64P, I64P, I64P, F32P, F32P, F32P, F32P, F32P, F32P, C.c_void_p]),
    "bv2_duration": (C.c_int, [P, C.c_int, C.c_int, F32P, I64P, I64P, F32P, C.c_float, F32P, F32P, C.c_void_p]),
    "bv2_flow_reverse": (C.c_int, [P, C.c_int, C.c_int, F32P, I64P, I64P, F32P, C.c_void_p]),
    "bv2_generator": (C.c_int, [P, C.c_int, C.c_int, F32P, F32P, F32P, C.c_void_p]),
    "bv2_debug_read": (C.c_int64, [P, C.c_char_p, C.c_void_p, C.c_int64]),
    "bv2_set_profiling": (C.c_int, [P, C.c_int]),
    "bv2_stage_ms": (C.c_float, [P, C.c_char_p]),
    "bv2_launch_count": (C.c_int64, [P]),
    "bv2_workspace_bytes": (C.c_int64, [P]),
    "bv2_workspace_grows": (C.c_int64, [P]),
    "bv2_peer_slab_alloc": (C.c_int, [C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "bv2_peer_slab_open": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "bv2_peer_slab_close": (C.c_int, [C.c_int, C.c_void_p]),
    "bv2_peer_slab_free": (C.c_int, [C.c_int, C.c_void_p]),
    "bv2_peer_write": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "bv2_last_error": (C.c_char_p, [P]),
    "bv2_destroy": (None, [P]),
}

_lock = threading.Lock()
_lib = None


def needs_build() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS if os.path.isfile(p))


def build(force: bool = False, verbose: bool = False, tuning: bool = False) -> str:
    """Compile libbv2.so in-tree for sm_100a with nvcc (cross-compiles without a GPU).
    tuning=True builds libbv2_tuning.so instead (same sources, -DBV2_TUNING); it is only ever loaded when BV2_LIB points at it."""
    with _lock:
        out = TUNING_LIB_PATH if tuning else LIB_PATH
        if not tuning and not force and not needs_build():
            return LIB_PATH
        tmp = out + ".tmp"  # link to a temporary name, then rename: a reader never sees a half-written library
        cmd = ["nvcc"] + NVCC_FLAGS + ["-o", tmp] + SOURCES
        if tuning or os.environ.get("BV2_BUILD_TUNING"):  # development builds: BV2_* environment knobs of the probes become active (tc_conv.cuh tune_env)
            cmd.insert(1, "-DBV2_TUNING")
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        os.replace(tmp, out)
        if verbose:
            print(r.stderr)
        return out


def load():
    """dlopen libbv2.so and type every exported symbol.  Raises if the library is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("BV2_LIB") or LIB_PATH  # BV2_LIB: development A/B runs against libbv2_tuning.so
    if not os.path.isfile(path):
        raise RuntimeError(f"{path} not built; run `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "There is no CPU/PyTorch fallback for the engine.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
