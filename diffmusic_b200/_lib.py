"""ctypes binding of include/dm_abi.h.  There is NO fallback: if the CUDA library is missing or a call fails, this raises.

The library is built in-tree by `python -m diffmusic_b200.build` (or __graft_entry__.build()).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdiffmusic_b200.so")

c_f, c_i, c_ll, c_p = C.c_float, C.c_int, C.c_longlong, C.c_void_p


class StftTables(C.Structure):
    """struct dm_stft_tables (include/dm_abi.h)."""
    _fields_ = [("window", c_p), ("tw512", c_p), ("w1024", c_p), ("mel_kstart", c_p), ("mel_klen", c_p),
                ("mel_w", c_p), ("mel_wstride", c_i), ("bin_m0", c_p), ("bin_w0", c_p), ("bin_w1", c_p),
                ("warp_image", c_p), ("warp_image_floats", c_i), ("warp_na", c_i), ("warp_nb", c_i)]


_SIGNATURES = {
    "dm_version": (c_i, []),
    "dm_last_error": (C.c_char_p, []),
    "dm_launch_count": (C.c_ulonglong, []),
    "dm_reset_launch_count": (None, []),
    "dm_set_tuning": (c_i, [c_i, c_i]),
    "dm_get_tuning": (c_i, [c_i]),
    "dm_sched_x0": (c_i, [c_p, c_p, c_p, c_ll, c_f, c_f, c_i, c_f, c_p, c_p]),
    "dm_sched_ddim_update": (c_i, [c_p, c_p, c_p, c_ll, c_f, c_f, c_f, c_f, c_p, c_p]),
    "dm_sched_dps_update": (c_i, [c_p, c_p, c_p, c_p, c_p, c_ll, c_f, c_f, c_f, c_f, c_f, c_f, c_p, c_p]),
    "dm_sched_mpgd_update": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_f, c_f, c_f, c_f, c_f, c_f, c_p, c_p]),
    "dm_sched_dsg_update": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_ll, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_p,
                                  c_p]),
    "dm_sched_diffmusic_update": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_ll, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f,
                                        c_p, c_p]),
    "dm_sched_x0_io": (c_i, [c_p, c_p, c_p, c_p, c_p, c_f, c_ll, c_f, c_f, c_i, c_f, c_p, c_i, c_p]),
    "dm_sched_ddim_update_io": (c_i, [c_p, c_p, c_p, c_ll, c_f, c_f, c_f, c_f, c_p, c_i, c_p]),
    "dm_sched_dps_update_io": (c_i, [c_p, c_p, c_p, c_f, c_p, c_p, c_ll, c_f, c_f, c_f, c_f, c_f, c_f, c_p, c_i,
                                     c_p, c_i, c_p, c_p]),
    "dm_sched_mpgd_update_io": (c_i, [c_p, c_p, c_p, c_f, c_p, c_p, c_p, c_ll, c_f, c_f, c_f, c_f, c_f, c_f, c_p,
                                      c_i, c_p, c_i, c_p, c_p]),
    "dm_sched_dsg_update_io": (c_i, [c_p, c_p, c_p, c_f, c_p, c_p, c_i, c_ll, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f,
                                     c_p, c_i, c_p, c_p, c_p]),
    "dm_sched_diffmusic_update_io": (c_i, [c_p, c_p, c_p, c_f, c_p, c_p, c_i, c_ll, c_f, c_f, c_f, c_f, c_f, c_f,
                                           c_f, c_f, c_p, c_i, c_p, c_p, c_p]),
    "dm_stft_guidance_io": (c_i, [C.POINTER(StftTables), c_i, c_i, c_i, c_p, c_i, c_ll, c_ll, c_p, c_i, c_p, c_ll, c_p,
                                  c_f, c_p, c_p, c_p, c_i, c_p]),
    "dm_stft_guidance_fir2": (c_i, [C.POINTER(StftTables), c_i, c_i, c_p, c_ll, c_ll, c_p, c_i, c_p, c_ll, c_p, c_p,
                                    c_i, c_p]),
    "dm_residual_wav_io": (c_i, [c_p, c_i, c_ll, c_ll, c_i, c_p, c_p, c_ll, c_p, c_p, c_p]),
    "dm_fold_adjoint_io": (c_i, [c_p, c_i, c_ll, c_i, c_p, c_p, c_i, c_p, c_i, c_ll, c_p, c_p]),
    "dm_resample_fwd_io": (c_i, [c_p, c_i, c_ll, c_ll, c_i, c_p, c_i, c_i, c_i, c_i, c_p, c_ll, c_p]),
    "dm_resample_fwd_fill_io": (c_i, [c_p, c_i, c_ll, c_ll, c_i, c_p, c_i, c_i, c_i, c_i, c_p, c_ll, c_p, c_ll, c_p]),
    "dm_resample_adjoint_io": (c_i, [c_p, c_i, c_ll, c_i, c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_ll, c_ll,
                                     c_p, c_p]),
    "dm_rir_correlate_io": (c_i, [c_p, c_i, c_ll, c_ll, c_i, c_p, c_i, c_p, c_p, c_p, c_ll, c_p]),
    "dm_rir_adjoint_io": (c_i, [c_p, c_i, c_ll, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_i, c_ll, c_ll, c_p, c_p]),
    "dm_randn_offset_increment": (c_ll, [c_ll]),
    "dm_randn_clips": (c_i, [c_p, c_p, c_i, c_ll, c_i, c_p, c_p]),
    "dm_stft_set_engine": (c_i, [c_i]),
    "dm_stft_num_tiles": (c_i, [c_ll, c_i, c_i]),
    "dm_stft_guidance": (c_i, [C.POINTER(StftTables), c_i, c_i, c_i, c_p, c_ll, c_ll, c_p, c_i, c_p, c_ll, c_p, c_f,
                               c_p, c_p, c_p, c_i, c_p]),
    "dm_mel_project": (c_i, [C.POINTER(StftTables), c_p, c_i, c_ll, c_i, c_p, c_p]),
    "dm_mask_apply": (c_i, [c_p, c_ll, c_ll, c_i, c_p, c_p, c_p]),
    "dm_residual_wav": (c_i, [c_p, c_ll, c_ll, c_i, c_p, c_p, c_ll, c_p, c_p, c_p]),
    "dm_fold_adjoint": (c_i, [c_p, c_i, c_ll, c_i, c_p, c_p, c_i, c_p, c_ll, c_p, c_p]),
    "dm_resample_fwd": (c_i, [c_p, c_ll, c_ll, c_i, c_p, c_i, c_i, c_i, c_i, c_p, c_ll, c_p]),
    "dm_resample_adjoint": (c_i, [c_p, c_i, c_ll, c_i, c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_p, c_ll, c_ll, c_p, c_p]),
    "dm_rir_spectrum": (c_i, [c_p, c_i, c_p, c_p, c_p, c_p]),
    "dm_rir_correlate": (c_i, [c_p, c_ll, c_ll, c_i, c_p, c_i, c_p, c_p, c_p, c_ll, c_p]),
    "dm_rir_adjoint": (c_i, [c_p, c_i, c_ll, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_ll, c_ll, c_p, c_p]),
    "dm_add_scaled": (c_i, [c_p, c_p, c_f, c_ll, c_p]),
    "dm_copy_f32": (c_i, [c_p, c_p, c_ll, c_p]),
    "dm_fad_moments": (c_i, [c_p, c_ll, c_i, c_p, c_p]),
    "dm_fad_moments_ex": (c_i, [c_p, c_ll, c_i, c_p, c_i, c_p]),
    "dm_fad_finalize": (c_i, [c_p, c_i, c_p, c_p, c_p]),
    "dm_fad_finalize_sym": (c_i, [c_p, c_i, c_p, c_p, c_p]),
    "dm_enable_peer_access": (c_i, [c_i, c_i]),
    "dm_peer_alloc": (c_i, [c_i, c_ll, c_p]),
    "dm_peer_free": (c_i, [c_i, c_p]),
    "dm_ipc_export": (c_i, [c_i, c_p, c_p]),
    "dm_ipc_open": (c_i, [c_i, c_p, c_p]),
    "dm_ipc_close": (c_i, [c_i, c_p]),
    "dm_fad_packed_doubles": (c_ll, [c_i]),
    "dm_fad_flag_words": (c_i, []),
    "dm_fad_reset_shared": (c_i, [c_p, c_i, c_p, c_i, C.c_uint, c_p]),
    "dm_fad_allreduce_peers": (c_i, [c_p, c_p, c_i, c_i, c_i, C.c_uint, c_p, c_p]),
    "dm_fad_allreduce": (c_i, [c_p, c_p, c_i, c_p, c_p]),
    "dm_fad_allreduce_push": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, C.c_uint, c_p]),
    "dm_fad_finalize_shared": (c_i, [c_p, c_i, c_p, c_i, C.c_uint, c_p, c_p, c_p]),
    "dm_fad_pack_tri": (c_i, [c_p, c_i, c_p, c_p]),
    "dm_fad_unpack_tri": (c_i, [c_p, c_i, c_p, c_p]),
    "dm_workspace_bytes": (c_ll, [c_i, c_ll, c_ll, c_i]),
    "dm_fad_gather_rows": (c_i, [c_p, c_ll, c_i, c_p, c_ll, c_p, c_p]),
    "dm_sym_eig_jacobi": (c_i, [c_p, c_i, c_i, C.c_double, c_p, c_p, c_p]),
    "dm_frechet_workspace_doubles": (c_ll, [c_i]),
    "dm_lsd_frames": (c_i, [C.POINTER(StftTables), c_p, c_ll, c_p, c_ll, c_ll, c_i, c_i, c_i, c_i, c_f, c_p, c_p]),
    "dm_mse_num_chunks": (c_ll, [c_ll]),
    "dm_mse": (c_i, [c_p, c_ll, c_p, c_ll, c_ll, c_i, c_p, c_p, c_p]),
    "dm_stft_spectrogram": (c_i, [C.POINTER(StftTables), c_p, c_ll, c_ll, c_i, c_i, c_p, c_p, c_p]),
    "dm_istft_workspace_floats": (c_ll, [c_i, c_ll, c_i]),
    "dm_istft_mel_phase": (c_i, [C.POINTER(StftTables), c_p, c_p, c_ll, c_ll, c_ll, c_p, c_ll, c_i, c_ll, c_i, c_p, c_p,
                                 c_ll, c_p]),
    "dm_frechet_distance": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, C.c_double, c_p, c_p, c_p]),
}

EXPORTS = tuple(_SIGNATURES)
TUNING_ENV = ("DM_TUNE_STREAM_KERNELS", "DM_TUNE_PDL")  # index = knob number
_lib = None


class DiffMusicB200Error(RuntimeError):
    pass


def load():
    """Load the shared library once.  Raises if it has not been built -- never falls back to a CPU/torch path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DiffMusicB200Error(
                f"{LIB_PATH} not found: build it with `python -m diffmusic_b200.build` (nvcc, sm_100a). "
                "diffmusic_b200 has no CPU or eager-torch fallback for the guided path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        for knob, env in enumerate(TUNING_ENV):  # process-wide kernel-selection knobs (include/dm_abi.h DM_TUNE_*)
            if os.environ.get(env) is not None:
                lib.dm_set_tuning(knob, int(os.environ[env]))
        _lib = lib
    return _lib


def call(name, *args):
    """Invoke an int-returning entry point; raise with dm_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise DiffMusicB200Error(f"{name} failed ({rc}): {lib.dm_last_error().decode()}")


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


IO_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def stream(device=None):
    """raw cudaStream_t of torch's current stream (the one every kernel of this library launches on).  The private fast
    getter costs ~1 us of host time against ~10 us for the Stream object round trip -- it is called once per kernel."""
    try:
        idx = torch.cuda.current_device() if device is None or device.index is None else device.index
        return torch._C._cuda_getCurrentRawStream(idx)
    except AttributeError:  # pragma: no cover - older / newer torch without the private getter
        return torch.cuda.current_stream(device).cuda_stream


def launch_count():
    return int(load().dm_launch_count())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise DiffMusicB200Error("diffmusic_b200 kernels need CUDA tensors (no CPU fallback on the guided path)")
