"""ctypes binding of liblpic_b200.so (the C-ABI declared in include/lpic_b200.h).

The library is built in-tree by ``lambdapic_b200/csrc/Makefile`` (``__graft_entry__.build()``).  There is no
CPU fallback: if the shared library is missing, or no CUDA device is present when a context is created,
the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LPIC_B200_LIB", os.path.join(HERE, "liblpic_b200.so"))  # override: A/B builds of the library

FIELD_ATTRS = ["ex", "ey", "ez", "bx", "by", "bz", "jx", "jy", "jz", "rho"]  # core/fields.py:71-75
PART_ATTRS = ["x", "y", "z", "w", "ux", "uy", "uz", "inv_gamma",
              "ex_part", "ey_part", "ez_part", "bx_part", "by_part", "bz_part", "_id"]  # core/particles.py:63-67
P_IS_DEAD = len(PART_ATTRS)
SORT_BUCKET_COUNT, SORT_BOUND_MIN, SORT_BOUND_MAX, SORT_PARTICLE_INDEX = range(4)
PUSH_WRITE_PART = 1
PUSH_SLOT_ORDER = 2

_i64, _dbl, _vp, _int, _u32, _u64 = C.c_int64, C.c_double, C.c_void_p, C.c_int, C.c_uint32, C.c_uint64

# name -> (restype, argtypes); every symbol of include/lpic_b200.h
SIGNATURES = {
    "lpic_last_error": (C.c_char_p, []),
    "lpic_device_count": (_int, []),
    "lpic_version": (C.c_char_p, []),
    "lpic_host_alloc": (_vp, [_i64]),
    "lpic_host_free": (None, [_vp]),
    "lpic_create": (_vp, [_int, _i64, _i64, _i64, _i64, _i64, _dbl, _dbl, _dbl, _int, _int]),
    "lpic_destroy": (None, [_vp]),
    "lpic_set_patch_geometry": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "lpic_sync": (_int, [_vp]),
    "lpic_field_cells": (_i64, [_vp]),
    "lpic_upload_fields": (_int, [_vp, _u32, _vp]),
    "lpic_download_fields": (_int, [_vp, _u32, _vp]),
    "lpic_upload_field_ptrs": (_int, [_vp, _int, _vp]),
    "lpic_download_field_ptrs": (_int, [_vp, _int, _vp]),
    "lpic_species_alloc": (_int, [_vp, _int, _vp, _dbl, _i64, _int]),
    "lpic_species_layout": (_int, [_vp, _int, _vp, _vp, _vp, _vp]),
    "lpic_upload_particles": (_int, [_vp, _int, _int, _vp]),
    "lpic_download_particles": (_int, [_vp, _int, _int, _vp]),
    "lpic_upload_particle_records": (_int, [_vp, _int, _u32, _vp]),
    "lpic_download_particle_records": (_int, [_vp, _int, _u32, _vp]),
    "lpic_upload_particle_ptrs": (_int, [_vp, _int, _int, _vp]),
    "lpic_download_particle_ptrs": (_int, [_vp, _int, _int, _vp]),
    "lpic_species_extend": (_int, [_vp, _int, _vp, _vp, _vp]),
    "lpic_species_set_npart": (_int, [_vp, _int, _vp]),
    "lpic_update_efield": (_int, [_vp, _dbl]),
    "lpic_update_bfield": (_int, [_vp, _dbl]),
    "lpic_pml_configure": (_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64]),
    "lpic_pml_psi_words": (_i64, [_vp]),
    "lpic_pml_upload_psi": (_int, [_vp, _vp]),
    "lpic_pml_download_psi": (_int, [_vp, _vp]),
    "lpic_laser_bfields": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _dbl]),
    "lpic_sync_guard_fields": (_int, [_vp, _u32]),
    "lpic_sync_currents": (_int, [_vp]),
    "lpic_reset_currents": (_int, [_vp]),
    "lpic_push_deposit": (_int, [_vp, _int, _dbl, _dbl, _dbl, _int]),
    "lpic_interpolate": (_int, [_vp, _int]),
    "lpic_push_momentum": (_int, [_vp, _int, _dbl, _dbl, _dbl]),
    "lpic_push_position": (_int, [_vp, _int, _dbl]),
    "lpic_deposit": (_int, [_vp, _int, _dbl, _dbl]),
    "lpic_sort": (_int, [_vp, _int, _int, _i64, _i64, _i64, _dbl, _dbl, _dbl, _vp, _vp, _vp, _vp]),
    "lpic_sort_download": (_int, [_vp, _int, _int, _vp]),
    "lpic_weighted_drift": (_int, [_vp, _int, _vp]),
    "lpic_migrate_count": (_int, [_vp, _int, _vp, _vp, _vp, _vp]),
    "lpic_migrate_fill": (_int, [_vp, _int]),
    "lpic_count_alive": (_int, [_vp, _int, _vp]),
    "lpic_kinetic_sum": (_int, [_vp, _int, _vp]),
    "lpic_field_energy_sums": (_int, [_vp, _vp]),
    "lpic_species_init_uniform": (_int, [_vp, _int, _i64, _dbl, _dbl, _u64]),
    "lpic_halo_plan": (_int, [_vp, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lpic_halo_words": (_i64, [_vp, _int, _int]),
    "lpic_halo_pack": (_int, [_vp, _int, _u32, _int, _vp]),
    "lpic_halo_unpack": (_int, [_vp, _u32, _int, _vp]),
    "lpic_particle_record_words": (_int, [_vp, _int]),
    "lpic_remote_migrate_prepare": (_int, [_vp, _int, _vp, _vp]),
    "lpic_remote_migrate_relist": (_int, [_vp, _int]),
    "lpic_remote_migrate_pack": (_int, [_vp, _int, _int, _vp, _vp]),
    "lpic_remote_migrate_unpack": (_int, [_vp, _int, _vp, _vp]),
    "lpic_event_record": (_int, [_vp, _int]),
    "lpic_event_elapsed_ms": (_int, [_vp, _int, _int, _vp]),
    "lpic_launch_count": (_i64, []),
    "lpic_fp64_peak": (_int, [_vp, _vp]),
    "lpic_zero_patches": (_int, [_vp, _i64, _vp]),
    "lpic_upload_particles_patch": (_int, [_vp, _int, _int, _i64, _vp]),
    "lpic_device_pci_bus_id": (_int, [_vp, _vp, _int]),
    "lpic_download_field_slice": (_int, [_vp, _u32, _vp, _vp]),
    "lpic_comm_unique_id": (_int, [_vp]),
    "lpic_comm_nccl_version": (_int, []),
    "lpic_comm_init": (_int, [_vp, _vp, _int, _int, _vp]),
    "lpic_comm_bytes_sent": (_i64, [_vp]),
    "lpic_comm_update": (_int, [_vp, _vp]),
    "lpic_halo_start": (_int, [_vp, _u32, _int]),
    "lpic_halo_wait": (_int, [_vp]),
    "lpic_migrate_remote_start": (_int, [_vp, _int, _int, _vp, _vp]),
    "lpic_migrate_remote_wait": (_int, [_vp, _int]),
    "lpic_stream": (_vp, [_vp]),
}

_LIB = None


class LpicError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile liblpic_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(HERE, "csrc"), "-j8"], stdout=out)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise LpicError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(status: int):
    if status != 0:
        raise LpicError(lib().lpic_last_error().decode())
