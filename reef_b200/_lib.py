"""ctypes loader for libreef_b200.so -- fails loudly when the CUDA library is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "libreef_b200.so")


class ReefError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libreef_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class NlookupOut(C.Structure):
    _fields_ = [
        ("prev_running_claim", C.c_void_p),
        ("combined_q", C.c_void_p),
        ("combined_q_cap", C.c_uint32),
        ("num_cqs", C.c_uint32),
        ("claim_r", C.c_void_p),
        ("rounds", C.c_void_p),
        ("rounds_cap", C.c_uint32),
        ("ell", C.c_uint32),
        ("sc_last_claim", C.c_void_p),
        ("next_running_claim", C.c_void_p),
    ]


class NlookupSlots(C.Structure):
    """reef_nlookup_slots (include/reef_b200.h); 2^64 - 1 = do not write"""
    _fields_ = [("claim_r", C.c_uint64), ("rounds", C.c_uint64), ("last_claim", C.c_uint64), ("next_claim", C.c_uint64)]


class CmtNldoc(C.Structure):
    """reef_cmt_nldoc (include/reef_b200.h)"""
    _fields_ = [("single_gens", C.c_void_p), ("single_gens_len", C.c_uint64), ("hyrax_gen", C.c_void_p), ("hyrax_gen_len", C.c_uint64),
                ("num_vars", C.c_uint32), ("doc_codes", C.c_void_p), ("doc_len", C.c_uint64), ("row_commitments", C.c_void_p),
                ("blinds", C.c_void_p), ("rows", C.c_uint64), ("doc_commit_hash", C.c_void_p), ("hash_salt", C.c_void_p),
                ("cap_pk", C.c_void_p), ("cap_pk_len", C.c_uint64), ("cap_vk", C.c_void_p), ("cap_vk_len", C.c_uint64),
                ("q_len", C.c_uint64), ("orig_doc_len", C.c_uint64), ("udoc_len", C.c_uint64)]


def _load():
    if not os.path.exists(lib_path):
        raise ImportError(
            f"{lib_path} not found: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "reef_b200 has no CPU fallback.")
    return C.CDLL(lib_path)


lib = _load()

_vp, _u8p = C.c_void_p, C.c_void_p
_sig = {
    "reef_abi_version": (C.c_int, []),
    "reef_launch_count": (C.c_uint64, []),
    "reef_last_error": (C.c_char_p, []),
    "reef_init": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "reef_init_prio": (C.c_int, [C.c_int, C.c_int, C.POINTER(_vp)]),
    "reef_ctx_sm_count": (C.c_uint32, [_vp]),
    "reef_shutdown": (None, [_vp]),
    "reef_sync": (C.c_int, [_vp]),
    "reef_stream": (_vp, [_vp]),
    "reef_profile_enable": (C.c_int, [_vp, C.c_int]),
    "reef_profile_read": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp]),
    "reef_logmn": (C.c_uint32, [C.c_uint64]),
    "reef_doc_transform": (C.c_int, [_vp, C.c_uint32, _vp, C.c_uint64, _vp, C.c_uint64, C.POINTER(C.c_uint64)]),
    "reef_combined_q": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, C.c_uint32, C.POINTER(C.c_uint32)]),
    "reef_io_pattern_tag": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp]),
    "reef_poseidon_hash": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint64, _vp]),
    "reef_calc_d": (C.c_int, [_vp, _vp, _vp, _vp]),
    "reef_poseidon_sponge": (C.c_int, [_vp, _vp, C.c_uint32, _vp, C.c_uint32, C.c_uint32, _vp, C.c_uint32]),
    "reef_sponge_start": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.POINTER(_vp)]),
    "reef_sponge_absorb": (C.c_int, [_vp, _vp, C.c_uint32]),
    "reef_sponge_squeeze": (C.c_int, [_vp, C.c_uint32, _vp]),
    "reef_sponge_finish": (C.c_int, [_vp]),
    "reef_merkle_tree_elems": (C.c_uint64, [C.c_uint64]),
    "reef_merkle_build": (C.c_int, [_vp, _vp, C.c_uint64, _vp, _vp, C.POINTER(C.c_uint32), _vp]),
    "reef_merkle_build_dev": (C.c_int, [_vp, _vp, C.c_uint64, _vp, _vp, C.POINTER(C.c_uint32), _vp]),
    "reef_merkle_path_wits": (C.c_int, [_vp, C.c_uint64, _vp, _vp, C.c_uint32, C.c_uint64, _vp, _vp, _vp, _vp]),
    "reef_table_upload": (C.c_int, [_vp, _vp, C.c_uint64, C.POINTER(_vp)]),
    "reef_table_upload_u32": (C.c_int, [_vp, _vp, C.c_uint64, C.POINTER(_vp)]),
    "reef_table_upload_u32_async": (C.c_int, [_vp, _vp, C.c_uint64, C.POINTER(_vp)]),
    "reef_table_hybrid_u32": (C.c_int, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, _vp, C.c_uint64, C.POINTER(_vp)]),
    "reef_table_wrap_dev": (C.c_int, [_vp, _vp, C.c_uint64, C.c_int, C.POINTER(_vp)]),
    "reef_table_download": (C.c_int, [_vp, _vp, C.c_uint64]),
    "reef_table_len": (C.c_uint64, [_vp]),
    "reef_table_free": (None, [_vp]),
    "reef_nlookup_prove": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_uint32, _vp, _vp, _vp, C.POINTER(NlookupOut)]),
    "reef_nl_shard_begin": (C.c_int, [_vp, C.c_int, _vp, C.c_uint32, C.c_uint32, _vp, _vp, C.c_uint32, _vp, _vp, _vp, C.POINTER(NlookupOut), C.POINTER(_vp)]),
    "reef_nl_shard_round_local": (C.c_int, [_vp, _vp]),
    "reef_nl_shard_round_finish": (C.c_int, [_vp, _vp]),
    "reef_nl_shard_export": (C.c_int, [_vp, _vp]),
    "reef_nl_shard_round_p2p": (C.c_int, [_vp]),
    "reef_nl_shard_finish_p2p": (C.c_int, [_vp, C.POINTER(NlookupOut)]),
    "reef_nl_shard_finish": (C.c_int, [_vp, _vp, C.POINTER(NlookupOut)]),
    "reef_nl_shard_free": (None, [_vp]),
    "reef_gen_eq_table": (C.c_int, [_vp, _vp, _vp, C.c_uint32, _vp, C.c_uint32, _vp]),
    "reef_linear_mle_product": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_uint32, _vp, _vp]),
    "reef_verifier_mle_eval": (C.c_int, [_vp, _vp, _vp, C.c_uint32, _vp]),
    "reef_prover_mle_partial_eval": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_int32, _vp, _vp]),
    "reef_hyrax_lz": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64, _vp, _vp]),
    "reef_bases_register": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64, C.c_uint32, C.POINTER(_vp)]),
    "reef_bases_free": (None, [_vp]),
    "reef_bases_windows": (C.c_uint32, [_vp]),
    "reef_bases_window_bits": (C.c_uint32, [_vp]),
    "reef_msm": (C.c_int, [_vp, _vp, _vp, C.c_uint64, _vp]),
    "reef_msm_dev": (C.c_int, [_vp, _vp, _vp, C.c_uint64, _vp]),
    "reef_msm_u32": (C.c_int, [_vp, _vp, _vp, C.c_uint64, _vp]),
    "reef_msm_rows_u32": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint64, C.c_uint32, _vp, _vp]),
    "reef_msm_rows": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint64, _vp, _vp]),
    "reef_msm_rows_dev": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint64, _vp]),
    "reef_msm_partial_dev": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, _vp]),
    "reef_msm_combine": (C.c_int, [_vp, C.c_int, _vp, C.c_uint32, _vp]),
    "reef_msm_sharded_dev": (C.c_int, [_vp, _vp, _vp, C.c_uint64, _vp]),
    "reef_merkle_subtree": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64, _vp, _vp]),
    "reef_merkle_top": (C.c_int, [_vp, _vp, C.c_uint32, _vp, _vp]),
    # test hooks (include/reef_b200_testing.h)
    "reef_hosttest_ec_op": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _vp]),
    "reef_mailbox_create": (C.c_int, [_vp, C.c_uint32, _vp]),
    "reef_mailbox_ptr": (_vp, [_vp]),
    "reef_mailbox_connect": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp]),
    "reef_mailbox_connect_local": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp]),
    "reef_p2p_allgather": (C.c_int, [_vp, _vp, C.c_uint32, _vp]),
    "reef_p2p_status": (C.c_int, [_vp]),
    "reef_sumcheck_begin": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_uint64, C.POINTER(_vp)]),
    "reef_sumcheck_round": (C.c_int, [_vp, _vp, _vp]),
    "reef_sumcheck_final": (C.c_int, [_vp, _vp, _vp]),
    "reef_sumcheck_free": (None, [_vp]),
    "reef_r1cs_spmv": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_uint64, C.c_uint64, _vp, _vp]),
    "reef_ipa_fold_bases": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64, _vp, _vp, _vp]),
    "reef_eq_table": (C.c_int, [_vp, C.c_int, _vp, C.c_uint32, _vp]),
    "reef_bases_cache_stats": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "reef_witness_create": (C.c_int, [_vp, C.c_uint64, C.POINTER(_vp)]),
    "reef_witness_set": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "reef_witness_set_u64": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "reef_witness_read": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp]),
    "reef_witness_dev": (_vp, [_vp]),
    "reef_witness_len": (C.c_uint64, [_vp]),
    "reef_witness_free": (None, [_vp]),
    "reef_nlookup_prove_w": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_uint32, _vp, _vp, _vp, C.POINTER(NlookupOut), _vp, _vp]),
    "reef_nova_cross_term": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_uint64, _vp]),
    "reef_vec_axpy": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_uint64, _vp]),
    "reef_ipa_begin": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_uint64, C.POINTER(_vp)]),
    "reef_ipa_begin_bases": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_uint64, C.POINTER(_vp)]),
    "reef_ipa_round": (C.c_int, [_vp, _vp, _vp]),
    "reef_ipa_fold": (C.c_int, [_vp, _vp, _vp]),
    "reef_ipa_finish": (C.c_int, [_vp, _vp, _vp, _vp]),
    "reef_ipa_free": (None, [_vp]),
    "reef_poseidon_ro": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64, C.c_uint32, _vp]),
    "reef_poseidon_ro_points": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64, C.c_uint32, _vp]),
    "reef_doc_commit_u32": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint64, C.c_uint32, _vp, _vp, _vp]),
    "reef_cmt_merkle_size": (C.c_uint64, [_vp, C.c_uint32, C.c_uint64]),
    "reef_cmt_merkle_write": (C.c_int, [_vp, _vp, _vp, C.c_uint32, _vp, C.c_uint64, C.c_uint64, _vp, C.c_uint64, C.POINTER(C.c_uint64)]),
    "reef_cmt_probe": (C.c_int, [_vp, C.c_uint64, C.POINTER(C.c_int)]),
    "reef_cmt_merkle_read": (C.c_int, [_vp, C.c_uint64, _vp, _vp, C.c_uint64, _vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                       _vp, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "reef_point_compress": (C.c_int, [_vp, _vp]),
    "reef_cmt_nldoc_size": (C.c_uint64, [_vp]),
    "reef_cmt_nldoc_write": (C.c_int, [_vp, _vp, C.c_uint64, C.POINTER(C.c_uint64)]),
    "reef_hosttest_poseidon_ro": (C.c_int, [C.c_int, _vp, C.c_uint64, _vp]),
    "reef_hosttest_poseidon_ro_constants": (C.c_int, [C.c_int, _vp, _vp]),
    "reef_hosttest_poseidon_ro_fast_ok": (C.c_int, [C.c_int]),
    "reef_hosttest_field_op": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _vp]),
    "reef_hosttest_mul_wide": (C.c_int, [_vp, _vp, _vp]),
    "reef_hosttest_poseidon_permute": (C.c_int, [_vp, _vp]),
    "reef_gputest_poseidon_permute_lp": (C.c_int, [_vp, _vp, C.c_uint32, _vp, _vp]),
}
_missing = []
for _name, (_res, _args) in _sig.items():
    try:
        _f = getattr(lib, _name)
    except AttributeError:
        _missing.append(_name)
        continue
    _f.restype = _res
    _f.argtypes = _args
if _missing:
    raise ImportError(f"libreef_b200.so is stale: missing symbols {_missing}; rebuild with `make`")


def check(rc: int):
    if rc != 0:
        raise ReefError(rc, lib.reef_last_error().decode("utf-8", "replace"))
