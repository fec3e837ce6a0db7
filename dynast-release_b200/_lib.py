"""ctypes binding of libdynast_b200.so (include/dynast_b200.h). Fails loudly when the library is absent."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('DYNAST_B200_LIB') or os.path.join(_HERE, 'libdynast_b200.so')


class DynastB200Error(RuntimeError):
    pass


_lib = None

_vp, _i32, _i64, _u32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32

# name -> (restype, argtypes); kept in the order of include/dynast_b200.h
PROTOTYPES = {
    'dyn_last_error': (C.c_char_p, []),
    'dyn_version': (_i32, []),
    'dyn_ctx_create': (_i32, [_i32, C.POINTER(_vp)]),
    'dyn_ctx_destroy': (_i32, [_vp]),
    'dyn_ctx_sm_count': (_i32, [_vp]),
    'dyn_tally_reads': (_i32, [_vp, _vp, _vp, _i64, _vp, _i64, _i32, _u32, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    'dyn_tally_reads_host':
        (_i32, [_vp, _vp, _vp, _i64, _vp, _i64, _i32, _u32, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    'dyn_em_pc': (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp, _vp]),
    'dyn_em_pc_host': (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp]),
    'dyn_count_records': (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _i32, _vp, _vp, _vp]),
    'dyn_complement_counts': (_i32, [_vp, _vp, _vp, _i64, _vp]),
    'dyn_barcode_first_pos': (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    'dyn_split_plan': (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    'dyn_dedup_umi': (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _u32,
                             _i32, _vp, _i32, _vp, _vp, _vp]),
    'dyn_group_key': (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    'dyn_group_sums': (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _u32, _vp, _vp, _vp, _vp]),
    'dyn_rates_pe': (_i32, [_vp, _vp, _i64, _u32, _vp, _vp, _vp]),
    'dyn_alpha': (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    'dyn_aggregate_kn': (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _u32, _u32, _u32, _i32, _i32, _i32, _vp, _vp, _vp,
                                _vp, _vp]),
    'dyn_kn_csr': (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    'dyn_kn_csr_pairs': (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'dyn_pi_fit': (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    'dyn_consensus': (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    'dyn_coverage': (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp]),
    'dyn_unique_counts': (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    'dyn_velocity': (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    'dyn_owner_partition': (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    'dyn_pack_reads': (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'dyn_unpack_reads': (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'dyn_peer_alloc': (_i32, [_vp, _i64, C.POINTER(_vp), _vp]),
    'dyn_peer_open': (_i32, [_vp, _vp, C.POINTER(_vp)]),
    'dyn_peer_close': (_i32, [_vp, _vp]),
    'dyn_peer_free': (_i32, [_vp, _vp]),
    'dyn_push_plan': (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    'dyn_push_reads': (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'dyn_merge_kn': (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    'dyn_bam_pack': (_i32, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, _i32, _i32, C.POINTER(_vp)]),
    'dyn_bam_pack_ex': (_i32, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, _i32, _i32, _u32, C.POINTER(_vp)]),
    'dyn_bam_peek': (_i32, [C.c_char_p, _i64, _vp, _i32, _vp, _vp, _vp]),
    'dyn_bam_result_free': (None, [_vp]),
    'dyn_bam_n_units': (_i64, [_vp]),
    'dyn_bam_n_records_seen': (_i64, [_vp]),
    'dyn_bam_n_refs': (_i64, [_vp]),
    'dyn_bam_field': (_vp, [_vp, _i32, C.POINTER(_i64)]),
    'dyn_bam_stream_open': (_i32, [C.c_char_p, _vp, C.POINTER(_vp)]),
    'dyn_bam_stream_next': (_i32, [_vp, C.POINTER(_vp)]),
    'dyn_bam_stream_close': (None, [_vp]),
    'dyn_bam_stream_n_refs': (_i64, [_vp]),
    'dyn_bam_stream_field': (_vp, [_vp, _i32, C.POINTER(_i64)]),
    'dyn_bam_dev_open': (_i32, [_vp, C.c_char_p, _vp, C.POINTER(_vp)]),
    'dyn_bam_dev_next': (_i32, [_vp, _vp, _vp]),
    'dyn_bam_dev_close': (None, [_vp]),
    'dyn_bam_dev_n_refs': (_i64, [_vp]),
    'dyn_bam_dev_field': (_vp, [_vp, _i32, C.POINTER(_i64)]),
    'dyn_inflate_blocks': (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    'dyn_bam_write_ex': (_i32, [C.c_char_p, _vp]),
    'dyn_bam_write': (_i32, [C.c_char_p, C.c_char_p, _i32, C.POINTER(C.c_char_p), _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                             _vp, _i32, _i32, _i32]),
    'dyn_synth_offsets': (_i32, [_vp, _vp, _i64, _vp, _vp]),
    'dyn_synth_fill': (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}


def load():
    """Load the shared library (once). Raises DynastB200Error if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DynastB200Error(
            f'{LIB_PATH} is missing: build it with `python build.py` (nvcc, sm_100a). '
            'dynast_b200 has no CPU fallback.'
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().dyn_last_error()
        raise DynastB200Error(f'dynast_b200 error {rc}: {msg.decode(errors="replace") if msg else ""}')


class Context:
    """One per GPU per process (dyn_ctx_create)."""

    _cache = {}

    def __init__(self, device: int = 0):
        lib = load()
        h = C.c_void_p()
        check(lib.dyn_ctx_create(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)
        self.lib = lib

    @classmethod
    def get(cls, device: int = 0) -> 'Context':
        if device not in cls._cache:
            cls._cache[device] = cls(device)
        return cls._cache[device]

    @property
    def sm_count(self) -> int:
        return self.lib.dyn_ctx_sm_count(self.handle)

    def close(self):
        if self.handle:
            self.lib.dyn_ctx_destroy(self.handle)
            self.handle = None
        Context._cache.pop(self.device, None)


class SynthParams(C.Structure):
    """dyn_synth_params_t"""
    _fields_ = [
        ('seed', C.c_uint64), ('read_len', C.c_int32), ('n_barcodes', C.c_int32), ('n_genes', C.c_int32),
        ('umi_len', C.c_int32), ('n_groups', C.c_int64), ('pos_span', C.c_int64), ('p_c', C.c_double),
        ('p_e', C.c_double), ('frac_n', C.c_double), ('d_bc_cdf', C.c_void_p), ('d_gene_cdf', C.c_void_p),
        ('d_gene_pi', C.c_void_p), ('d_gene_strand', C.c_void_p), ('d_gene_contig', C.c_void_p), ('lean', C.c_int32),
    ]


class BamStreamOpts(C.Structure):
    """dyn_bam_stream_opts_t"""
    _fields_ = [('barcode_tag', C.c_char_p), ('umi_tag', C.c_char_p), ('gene_tag', C.c_char_p), ('require_gene', C.c_int32),
                ('n_threads', C.c_int32), ('part', C.c_int32), ('n_parts', C.c_int32), ('chunk_bytes', C.c_int64),
                ('lean', C.c_int32), ('keys', C.c_int32), ('gene_ids', C.POINTER(C.c_char_p)), ('n_gene_ids', C.c_int64)]


class BamWriteOpts(C.Structure):
    """dyn_bam_write_opts_t"""
    _fields_ = [('header_text', C.c_char_p), ('n_ref', C.c_int32), ('ref_names', C.POINTER(C.c_char_p)), ('ref_lens', C.c_void_p),
                ('h_records', C.c_void_p), ('h_rec_off', C.c_void_p), ('n_units', C.c_int64), ('h_order', C.c_void_p),
                ('h_names', C.c_void_p), ('h_name_off', C.c_void_p), ('h_aux', C.c_void_p), ('h_aux_off', C.c_void_p),
                ('h_mapq', C.c_void_p), ('h_flag', C.c_void_p), ('level', C.c_int32), ('n_threads', C.c_int32),
                ('write_index', C.c_int32), ('reserved', C.c_int32), ('h_raw', C.c_void_p), ('h_raw_start', C.c_void_p),
                ('h_raw_len', C.c_void_p), ('h_retag', C.c_void_p), ('h_retag_off', C.c_void_p), ('gene_tag', C.c_char_p),
                ('h_md', C.c_void_p), ('h_md_off', C.c_void_p)]


class BamDevChunk(C.Structure):
    """dyn_bam_dev_chunk_t"""
    _fields_ = [('n_units', C.c_int64), ('n_records_seen', C.c_int64), ('records_bytes', C.c_int64), ('compressed_bytes', C.c_int64),
                ('inflated_bytes', C.c_int64), ('end', C.c_int32), ('reserved', C.c_int32)] + [(name, C.c_void_p) for name in (
                    'd_records', 'd_rec_off', 'd_barcode_key', 'd_umi_key', 'd_gene_idx', 'd_score', 'd_ref_id', 'd_pos', 'd_hi',
                    'd_flag', 'd_gx_assigned')]


class Annotation(C.Structure):
    """dyn_annotation_t"""
    _fields_ = [('n_ref', C.c_int32), ('n_genes', C.c_int32)] + [(name, C.c_void_p) for name in (
        'd_contig_off', 'd_gene_start', 'd_gene_end', 'd_gene_pmax', 'd_gene_strand', 'd_gene_tr_off', 'd_tr_exon_off',
        'd_tr_intron_off', 'd_exon_start', 'd_exon_end', 'd_intron_start', 'd_intron_end')]


def ptr(arr):
    """Data pointer of a numpy array / torch tensor / None."""
    if arr is None:
        return None
    if hasattr(arr, 'data_ptr'):
        return C.c_void_p(arr.data_ptr())
    return C.c_void_p(arr.ctypes.data)


def current_device_index() -> int:
    """Index of the current CUDA device (the operators run on torch's current device)."""
    import torch
    if not torch.cuda.is_available():
        raise DynastB200Error('no CUDA device: dynast_b200 has no CPU fallback')
    return torch.cuda.current_device()


def current_context() -> Context:
    return Context.get(current_device_index())


def current_torch_device():
    import torch
    return torch.device('cuda', current_device_index())
