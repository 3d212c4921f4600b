"""ctypes view of include/pangenie_b200.h.

The structures are shared by the product library (prefix ``pg_``) and — in tests only — by the CPU
oracles, which export the same signatures under the prefixes ``pgo_`` (oracle/pg_oracle.cpp) and
``pgr_`` (oracle/ref_shim.cpp).  This module never loads an oracle itself.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PG_OK, PG_ERR_ARG, PG_ERR_CUDA, PG_ERR_FORMAT, PG_ERR_FULL, PG_ERR_IO = range(6)
PG_OP_COUNT, PG_OP_PRIME, PG_OP_UPDATE = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpangenie_b200.so")


class PgPanel(C.Structure):
    _fields_ = [
        ("n_variants", C.c_uint32),
        ("n_paths", C.c_uint32),
        ("positions", C.c_void_p),
        ("path_to_allele", C.c_void_p),
        ("coverage", C.c_void_p),
        ("kmer_offsets", C.c_void_p),
        ("kmer_counts", C.c_void_p),
        ("allele_offsets", C.c_void_p),
        ("allele_ids", C.c_void_p),
        ("allele_undefined", C.c_void_p),
        ("allele_kmer_offset", C.c_void_p),
        ("allele_kmer_mask", C.c_void_p),
        ("kmer_codes", C.c_void_p),
        ("flank_offsets", C.c_void_p),
        ("flank_codes", C.c_void_p),
    ]


class PgProbTable(C.Structure):
    _fields_ = [
        ("cov_min", C.c_uint16),
        ("cov_max", C.c_uint16),
        ("count_max", C.c_uint16),
        ("regularization", C.c_double),
        ("log_p", C.c_void_p),
    ]


class PgHmmParams(C.Structure):
    _fields_ = [
        ("recombrate", C.c_double),
        ("effective_N", C.c_double),
        ("uniform", C.c_int),
        ("normalize", C.c_int),
        ("only_paths", C.c_void_p),
        ("n_only_paths", C.c_uint32),
    ]


class PgHmmResult(C.Structure):
    _fields_ = [
        ("gl_offsets", C.c_void_p),
        ("likelihoods", C.c_void_p),
        ("is_column", C.c_void_p),
        ("genotype", C.c_void_p),
        ("quality", C.c_void_p),
        ("unique_kmers", C.c_void_p),
        ("coverage", C.c_void_p),
    ]


class PgGenotypeInput(C.Structure):
    _fields_ = [
        ("reads", C.c_void_p),
        ("reads_len", C.c_uint64),
        ("segments", C.c_void_p),
        ("segments_len", C.c_uint64),
        ("k", C.c_uint32),
        ("hash_size", C.c_uint64),
        ("regularization", C.c_double),
        ("histogram_path", C.c_char_p),
    ]


class PgVariants(C.Structure):
    _fields_ = [
        ("n_variants", C.c_uint32),
        ("n_paths", C.c_uint32),
        ("k", C.c_uint32),
        ("positions", C.c_void_p),
        ("end_positions", C.c_void_p),
        ("path_to_allele", C.c_void_p),
        ("allele_offsets", C.c_void_p),
        ("allele_undefined", C.c_void_p),
        ("seq_offsets", C.c_void_p),
        ("seq", C.c_void_p),
        ("left_offsets", C.c_void_p),
        ("left_seq", C.c_void_p),
        ("right_offsets", C.c_void_p),
        ("right_seq", C.c_void_p),
    ]


class PgTimings(C.Structure):
    _fields_ = [
        ("count_ms", C.c_double),
        ("histogram_ms", C.c_double),
        ("fill_ms", C.c_double),
        ("emission_ms", C.c_double),
        ("hmm_skeleton_ms", C.c_double),
        ("hmm_blocks_ms", C.c_double),
        ("finalize_ms", C.c_double),
        ("hmm_columns", C.c_uint64),
        ("hmm_block_launches", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("kmers_counted", C.c_uint64),
        ("text_bytes", C.c_uint64),
        ("prime_ms", C.c_double),
        ("hmm_scan_used", C.c_uint64),
        ("count_probe_ms", C.c_double),
        ("count_probe_passes", C.c_uint64),
    ]


# every symbol include/pangenie_b200.h declares (tests check the .so exports each one)
EXPORTS = [
    "pg_last_error", "pg_version", "pg_device_count", "pg_kernel_launches", "pg_count_last_probe_ms",
    "pg_count_create", "pg_count_create_from_buffers", "pg_count_new", "pg_count_feed", "pg_count_feed_device",
    "pg_count_lookup_ascii", "pg_count_lookup", "pg_count_kmer_coverage", "pg_count_histogram",
    "pg_count_compute_histogram", "pg_count_distinct", "pg_count_capacity", "pg_count_destroy", "pg_histogram_peak",
    "pg_probtable_init", "pg_probtable_modify", "pg_probtable_get", "pg_probtable_free",
    "pg_result_layout", "pg_engine_create", "pg_engine_destroy", "pg_hmm_run", "pg_emission_run",
    "pg_fill_counts", "pg_genotype_run", "pg_engine_timings",
    "pg_count_device_arrays", "pg_count_export_counts", "pg_count_import_counts", "pg_count_kmers_seen", "pg_count_last_ms", "pg_count_clear",
    "pg_count_canonicalize", "pg_count_exchange_buffer", "pg_count_export_range", "pg_count_import_range", "pg_haplotype_sample",
    "pg_engine_load", "pg_engine_run_resident", "pg_engine_fetch", "pg_engine_run_counted", "pg_hmm_run_subsets",
    "pg_index_open", "pg_index_open_archive", "pg_index_close", "pg_index_kmer_size", "pg_index_n_chromosomes",
    "pg_index_chromosome_name", "pg_index_add_reference", "pg_index_segments_path", "pg_index_panel",
    "pg_hmm_run_samples", "pg_unique_kmers_compute", "pg_unique_kmers_panel", "pg_unique_kmers_stats", "pg_unique_kmers_write_tsv", "pg_unique_kmers_free",
]


def _sig(lib, name, restype, argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = argtypes
    return f


def bind(lib: C.CDLL, prefix: str = "pg_") -> C.CDLL:
    """Declares argument / result types of the functions `lib` exports under `prefix`."""
    p = prefix
    vp, u64, u32, i32, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_double

    def has(n):
        return hasattr(lib, p + n)

    _sig(lib, p + "last_error", C.c_char_p, [])
    if has("version"):
        _sig(lib, p + "version", C.c_char_p, [])
        _sig(lib, p + "device_count", i32, [])
        _sig(lib, p + "kernel_launches", u64, [])
    if has("count_new"):
        if p == "pg_":
            _sig(lib, p + "count_new", vp, [u32, u64, i32])
            _sig(lib, p + "count_create", vp, [C.c_char_p, C.c_char_p, u32, u64, i32])
            _sig(lib, p + "count_create_from_buffers", vp, [vp, u64, vp, u64, u32, u64, i32])
            _sig(lib, p + "count_feed_device", i32, [vp, vp, u64, i32])
            _sig(lib, p + "count_capacity", u64, [vp])
            _sig(lib, p + "count_device_arrays", i32, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)])
            _sig(lib, p + "count_export_counts", i32, [vp])
            _sig(lib, p + "count_import_counts", i32, [vp])
            _sig(lib, p + "count_kmers_seen", u64, [vp])
            _sig(lib, p + "count_last_ms", dbl, [vp])
            _sig(lib, p + "count_last_probe_ms", dbl, [vp, C.POINTER(u32)])
            _sig(lib, p + "count_clear", i32, [vp])
            _sig(lib, p + "count_canonicalize", i32, [vp])
            _sig(lib, p + "count_exchange_buffer", i32, [vp, u64, C.POINTER(u64)])
            _sig(lib, p + "count_export_range", i32, [vp, u64, u64])
            _sig(lib, p + "count_import_range", i32, [vp, u64, u64])
        else:
            _sig(lib, p + "count_new", vp, [u32])
            _sig(lib, p + "count_create_from_buffers", vp, [vp, u64, vp, u64, u32])
            _sig(lib, p + "count_feed_mt", i32, [vp, vp, u64, i32, i32])
        _sig(lib, p + "count_feed", i32, [vp, vp, u64, i32])
        _sig(lib, p + "count_lookup_ascii", i32, [vp, vp, u64, vp])
        _sig(lib, p + "count_lookup", i32, [vp, vp, u64, vp])
        _sig(lib, p + "count_kmer_coverage", i32, [vp, u64, C.POINTER(u64)])
        _sig(lib, p + "count_histogram", i32, [vp, u64, vp])
        _sig(lib, p + "count_compute_histogram", i32, [vp, u64, i32, C.c_char_p, C.POINTER(u64)])
        _sig(lib, p + "count_distinct", u64, [vp])
        _sig(lib, p + "count_destroy", None, [vp])
    if has("unique_kmers_compute"):
        if p == "pg_":
            _sig(lib, p + "unique_kmers_compute", vp, [i32, vp, C.POINTER(PgVariants)])
            _sig(lib, p + "unique_kmers_stats", i32, [vp, C.POINTER(dbl), C.POINTER(u64)])
            _sig(lib, p + "unique_kmers_write_tsv", i32, [vp, C.c_char_p, vp, C.c_char_p])
        else:
            _sig(lib, p + "unique_kmers_compute", vp, [vp, C.POINTER(PgVariants)])
        _sig(lib, p + "unique_kmers_panel", i32, [vp, C.POINTER(PgPanel)])
        _sig(lib, p + "unique_kmers_free", None, [vp])
    if has("histogram_peak"):
        _sig(lib, p + "histogram_peak", i32, [vp, u64, i32, C.POINTER(u64)])
    if has("probtable_init"):
        _sig(lib, p + "probtable_init", i32, [C.POINTER(PgProbTable), C.c_uint16, C.c_uint16, C.c_uint16, dbl])
        _sig(lib, p + "probtable_modify", i32, [C.POINTER(PgProbTable), C.c_uint16, C.c_uint16, dbl, dbl, dbl])
        _sig(lib, p + "probtable_get", dbl, [C.POINTER(PgProbTable), C.c_uint16, C.c_uint16, i32])
        _sig(lib, p + "probtable_free", None, [C.POINTER(PgProbTable)])
        _sig(lib, p + "result_layout", i32, [C.POINTER(PgPanel), vp])
    if has("engine_create"):
        _sig(lib, p + "engine_create", vp, [i32])
        _sig(lib, p + "engine_destroy", None, [vp])
        _sig(lib, p + "engine_timings", i32, [vp, C.POINTER(PgTimings)])
        _sig(lib, p + "hmm_run", i32, [vp, u32, C.POINTER(PgPanel), C.POINTER(PgProbTable), C.POINTER(PgHmmParams), C.POINTER(PgHmmResult)])
        _sig(lib, p + "hmm_run_samples", i32, [vp, u32, u32, C.POINTER(PgPanel), vp, vp, vp, C.POINTER(PgHmmParams), C.POINTER(PgHmmResult)])
        _sig(lib, p + "index_open", vp, [C.c_char_p, i32])
        _sig(lib, p + "index_open_archive", vp, [C.c_char_p])
        _sig(lib, p + "index_close", None, [vp])
        _sig(lib, p + "index_kmer_size", u32, [vp])
        _sig(lib, p + "index_n_chromosomes", u32, [vp])
        _sig(lib, p + "index_chromosome_name", C.c_char_p, [vp, u32])
        _sig(lib, p + "index_add_reference", i32, [vp])
        _sig(lib, p + "index_segments_path", C.c_char_p, [vp])
        _sig(lib, p + "index_panel", i32, [vp, u32, C.POINTER(PgPanel)])
        _sig(lib, p + "hmm_run_subsets", i32, [vp, u32, C.POINTER(PgPanel), C.POINTER(PgProbTable), C.POINTER(PgHmmParams), u32, vp, vp,
                                                 C.POINTER(PgHmmResult)])
        _sig(lib, p + "emission_run", i32, [vp, C.POINTER(PgPanel), C.POINTER(PgProbTable), vp, vp, vp])
        _sig(lib, p + "fill_counts", i32, [vp, vp, u64, u32, C.POINTER(PgPanel)])
        _sig(lib, p + "engine_load", i32, [vp, u32, C.POINTER(PgPanel), C.POINTER(PgHmmResult)])
        _sig(lib, p + "engine_run_resident", i32, [vp, vp, u64, vp, u64, u32, u64, dbl, C.POINTER(PgHmmParams), C.POINTER(u64)])
        _sig(lib, p + "engine_fetch", i32, [vp, u32, C.POINTER(PgPanel), C.POINTER(PgHmmResult)])
        _sig(lib, p + "engine_run_counted", i32, [vp, vp, i32, dbl, C.POINTER(PgHmmParams), C.POINTER(u64)])
        _sig(lib, p + "haplotype_sample", i32, [i32, C.POINTER(PgPanel), u32, dbl, dbl, i32, C.c_uint16, vp, vp, vp, vp, vp])
        _sig(lib, p + "genotype_run", i32, [vp, C.POINTER(PgGenotypeInput), u32, C.POINTER(PgPanel), C.POINTER(PgHmmParams), C.POINTER(PgHmmResult), C.POINTER(u64)])
    else:
        # oracles: same data arguments, no engine handle
        if has("hmm_run"):
            _sig(lib, p + "hmm_run", i32, [u32, C.POINTER(PgPanel), C.POINTER(PgProbTable), C.POINTER(PgHmmParams), C.POINTER(PgHmmResult)])
            _sig(lib, p + "hmm_run_mt", i32, [u32, C.POINTER(PgPanel), C.POINTER(PgProbTable), C.POINTER(PgHmmParams), C.POINTER(PgHmmResult), i32])
            _sig(lib, p + "emission_run", i32, [C.POINTER(PgPanel), C.POINTER(PgProbTable), vp, vp, vp])
        if has("fill_counts"):
            _sig(lib, p + "fill_counts", i32, [vp, u64, u32, C.POINTER(PgPanel)])
    return lib


_lib = None


def load() -> C.CDLL:
    """Loads libpangenie_b200.so (built in-tree by `make lib` / __graft_entry__.build()).

    There is deliberately no fallback: a missing or unloadable CUDA library is an error.
    """
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `make lib` (or __graft_entry__.build()); "
                               "pangenie_b200 has no CPU fallback")
        _lib = bind(C.CDLL(LIB_PATH), "pg_")
    return _lib


def ptr(a) -> int | None:
    """Address of a numpy array (None -> NULL)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "need a C-contiguous numpy array"
    return a.ctypes.data
