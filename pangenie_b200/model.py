"""Host-side mirror of the reference's interface for the genotyping hot path, over the C-ABI.

Names and argument meaning follow the reference (file:line into the reference tree):
  KmerCounter / JellyfishCounter   src/kmercounter.hpp:9-24, src/jellyfishcounter.cpp:26-153
  ProbabilityTable                 src/probabilitytable.hpp:12-30
  HMM                              src/hmm.hpp:26-47
Everything computes on the GPU through libpangenie_b200.so; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi
from .capi import PG_OK, PgGenotypeInput, PgHmmParams, PgHmmResult, PgPanel, PgProbTable, PgTimings, ptr
from .panel import Panel, Result


class PgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pangenie_b200 error {code}: {msg}")
        self.code = code


def _check(lib, st: int):
    if st != PG_OK:
        raise PgError(st, lib.pg_last_error().decode())


def _bytes_arg(b):
    """(address, length, keepalive) for bytes / bytearray / numpy uint8 / torch uint8 tensors."""
    if b is None:
        return None, 0, None
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b.view(np.uint8))
        return a.ctypes.data, a.size, a
    if hasattr(b, "data_ptr"):  # torch tensor (host pinned or device)
        return b.data_ptr(), b.numel() * b.element_size(), b
    a = np.frombuffer(b, dtype=np.uint8)
    return a.ctypes.data, a.size, a


class ProbabilityTable:
    """ProbabilityTable(cov_min, cov_max, count_max, regularization_const) (src/probabilitytable.cpp:28-45)."""

    def __init__(self, cov_min: int, cov_max: int, count_max: int, regularization_const: float):
        self._lib = capi.load()
        self.t = PgProbTable()
        _check(self._lib, self._lib.pg_probtable_init(C.byref(self.t), cov_min, cov_max, count_max, regularization_const))

    def modify_probability(self, kmer_coverage: int, read_kmer_count: int, p0: float, p1: float, p2: float):
        """modify_probability(cov, count, CopyNumber(p0,p1,p2)) (src/probabilitytable.cpp:67-73)."""
        _check(self._lib, self._lib.pg_probtable_modify(C.byref(self.t), kmer_coverage, read_kmer_count, p0, p1, p2))

    def get_probability(self, kmer_coverage: int, read_kmer_count: int):
        return tuple(self._lib.pg_probtable_get(C.byref(self.t), kmer_coverage, read_kmer_count, cn) for cn in range(3))

    def __del__(self):
        try:
            self._lib.pg_probtable_free(C.byref(self.t))
        except Exception:
            pass


def copy_number(p0, p1, p2, regularization=None):
    """CopyNumber(cn_0, cn_1, cn_2[, regularization_const]) (src/copynumber.cpp:14-28) -> (p0, p1, p2)."""
    if regularization is None:
        return (p0, p1, p2)
    s = p0 + p1 + p2 + 3.0 * regularization
    a, b = (p0 + regularization) / s, (p1 + regularization) / s
    return (a, b, 1.0 - a - b)


class KmerCounter:
    """JellyfishCounter (src/jellyfishcounter.cpp:26-153) on the device k-mer table."""

    def __init__(self, reads=None, segments=None, kmer_size: int = 31, hash_size: int = 3_000_000_000, device: int = 0,
                 max_distinct: int | None = None):
        self._lib = capi.load()
        self.k = kmer_size
        self._h = None
        if reads is None:
            h = self._lib.pg_count_new(kmer_size, max_distinct or hash_size, device)
        elif isinstance(reads, str):
            h = self._lib.pg_count_create(reads.encode(), segments.encode() if segments else None, kmer_size, hash_size, device)
        else:
            ra, rl, _k1 = _bytes_arg(reads)
            sa, sl, _k2 = _bytes_arg(segments)
            h = self._lib.pg_count_create_from_buffers(ra, rl, sa, sl, kmer_size, hash_size, device)
        if not h:
            raise PgError(-1, self._lib.pg_last_error().decode())
        self._h = h

    @property
    def handle(self):
        return self._h

    def feed(self, text, op: int):
        a, n, _keep = _bytes_arg(text)
        on_device = hasattr(text, "is_cuda") and text.is_cuda
        f = self._lib.pg_count_feed_device if on_device else self._lib.pg_count_feed
        _check(self._lib, f(self._h, a, n, op))

    def getKmerAbundance(self, kmer: str) -> int:
        out = np.zeros(1, np.uint64)
        b = np.frombuffer(kmer.encode(), dtype=np.uint8)
        assert len(b) == self.k
        _check(self._lib, self._lib.pg_count_lookup_ascii(self._h, b.ctypes.data, 1, out.ctypes.data))
        return int(out[0])

    def lookup(self, codes: np.ndarray) -> np.ndarray:
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        out = np.zeros(len(codes), np.uint64)
        _check(self._lib, self._lib.pg_count_lookup(self._h, ptr(codes), len(codes), ptr(out)))
        return out

    def computeKmerCoverage(self, genome_kmers: int) -> int:
        out = C.c_uint64(0)
        _check(self._lib, self._lib.pg_count_kmer_coverage(self._h, genome_kmers, C.byref(out)))
        return out.value

    def histogram(self, max_count: int = 10000) -> np.ndarray:
        bins = np.zeros(max_count + 1, np.uint64)
        _check(self._lib, self._lib.pg_count_histogram(self._h, max_count, ptr(bins)))
        return bins

    def computeHistogram(self, max_count: int, largest_peak: bool, filename: str = "") -> int:
        out = C.c_uint64(0)
        _check(self._lib, self._lib.pg_count_compute_histogram(self._h, max_count, int(largest_peak),
                                                                 filename.encode() if filename else None, C.byref(out)))
        return out.value

    def export_counts(self):
        _check(self._lib, self._lib.pg_count_export_counts(self._h))

    def import_counts(self):
        _check(self._lib, self._lib.pg_count_import_counts(self._h))

    def clear(self):
        _check(self._lib, self._lib.pg_count_clear(self._h))

    def canonicalize(self):
        """Deterministic key layout after PRIME (pg_count_canonicalize): identical on every GPU that primed the same file."""
        _check(self._lib, self._lib.pg_count_canonicalize(self._h))

    def exchange_buffer(self, n_slots: int) -> int:
        a = C.c_uint64(0)
        _check(self._lib, self._lib.pg_count_exchange_buffer(self._h, n_slots, C.byref(a)))
        return a.value

    def export_range(self, first_slot: int, n_slots: int):
        _check(self._lib, self._lib.pg_count_export_range(self._h, first_slot, n_slots))

    def import_range(self, first_slot: int, n_slots: int):
        _check(self._lib, self._lib.pg_count_import_range(self._h, first_slot, n_slots))

    def distinct(self) -> int:
        return int(self._lib.pg_count_distinct(self._h))

    def capacity(self) -> int:
        return int(self._lib.pg_count_capacity(self._h))

    def kmers_seen(self) -> int:
        return int(self._lib.pg_count_kmers_seen(self._h))

    def last_ms(self) -> float:
        return float(self._lib.pg_count_last_ms(self._h))

    def last_probe_ms(self):
        """(device ms spent in the probe passes of the last partitioned feed, number of passes)."""
        n = C.c_uint32(0)
        return float(self._lib.pg_count_last_probe_ms(self._h, C.byref(n))), int(n.value)

    def device_arrays(self):
        k, c, cap = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _check(self._lib, self._lib.pg_count_device_arrays(self._h, C.byref(k), C.byref(c), C.byref(cap)))
        return k.value, c.value, cap.value

    def close(self):
        if self._h:
            self._lib.pg_count_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hmm_params(recombrate=1.26, uniform=False, effective_N=25000.0, only_paths=None, normalize=True):
    """Arguments of HMM::HMM (src/hmm.hpp:38) after `probabilities`/`run_*`; returns (struct, keepalive)."""
    p = PgHmmParams()
    p.recombrate = float(recombrate)
    p.effective_N = float(effective_N)
    p.uniform = int(bool(uniform))
    p.normalize = int(bool(normalize))
    keep = None
    if only_paths is not None:
        keep = np.ascontiguousarray(only_paths, dtype=np.uint16)
        p.only_paths = keep.ctypes.data
        p.n_only_paths = len(keep)
    return p, keep


class Engine:
    """Per-device engine (pg_engine)."""

    def __init__(self, device: int = 0):
        self._lib = capi.load()
        self._h = self._lib.pg_engine_create(device)
        if not self._h:
            raise PgError(-1, self._lib.pg_last_error().decode())
        self.device = device

    def _panel_array(self, panels):
        arr = (PgPanel * len(panels))()
        for i, p in enumerate(panels):
            arr[i] = p.as_struct()
        return arr

    def hmm_run(self, panels, table: ProbabilityTable, results=None, **kw):
        """HMM forward-backward for a list of chromosome panels -> list[Result]."""
        if results is None:
            results = [Result(p) for p in panels]
        prm, _keep = hmm_params(**kw)
        pa = self._panel_array(panels)
        ra = (PgHmmResult * len(panels))()
        for i, r in enumerate(results):
            ra[i] = r.as_struct()
        _check(self._lib, self._lib.pg_hmm_run(self._h, len(panels), pa, C.byref(table.t), C.byref(prm), ra))
        return results

    def hmm_run_samples(self, panels, counts, coverages, tables, **kw):
        """Multi-sample batching (pg_hmm_run_samples): S samples on ONE index.  `panels` = the chromosomes'
        structure, `counts[s][c]` / `coverages[s][c]` = the filled k-mer counts / local coverages of sample s on chromosome c,
        `tables[s]` = its ProbabilityTable.  -> results[s][c]."""
        S, Cn = len(counts), len(panels)
        results = [[Result(p) for p in panels] for _ in range(S)]
        prm, _keep = hmm_params(**kw)
        pa = self._panel_array(panels)
        ra = (PgHmmResult * (S * Cn))()
        kc = (C.c_void_p * (S * Cn))()
        cv = (C.c_void_p * (S * Cn))()
        tb = (C.POINTER(type(tables[0].t)) * S)()
        keep = []
        for s_ in range(S):
            tb[s_] = C.pointer(tables[s_].t)
            for c_ in range(Cn):
                a = np.ascontiguousarray(counts[s_][c_], np.uint16)
                b = np.ascontiguousarray(coverages[s_][c_], np.uint16)
                if a.size == 0:
                    a = np.zeros(1, np.uint16)
                keep += [a, b]
                kc[s_ * Cn + c_], cv[s_ * Cn + c_] = a.ctypes.data, b.ctypes.data
                ra[s_ * Cn + c_] = results[s_][c_].as_struct()
        _check(self._lib, self._lib.pg_hmm_run_samples(self._h, S, Cn, pa, kc, cv, tb, C.byref(prm), ra))
        return results

    def hmm_run_subsets(self, panels, table: ProbabilityTable, subsets, results=None, **kw):
        """The reference's `-a` mode (src/commands.cpp:916-993): one un-normalised run per path subset, likelihoods added
        per variant (run_genotyping, :166-176), normalised at the end.  `subsets` = list of lists of path ids."""
        if results is None:
            results = [Result(p) for p in panels]
        prm, _keep = hmm_params(**kw)
        off = np.zeros(len(subsets) + 1, np.uint32)
        off[1:] = np.cumsum([len(x) for x in subsets])
        paths = np.ascontiguousarray(np.concatenate([np.asarray(x, np.uint16) for x in subsets]))
        pa = self._panel_array(panels)
        ra = (PgHmmResult * len(panels))()
        for i, r in enumerate(results):
            ra[i] = r.as_struct()
        _check(self._lib, self._lib.pg_hmm_run_subsets(self._h, len(panels), pa, C.byref(table.t), C.byref(prm), len(subsets),
                                                        ptr(off), ptr(paths), ra))
        return results

    def emission_run(self, panel: Panel, table: ProbabilityTable):
        """EmissionProbabilityComputer for every variant -> (offsets, dense (maxA+1)^2 matrices, log_scale)."""
        V = panel.n_variants
        off = np.zeros(V + 1, np.uint64)
        for v in range(V):
            n = panel.nr_alleles(v)
            off[v + 1] = off[v] + n * n
        em = np.zeros(int(off[-1]), np.float64)
        ls = np.zeros(V, np.float64)
        ps = panel.as_struct()
        _check(self._lib, self._lib.pg_emission_run(self._h, C.byref(ps), C.byref(table.t), ptr(off), ptr(em), ptr(ls)))
        return off, em, ls

    def fill_counts(self, counter: KmerCounter, kmer_abundance_peak: int, panels):
        pa = self._panel_array(panels)
        _check(self._lib, self._lib.pg_fill_counts(self._h, counter.handle, kmer_abundance_peak, len(panels), pa))

    def genotype_run(self, reads, segments, panels, k=31, hash_size=3_000_000_000, regularization=0.01,
                     histogram_path=None, results=None, **kw):
        """The `PanGenie -f` stage (src/commands.cpp:730-1084) with host buffers -> (results, peak)."""
        if results is None:
            results = [Result(p) for p in panels]
        prm, _keep = hmm_params(**kw)
        inp = PgGenotypeInput()
        ra_, rl, _k1 = _bytes_arg(reads)
        sa, sl, _k2 = _bytes_arg(segments)
        inp.reads, inp.reads_len, inp.segments, inp.segments_len = ra_, rl, sa, sl
        inp.k, inp.hash_size, inp.regularization = k, hash_size, regularization
        inp.histogram_path = histogram_path.encode() if histogram_path else None
        pa = self._panel_array(panels)
        ra = (PgHmmResult * len(panels))()
        for i, r in enumerate(results):
            ra[i] = r.as_struct()
        peak = C.c_uint64(0)
        _check(self._lib, self._lib.pg_genotype_run(self._h, C.byref(inp), len(panels), pa, C.byref(prm), ra, C.byref(peak)))
        return results, peak.value

    def load(self, panels, results=None):
        """Uploads the panels (incl. k-mer codes) to HBM; returns the Result buffers to fetch into."""
        if results is None:
            results = [Result(p) for p in panels]
        self._resident = (panels, results)
        pa = self._panel_array(panels)
        ra = (PgHmmResult * len(panels))()
        for i, r in enumerate(results):
            ra[i] = r.as_struct()
        _check(self._lib, self._lib.pg_engine_load(self._h, len(panels), pa, ra))
        return results

    def run_resident(self, d_reads, d_segments, k=31, hash_size=3_000_000_000, regularization=0.01, **kw) -> int:
        """Whole stage from device-resident text (torch uint8 cuda tensors); returns the k-mer abundance peak."""
        prm, _keep = hmm_params(**kw)
        ra_, rl, _k1 = _bytes_arg(d_reads)
        sa, sl, _k2 = _bytes_arg(d_segments)
        peak = C.c_uint64(0)
        _check(self._lib, self._lib.pg_engine_run_resident(self._h, ra_, rl, sa, sl, k, hash_size, regularization, C.byref(prm), C.byref(peak)))
        return peak.value

    def run_counted(self, counter: "KmerCounter", largest_peak: bool = True, regularization=0.01, **kw) -> int:
        """Histogram peak -> table -> fill -> HMM on the loaded panels against an externally filled counter."""
        prm, _keep = hmm_params(**kw)
        peak = C.c_uint64(0)
        _check(self._lib, self._lib.pg_engine_run_counted(self._h, counter.handle, int(largest_peak), regularization, C.byref(prm), C.byref(peak)))
        return peak.value

    def fetch(self):
        panels, results = self._resident
        pa = self._panel_array(panels)
        ra = (PgHmmResult * len(panels))()
        for i, r in enumerate(results):
            ra[i] = r.as_struct()
        _check(self._lib, self._lib.pg_engine_fetch(self._h, len(panels), pa, ra))
        return results

    def timings(self) -> dict:
        t = PgTimings()
        _check(self._lib, self._lib.pg_engine_timings(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in PgTimings._fields_}

    def close(self):
        if self._h:
            self._lib.pg_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def haplotype_sample(panel: Panel, size: int, recombrate: float = 1.26, effective_N: float = 25000.0, add_reference: bool = False,
                     allele_penalty: int = 10, device: int = 0):
    """HaplotypeSampler(&unique_kmers, size, recombrate, effective_N, &best_scores, add_reference, "", chromosome,
    allele_penalty) (src/haplotypesampler.cpp:20-78) on the device -> (sampled_paths [n_out, V], best_scores [size],
    new_path_to_allele [V, n_out], new_kmer_count [V], new_counts)."""
    lib = capi.load()
    V, n_out = panel.n_variants, size + (1 if add_reference else 0)
    paths = np.zeros((n_out, V), np.uint64)
    scores = np.zeros(size, np.uint32)
    p2a = np.zeros((V, n_out), np.uint16)
    nk = np.zeros(V, np.uint32)
    counts = np.zeros(max(len(panel.kmer_counts), 1), np.uint16)
    ps = panel.as_struct()
    _check(lib, lib.pg_haplotype_sample(device, C.byref(ps), size, recombrate, effective_N, int(add_reference), allele_penalty,
                                        ptr(paths), ptr(scores), ptr(p2a), ptr(nk), ptr(counts)))
    return paths, scores, p2a, nk, counts[:int(nk.sum())]


class HMM:
    """HMM(unique_kmers, probabilities, run_genotyping, run_phasing, recombrate, uniform, effective_N,
    only_paths, normalize) (src/hmm.hpp:38).  Viterbi phasing is out of scope."""

    _engines: dict = {}

    def __init__(self, unique_kmers: Panel, probabilities: ProbabilityTable, run_genotyping=True, run_phasing=False,
                 recombrate=1.26, uniform=False, effective_N=25000.0, only_paths=None, normalize=True, device=0):
        if run_phasing:
            raise NotImplementedError("Viterbi phasing (-p) is outside the accelerated path")
        eng = HMM._engines.get(device)
        if eng is None:
            eng = HMM._engines[device] = Engine(device)
        self.panel = unique_kmers
        self.result = Result(unique_kmers)
        if run_genotyping:
            eng.hmm_run([unique_kmers], probabilities, [self.result], recombrate=recombrate, uniform=uniform,
                        effective_N=effective_N, only_paths=only_paths, normalize=normalize)

    def get_genotyping_result(self) -> Result:
        return self.result


def panel_from_struct(ps: PgPanel) -> Panel:
    """numpy COPIES of the library-owned arrays of a pg_panel."""
    V, P = ps.n_variants, ps.n_paths

    def arr(addr, n, dt):
        if not addr or n == 0:
            return np.zeros(0, dt) if addr or n == 0 else None
        return np.ctypeslib.as_array(C.cast(addr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy()
    koff = arr(ps.kmer_offsets, V + 1, np.uint32)
    aoff = arr(ps.allele_offsets, V + 1, np.uint32)
    K, A = int(koff[-1]) if V else 0, int(aoff[-1]) if V else 0
    pan = Panel(P, arr(ps.positions, V, np.uint64), arr(ps.path_to_allele, V * P, np.uint16), arr(ps.coverage, V, np.uint16),
                koff, arr(ps.kmer_counts, K, np.uint16), aoff, arr(ps.allele_ids, A, np.uint16),
                arr(ps.allele_undefined, A, np.uint8), arr(ps.allele_kmer_offset, A, np.uint16), arr(ps.allele_kmer_mask, A, np.uint32))
    if ps.flank_offsets:
        foff = arr(ps.flank_offsets, V + 1, np.uint32)
        pan.kmer_codes = arr(ps.kmer_codes, K, np.uint64)
        pan.flank_offsets = foff
        pan.flank_codes = arr(ps.flank_codes, int(foff[-1]) if V else 0, np.uint64)
    return pan


def variants_struct(flat: dict):
    """pg_variants over the flat numpy arrays of one chromosome (keys = the struct's field names); returns
    (struct, keep-alive list)."""
    from .capi import PgVariants
    vs = PgVariants()
    vs.n_variants, vs.n_paths, vs.k = int(flat["n_variants"]), int(flat["n_paths"]), int(flat["k"])
    keep = []
    for name, dt in (("positions", np.uint64), ("end_positions", np.uint64), ("path_to_allele", np.uint16),
                     ("allele_offsets", np.uint32), ("allele_undefined", np.uint8), ("seq_offsets", np.uint64), ("seq", np.uint8),
                     ("left_offsets", np.uint64), ("left_seq", np.uint8), ("right_offsets", np.uint64), ("right_seq", np.uint8)):
        a = flat.get(name)
        if a is None:
            continue
        a = np.ascontiguousarray(a, dt)
        if a.size == 0:
            a = np.zeros(1, dt)          # a valid address for empty arrays
        keep.append(a)
        setattr(vs, name, a.ctypes.data)
    return vs, keep


class UniqueKmerSelection:
    """The index stage's unique-k-mer selection for one chromosome on the device (pg_unique_kmers_*; reference
    src/stepwiseuniquekmercomputer.cpp:95-197).  `graph_counts` is a Counter holding the COUNT of the path segments."""

    def __init__(self, graph_counts: "KmerCounter", flat: dict, device: int = 0):
        self._lib = capi.load()
        vs, keep = variants_struct(flat)
        self._h = self._lib.pg_unique_kmers_compute(device, graph_counts._h, C.byref(vs))
        if not self._h:
            raise PgError(-1, self._lib.pg_last_error().decode())

    def panel(self) -> Panel:
        ps = PgPanel()
        _check(self._lib, self._lib.pg_unique_kmers_panel(self._h, C.byref(ps)))
        return panel_from_struct(ps)

    def stats(self):
        ms, n = C.c_double(), C.c_uint64()
        _check(self._lib, self._lib.pg_unique_kmers_stats(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def write_tsv(self, chromosome: str, end_positions, path: str):
        ends = np.ascontiguousarray(end_positions, np.uint64)
        _check(self._lib, self._lib.pg_unique_kmers_write_tsv(self._h, chromosome.encode(), ends.ctypes.data, os.fsencode(path)))

    def close(self):
        if self._h:
            self._lib.pg_unique_kmers_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Index:
    """The index artefacts of `PanGenie-index` read by the native reader (csrc/index_io.cu, pg_index_*): what
    `PanGenie -f <prefix>` loads before the hot path starts (reference src/commands.cpp:760-790, 98-137)."""

    def __init__(self, prefix: str | None = None, with_kmers: bool = True, archive: str | None = None):
        self._lib = capi.load()
        if archive is not None:
            self._h = self._lib.pg_index_open_archive(os.fsencode(archive))
        else:
            self._h = self._lib.pg_index_open(os.fsencode(prefix), int(with_kmers))
        if not self._h:
            raise PgError(-1, self._lib.pg_last_error().decode())

    @property
    def kmer_size(self) -> int:
        return int(self._lib.pg_index_kmer_size(self._h))

    @property
    def add_reference(self) -> bool:
        return bool(self._lib.pg_index_add_reference(self._h))

    @property
    def segments_path(self) -> str:
        return self._lib.pg_index_segments_path(self._h).decode()

    @property
    def chromosomes(self):
        return [self._lib.pg_index_chromosome_name(self._h, i).decode() for i in range(self._lib.pg_index_n_chromosomes(self._h))]

    def panel(self, i: int) -> Panel:
        """Chromosome i as a Panel (numpy COPIES of the index-owned arrays)."""
        ps = PgPanel()
        _check(self._lib, self._lib.pg_index_panel(self._h, i, C.byref(ps)))
        return panel_from_struct(ps)

    def close(self):
        if self._h:
            self._lib.pg_index_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
