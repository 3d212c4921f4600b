"""Native reader of the reference's index artefacts (csrc/index_io.cu) against the reference's own
test-data files (byte copies under tests/golden/counting/, see test_golden_reference_fixture.py) and the independent
Python decoder pangenie_b200/refindex.py.  Host code only: runs without a GPU."""
import gzip
import os
import shutil

import numpy as np
import pytest

import pangenie_b200 as pg
from pangenie_b200 import refindex

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "counting")
FIELDS = ("positions", "path_to_allele", "coverage", "kmer_offsets", "kmer_counts", "allele_offsets", "allele_ids",
          "allele_undefined", "allele_kmer_offset", "allele_kmer_mask")


def _same(a, b, with_kmers):
    assert a.n_paths == b.n_paths and a.n_variants == b.n_variants
    for f in FIELDS + (("kmer_codes", "flank_offsets", "flank_codes") if with_kmers else ()):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def test_prefix_index_matches_python_decoder():
    ix = pg.Index(os.path.join(G, "index"))
    assert ix.kmer_size == 31 and ix.chromosomes == ["chr1"]
    assert ix.segments_path.endswith("index_path_segments.fasta") and os.path.exists(ix.segments_path)
    k, panels, add_ref = refindex.read_unique_kmers_map(os.path.join(G, "index_UniqueKmersMap.cereal"))
    want = refindex.attach_kmers_tsv(panels["chr1"], os.path.join(G, "index_chr1_kmers.tsv.gz"))
    assert ix.add_reference == add_ref
    got = ix.panel(0)
    assert got.n_paths == 215 and got.n_variants == 2          # the reference's CommandsTest fixture
    _same(got, want, True)


def test_single_archive_with_filled_counts():
    ix = pg.Index(archive=os.path.join(G, "region_UniqueKmersList.cereal"))
    _, panels, _ = refindex.read_unique_kmers_map(os.path.join(G, "region_UniqueKmersList.cereal"))
    got = ix.panel(0)
    _same(got, panels["chr1"], False)
    assert got.kmer_counts.max() > 0 and got.kmer_codes is None     # counts as produced by the real PanGenie + jellyfish


def test_errors_are_reported(tmp_path):
    with pytest.raises(pg.PgError, match="cannot be opened"):
        pg.Index(str(tmp_path / "nothing"))
    data = open(os.path.join(G, "index_UniqueKmersMap.cereal"), "rb").read()
    (tmp_path / "cut_UniqueKmersMap.cereal").write_bytes(data[:len(data) // 2])
    with pytest.raises(pg.PgError, match="truncated|corrupt|unknown"):
        pg.Index(str(tmp_path / "cut"), with_kmers=False)
    (tmp_path / "tail_UniqueKmersMap.cereal").write_bytes(data + b"\\0")
    with pytest.raises(pg.PgError, match="trailing"):
        pg.Index(str(tmp_path / "tail"), with_kmers=False)
    # a k-mer table that does not belong to the archive
    shutil.copy(os.path.join(G, "index_UniqueKmersMap.cereal"), tmp_path / "bad_UniqueKmersMap.cereal")
    with gzip.open(os.path.join(G, "index_chr1_kmers.tsv.gz"), "rt") as f:
        lines = f.read().splitlines()
    with gzip.open(tmp_path / "bad_chr1_kmers.tsv.gz", "wt") as f:
        f.write("\\n".join(lines[:-1]) + "\\n")                        # one variant missing
    with pytest.raises(pg.PgError, match="fewer variants"):
        pg.Index(str(tmp_path / "bad"))
