"""Synthetic workloads of the shapes BASELINE.json names.

Test and bench tooling only: nothing in the product package (pangenie_b200/) imports this.
  small.py  numpy generator (SNP bubbles, host memory) used by the unit-sized parity tests
  large.py  torch generator (SNPs + indels + tri-allelic + undefined alleles, per-chromosome streams, runs on the GPU when
            one is present) used for the full BASELINE.json configurations and the parity tests at size
"""
