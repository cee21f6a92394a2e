"""CPU ORACLE — test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this
package.  The product path (openea_b200, openea) never does; it fails loudly when liboea.so is missing.

  oracle.triple  : ctypes binding of oea_oracle.c (path (i), TF-1 graph restated; PARITY UNPINNED, see the
                   header of oea_oracle.c)
  oracle.triple_ext : float64 torch-autograd restatement of the TransH / TransD / DistMult / SimplE graphs (PARITY UNPINNED)
  oracle.finding : NumPy restatement of modules/finding/{similarity,alignment}.py and the filter/top-k part of
                   modules/bootstrapping/alignment_finder.py (path (iii)); pinned against golden vectors
                   generated from the reference itself (tests/golden/make_golden.py)
  oracle.sampler : pure-Python restatement of modules/train/batch.py (small cases only)
"""
