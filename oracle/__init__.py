"""CPU ORACLE — test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this
package.  The product path (openea_b200, openea) never does; it fails loudly when liboea.so is missing.

  oracle.triple  : ctypes binding of oea_oracle.c (path (i), TF-1 graph restated; pinned to the reference's own graph
                   code executed on oracle.tf1_shim — tests/test_reference_graph_goldens.py; TF's op / optimiser
                   semantics themselves stay unpinned, see the header of oea_oracle.c)
  oracle.triple_ext : float64 torch-autograd restatement of the TransH / TransD / DistMult / SimplE graphs (pinned the same way)
  oracle.tf1_shim   : a TensorFlow-1 graph interpreter on torch float64; scripts/make_golden_path_i.py runs the reference's
                   graph-definition code on it to produce tests/golden/path_i_*.npz and path_ii_*.npz
  oracle.reference_source : pure-Python helper functions of the reference compiled from its source (where it is present)
  oracle.finding : NumPy restatement of modules/finding/{similarity,alignment}.py and the filter/top-k part of
                   modules/bootstrapping/alignment_finder.py (path (iii)); pinned against golden vectors
                   generated from the reference itself (tests/golden/make_golden.py)
  oracle.sampler : pure-Python restatement of modules/train/batch.py (small cases only)
"""
