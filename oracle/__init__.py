"""ORACLE — CPU restatements of the reference algorithm for the frame-scoring
path.  TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs, never from the product
package (ivos-w_b200/), which fails loudly when its CUDA library is missing.

Parity status of each file is stated in its header (all PINNED against goldens
generated from the reference's own modules, except where a header says
"parity unpinned").
"""
