"""CPU oracle for the OCR forward hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this package.  The product (advancedliteratemachinery_b200) never does: it fails loudly when the
CUDA library is missing instead of falling back to anything here.
"""
