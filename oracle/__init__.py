"""CPU oracle for the graph-CF hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement (numpy for the integer/index work, torch-CPU fp32/fp64 for the floating-point work) of the
reference algorithms that recommendation_b200's CUDA kernels replace.  Every function cites the reference
file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; nothing under recommendation_b200/ does.

Pinning: the reference (Cmint22/Recommendation) ships no tests, seeds or golden vectors ("parity unpinned" by
its own tests).  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF: the
importable reference modules were run in the build container on fixed-seed inputs by
tests/golden/make_golden.py and their inputs/outputs/gradients committed as tests/golden/*.npz;
tests/test_oracle_golden.py checks every oracle function against them.  lightgcn.py cannot be imported
(torch_geometric absent, version unpinned): its LGConv arithmetic is restated from PyG's published
gcn_norm / propagate semantics in oracle/lightgcn_ref.py and cross-checked against the importable
selfcf.LGCN_Encoder fixtures (same operator on symmetric graphs).
"""
