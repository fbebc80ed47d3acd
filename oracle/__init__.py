"""CPU oracle of the SPN4CIR hot path -- TEST INFRASTRUCTURE ONLY.

A restatement of the reference's algorithm (plain torch fp32/fp64 on CPU and numpy), each function
citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it, and only as the checker or the timed CPU baseline;
nothing under spn4cir_b200/ imports it.

Pinning: the reference ships no tests or golden vectors.  The oracle is pinned
against the reference's own code executed in the build container (oracle/make_golden.py imports
/root/reference/clip4cir/{models_negplus,validate}.py unmodified) through the committed fixtures in
tests/golden/.  Two parts of the path live in files that cannot be imported offline -- the BLIP-2 tails (blip24cir
needs omegaconf / timm / iopath) and the pair-mining script zscir/srm_utils.py (module globals + dataset files).  Their
hot-path lines are straight-line tensor code, so make_golden.py EXECUTES THOSE SOURCE LINES VERBATIM
(exec_reference_lines: blip2_qformer_cir_align_prompt.py:254-265 and :314-318, srm_utils.py:121-124 and :145-180, and the
submission writer clip4cir/cirr_test_submission.py:84-109) on prepared locals, checks the restatements against the result
and commits it as fixtures (blip2_ref_*.npz, srm_ref.npz, cirr_submit_ref.npz).
"""
