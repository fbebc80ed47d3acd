"""CPU oracle of the SPN4CIR hot path -- TEST INFRASTRUCTURE ONLY.

A restatement of the reference's algorithm (plain torch fp32/fp64 on CPU and numpy), each function
citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it, and only as the checker or the timed CPU baseline;
nothing under spn4cir_b200/ imports it.

Pinning: the reference ships no tests or golden vectors.  The oracle is pinned
against the reference's own code executed in the build container (oracle/make_golden.py imports
/root/reference/clip4cir/{models_negplus,validate}.py unmodified) through the committed fixtures in
tests/golden/.  The BLIP-2 tails (blip24cir) cannot be imported offline (omegaconf/timm missing):
oracle.bank_ce.blip2_stage2_loss_loop restates blip2_qformer_cir_align_prompt.py:254-268 line by line and is
PARITY-UNPINNED against an execution of the reference itself; it is cross-checked only against its own
closed form.
"""
