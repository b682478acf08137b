"""CPU oracle for the GAOT forward/backward hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (torch-CPU tensors, numpy for the
integer graph work) of the reference algorithm, pinned against golden vectors that
`oracle/make_golden.py` produced by running the *unmodified* reference `src/model` package in the
build container (see `tests/golden/`).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it - as the checker / the CPU baseline, never
as the thing shipped.  Nothing under `gaot_b200/` imports `oracle`.
"""
