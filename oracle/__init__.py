"""CPU oracle for the EIGen fitness hot path.  TEST INFRASTRUCTURE ONLY.

A from-scratch restatement (torch-CPU / numpy / cv2) of the reference's algorithm for
  genome -> CPPN render -> PredNet (20 + 2 steps) -> Shi-Tomasi + pyramidal LK -> motion score.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package, and only as the checker.  The product package
(`evolutionary_illusion_generator_b200`) never imports it.

Parity pinning status (see DESIGN.md "Oracle"):
  * cppn / grid / render  : pinned against the reference's own code run under import stubs
                            (tests/golden/make_golden.py; byte-equal) and its 4 test_cppn.py cases.
  * scoring               : pinned against the reference's fitness_calculator.py (same harness).
  * optical flow          : pinned against the cv2 4.13 binary the reference calls (cv2 is third-party,
                            unpinned by the reference; restated in numpy in flow.py).
  * PredNet + whole path  : pinned against the reference's own net.py / call_prednet.py / get_fitnesses_neat,
                            executed unmodified with Chainer replaced by the functional stand-in
                            tests/golden/chainer_shim (tests/golden/reference_pipeline.npz: frames within 1 LSB,
                            fitness within 4e-4 relative).  Unpinned: the arithmetic inside Chainer's own
                            primitives (Chainer is not installable offline; restated from its documentation).
"""
