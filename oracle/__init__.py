"""CPU oracle for the STMiner spatial-pattern hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``stminer_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
the timed CPU baseline, never as the product path.

Each function cites the reference file:line (relative to the STMiner source
tree, v1.1.4) whose arithmetic it restates.  Where the arithmetic lives in a
third-party dependency (scikit-learn GaussianMixture, scipy
linear_sum_assignment, POT emd2) the module header names the dependency, its
pinned version and how the restatement was pinned.
"""
