"""CPU oracle for the PQL learner hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pql_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do.  It restates the reference algorithm
(``/root/reference/pql/...``; every function cites the file:line it follows):

* ``oracle.replay``  - numpy restatement of the ring buffer, the P-learner
  observation ring and the n-step window (byte/index work, bit-exact).
* ``oracle.learner`` - torch-CPU fp32 restatement of the twin-Q / C51 critic
  update, the DPG actor update, clip + AdamW + Polyak (the reference's own
  arithmetic *is* PyTorch, so the float oracle is PyTorch fp32 on the CPU).

Parity pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the reference itself, produced in the
build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference`` under import stubs) and committed under ``tests/golden/``.
"""
