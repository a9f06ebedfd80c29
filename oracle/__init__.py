"""CPU oracle for the DeepHumor caption-generation path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deephumor_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do.

This is a functional (state_dict in, tensors out) restatement of the reference algorithm
in torch-CPU fp32 ops.  The arithmetic of the reference lives in third-party, un-vendored,
unpinned ``torch`` / ``torchvision`` (``/root/reference/requirements.txt:2-3``; installed
here: torch 2.11.0+cu128, torchvision 0.26.0+cu128).  Every function cites the reference
``file:line`` it follows.

Pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against outputs of the *unmodified reference itself*, imported in the build
container by ``oracle/make_golden.py`` and committed as fixtures under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks the oracle against those fixtures on any box, and
``tests/test_oracle_vs_reference.py`` re-checks live when ``/root/reference`` is present.
"""
