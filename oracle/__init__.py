"""TEST INFRASTRUCTURE ONLY -- never imported by the product package ``fumi_b200``.

``oracle/`` holds CPU restatements of the reference algorithm for the FuMI episodic
inner-loop path plus the shims that run the *unmodified* reference
from ``/root/reference/fumi`` in this container to pin those restatements.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import
from here, and only as the checker -- never as the thing measured or shipped.

PARITY STATUS: the reference ships no tests / golden vectors.  The
episode math (``episode_np``/``episode_torch``) is pinned against outputs of the reference
itself run here through ``ref_shims`` (fixtures in ``tests/golden/``, written by
``make_golden.py``).  The episodic *sampler* depends on torchmeta==1.7.0, which is not
vendored in /root/reference and not installed; ``torchmeta_shim`` restates its published
algorithm from memory, so for the sampler: **parity unpinned** against real torchmeta
(pinned only against the reference's own ``dataset/data.py`` driven through that shim).
"""
