"""CPU oracle: a torch-CPU fp32/fp64 restatement of the reference's forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tensorflow-image-models_b200/`` imports this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm do,
and there only as the checker / the timed CPU stand-in -- never as the product path.

Why a restatement: the reference's arithmetic lives in TensorFlow 2.12 / Keras 2.12
(poetry.lock:1405-1406, 608-609), which is not installed here or on the GPU box, and the
reference ships no golden vectors (its only numeric test is a live comparison with timm,
tests/test_timm.py:38-71).  Each function cites the reference file:line it follows.

Parity pin status (see DESIGN.md "Oracle"):
  * against TensorFlow itself: UNPINNED (TF cannot run in this image);
  * graph structure pinned two ways in tests/: (1) the reference's OWN model code executed on a
    torch-backed stand-in for the handful of Keras layers it uses (tools/ref_shim) -> committed
    golden logits under tests/golden/; (2) torchvision's independent implementations of the same
    timm-parameterised networks, with weights mapped by the reference's conversion rules
    (tfimm/utils/timm.py:39-106).
"""
