"""CPU oracle for the CALAMITY gain-and-foreground fit path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``calamity_b200``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
there only as the checker / the timed CPU baseline -- never as the thing shipped.

What is in here

* ``restatement.py``  -- NumPy restatement (float32 or float64) of the reference's
  marshalling, model, loss, analytic gradient, Keras-v2 optimizer rules and fit loop
  (``/root/reference/calamity/calibration.py``; each function cites the lines it follows).
* ``torch_port.py``   -- op-for-op torch-CPU transliteration of the TensorFlow graph
  (dense broadcast-multiply-reduce, autograd, same optimizers); used to cross-check the
  analytic gradient and as the timed CPU baseline (kind "port").
* ``tf_shim/``        -- a minimal fake ``tensorflow`` / ``pyuvdata`` / ``hera_filters``
  backed by torch+numpy, good enough to IMPORT AND RUN the reference's own, unmodified
  ``calamity/calibration.py`` in the build container.  ``tests/golden/make_golden.py``
  uses it to generate the committed golden fixtures from the reference's own code.

Parity pin status (also in DESIGN.md): the reference ships no golden loss / gradient /
gain vectors and TensorFlow itself is not installable here, so TensorFlow's autodiff and
Keras optimizer ARITHMETIC are restated from their published update rules ("parity
unpinned" for those two pieces).  Everything that lives in the reference's own source --
chunking, tensor layout, index construction, the model/loss graph, the fit-loop
semantics (warm-up step, pre-update loss, use_min, tol stop) -- is pinned by executing the
reference's own code under ``tf_shim`` and comparing against the committed fixtures.
"""
