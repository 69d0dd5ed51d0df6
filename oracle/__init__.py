"""CPU oracle for the ELEKTRONN2 volumetric-CNN hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``elektronn2_b200/`` may import this
package: it is the checker for the CUDA path, used by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py``.

What it is: a numpy float64 restatement of the arithmetic that the reference
(``/root/reference/elektronn2/neuromancer/computations.py`` + ``neural.py``)
asks Theano to do for Conv / UpConv / Pool / MFP / FragmentsToDense / Crop /
Concat, the loss nodes every shipped config uses, the Adam update, the shape
algebra and the ``predict_dense`` tiling.  Every function cites the reference
``file:line`` it follows.

PARITY PINNING (read this): the arithmetic itself lives in Theano
(``theano>=0.8,<0.10``, setup.py:53), which is not vendored under
/root/reference and cannot be imported here (python 3.12 / numpy 2).  The
reference's own tests hold no golden tensors for this path --
**numeric parity is therefore unpinned by the reference**.  What *is* pinned:

* semantics by known answers the reference does hold: ``np.convolve`` equality
  (tests/test_conv.py:89-104), shape / cost / parameter-count printouts
  (docs/examples.rst:100-140, 201-218; docs/predictions.rst:62-79) and the
  ``cnncalculator`` docstring example (utils/cnncalculator.py:241-260);
* the reference's importable pure-Python pieces (``TaggedShape``,
  ``cnncalculator``, ``initweights``), run in the build container by
  ``tests/golden/make_golden.py`` and committed as fixtures;
* two independent restatements per op (direct form vs the Theano-shaped
  decomposition: conv3d2d diagonal sum, pool_2d + z-maximum, unpool + conv)
  plus a torch-CPU float64 cross-check in the CPU test-suite.
"""

from . import ops, shapes, loss, adam, tiling, nets  # noqa: F401
