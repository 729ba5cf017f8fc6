"""CPU oracle for the recommend-tf2.0 hot path — TEST INFRASTRUCTURE ONLY.

numpy restatement of the reference's embedding lookup / pooling / feature-interaction layer
code (littlemesie/recommend-tf2.0, paths cited per function as src/...:line) and of the
TensorFlow/Keras op semantics those lines rely on.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package, and only as the checker or the reported CPU baseline.  The product
package (recommend-tf2.0_b200/) never imports it and has no CPU path.

PARITY PINNING: the reference ships no tests, golden vectors or fixtures, and TensorFlow is
not installable in the build image, so the TF kernels themselves could not be run.  The
restatement is pinned instead against the reference's OWN layer source executed over a numpy
stand-in for the handful of tf.* ops it calls (tools/tf_shim, tools/make_golden.py ->
tests/golden/*.npz).  TF's kernels are restated, not executed: "parity unpinned" in the
strict sense of the task statement, stated here and in DESIGN.md.
"""
