"""Sparse optimizer descriptors for the fused backward+update kernel (K5).

Semantics follow TF-1.15 (the reference's optimizer layer, wrapped by
hybridbackend/tensorflow/training/optimizer.py:60-185): sharded embedding
gradients are applied locally without averaging (training/gradient.py:216-217).
"""
from hybridbackend_b200 import _lib


class SparseOptimizer:
  kind = 'sgd'
  num_slots = 0

  def __init__(self, learning_rate, fast_math=False):
    self.learning_rate = float(learning_rate)
    self.step = 0
    # fast_math: approximate sqrt/divide of the GPU (<= 2 ulp each, the arithmetic
    # class of TF's GPU kernels) instead of the IEEE sequence of TF's CPU kernels
    self.fast_math = bool(fast_math)

  @property
  def flags(self):
    return _lib.OPT_FLAG_FAST_MATH if self.fast_math else 0

  def slot_init(self, slot):  # pylint: disable=unused-argument
    return 0.0

  def descriptor(self):
    return _lib.hbOptimizer(_lib.OPT[self.kind], self.learning_rate, 0.0, 0.0, 0.0,
                            self.flags, max(self.step, 1))


class SGD(SparseOptimizer):
  """tf.train.GradientDescentOptimizer sparse apply: w -= lr * g."""


class Adagrad(SparseOptimizer):
  """tf.train.AdagradOptimizer sparse apply (SparseApplyAdagrad):
  accum += g*g; w -= lr * g / sqrt(accum); accum starts at 0.1."""
  kind = 'adagrad'
  num_slots = 1

  def __init__(self, learning_rate, initial_accumulator_value=0.1, fast_math=False):
    super().__init__(learning_rate, fast_math)
    self.initial_accumulator_value = float(initial_accumulator_value)

  def slot_init(self, slot):
    return self.initial_accumulator_value


class LazyAdam(SparseOptimizer):
  """tf.contrib.opt.LazyAdamOptimizer: m/v of touched rows only.  (TF-1.15's
  AdamOptimizer._apply_sparse decays every row each step -- dense work that is
  meaningless at 1e9 rows; BASELINE config 5's "sparse Adam" is LazyAdam.)"""
  kind = 'lazy_adam'
  num_slots = 2

  def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, fast_math=False):
    super().__init__(learning_rate, fast_math)
    self.beta1, self.beta2, self.epsilon = float(beta1), float(beta2), float(epsilon)

  def descriptor(self):
    return _lib.hbOptimizer(_lib.OPT[self.kind], self.learning_rate, self.beta1,
                            self.beta2, self.epsilon, self.flags, max(self.step, 1))
