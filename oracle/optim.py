"""Oracle (test infrastructure, NOT product code): CPU restatement of the
reference's update rule, /root/reference/models/AcousticModel.py:388,404-406:

    clipped, _ = tf.clip_by_global_norm(accumulated_gradients, grad_clip)
    tf.train.AdamOptimizer(learning_rate_var).apply_gradients(...)

TensorFlow is absent (PARITY UNPINNED upstream); restated from TF 1.x:
  clip_by_global_norm: g * clip / max(||g||_2, clip), the norm over ALL tensors;
  ApplyAdam (beta1 .9, beta2 .999, eps 1e-8), step t = 1, 2, ...:
      lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
      m += (g - m) * (1 - beta1);  v += (g*g - v) * (1 - beta2)
      theta -= lr_t * m / (sqrt(v) + eps)          (eps OUTSIDE the corrected sqrt)
"""
import numpy as np


def global_norm(grads):
    return float(np.sqrt(sum(float(np.sum(np.asarray(g, np.float64) ** 2)) for g in grads)))


def clip_by_global_norm(flat_grad, clip):
    g = np.asarray(flat_grad, np.float64)
    norm = float(np.sqrt(np.sum(g * g)))
    return g * (clip / max(norm, clip)), norm


def adam_step(theta, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """All arrays float64, updated copies returned.  step is the 1-based count
    of this update."""
    lr_t = lr * np.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    m = m + (g - m) * (1.0 - beta1)
    v = v + (g * g - v) * (1.0 - beta2)
    theta = theta - lr_t * m / (np.sqrt(v) + eps)
    return theta, m, v


def clip_adam_step(theta, grad, m, v, step, lr, clip, beta1=0.9, beta2=0.999, eps=1e-8):
    g, norm = clip_by_global_norm(grad, clip)
    theta, m, v = adam_step(np.asarray(theta, np.float64), g, np.asarray(m, np.float64),
                            np.asarray(v, np.float64), step, lr, beta1, beta2, eps)
    return theta, m, v, norm
