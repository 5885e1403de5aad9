"""TF-1 Adam over the DMT parameter store (tf.train.AdamOptimizer, inference_mlp.py:272-273).

Placeholder until the fused Adam kernel (K10) lands: constructing it is allowed so that
`Inference.get_optimizer` keeps the reference's surface, stepping raises.
"""


class TFAdam(object):
    def __init__(self, model, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.model, self.learning_rate = model, learning_rate
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self.t = 0

    def step(self, grads, lr=None):
        raise NotImplementedError("dmt_adam_* kernels are not built yet")
