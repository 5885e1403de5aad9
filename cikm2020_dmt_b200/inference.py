"""Drop-in for `DMT_code/model/inference_mlp.py::Inference` -- the reference's plugin boundary.

Same construction and call protocol (inference_mlp.py:18-68,117-118,170-171,260-280):

    inf = Inference(wnd_conf)                       # picks the class named by [model] model_type
    logits = inf.inference(features, is_train=...)  # ((click, order), y_bias) | (click, order)
    loss = inf.loss_multi_task_unbias(logits, labels, mask, is_train=..., loss_unbias_method=...,
                                      loss_ctr_rel_method=...)
    opt = inf.get_optimizer(name, lr)

Error convention kept from the reference: an unknown model or optimizer prints a message and
exits with status 1 (inference_mlp.py:66-68,278-280).
"""
import importlib
import sys

from . import keys as K


class Inference(object):
    def __init__(self, wnd_conf, **model_kwargs):
        self.wnd_conf = wnd_conf
        self.model_type = wnd_conf[K.MODEL][K.MODEL_TYPE]
        try:
            # same mechanism as inference_mlp.py:25: module `net.<model_type>`, class of that name
            self.module = importlib.import_module(".net.%s" % self.model_type, __package__)
            cls = getattr(self.module, self.model_type)
        except (ImportError, AttributeError):
            print("Unknown model, exit now")
            sys.exit(1)
        self.model = cls(wnd_conf, **model_kwargs)

    def inference(self, inputs, is_train=True, is_predict=False):
        return self.model.inference(inputs, is_train, is_predict)

    def loss_multi_task_unbias(self, logits, labels, mask, is_train=True, loss_unbias_method="two_head_add",
                               loss_ctr_rel_method="ctr"):
        # `labels` is unused by the reference too (inference_mlp.py:173-223 reads only `mask`)
        return self.model.loss(logits, mask, loss_unbias_method=loss_unbias_method,
                               loss_ctr_rel_method=loss_ctr_rel_method)

    def l2_norm(self, inputs):
        return self.model.l2_norm(inputs)

    def embedding_update(self, sess=None):
        return self.model.embedding_update(sess)

    def get_optimizer(self, optimizer, learning_rate):
        print("Use the optimizer: {}".format(optimizer))
        if optimizer == "adam":
            from .optim import TFAdam   # tf.train.AdamOptimizer semantics (inference_mlp.py:272-273)
            return TFAdam(self.model, learning_rate)
        if optimizer == "sgd":
            from .optim import TFGradientDescent     # inference_mlp.py:266-267
            return TFGradientDescent(self.model, learning_rate)
        if optimizer == "adagrad":
            from .optim import TFAdagrad             # inference_mlp.py:270-271
            return TFAdagrad(self.model, learning_rate)
        if optimizer in ("adadelta", "ftrl", "rmsprop"):
            raise NotImplementedError("optimizer %r is accepted by the reference (inference_mlp.py:264-277) "
                                      "but only 'adam' (dmt.conf), 'sgd' and 'adagrad' are built" % optimizer)
        print("Unknow optimizer, exit now")
        sys.exit(1)


DMTInference = Inference
