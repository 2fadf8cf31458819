"""Entry points of nets/mj_loss.py kept callable for signature compatibility.  They are legacy pair /
verification losses that the 3-modality path never compiles into a model (only VerifLossLayer is used,
by the legacy UWYHNet.build); they evaluate eagerly on torch tensors."""
import torch

HUBER_DELTA = 0.5


def _t(x):
    return x if torch.is_tensor(x) else torch.as_tensor(x, dtype=torch.float32)


def mj_l2normalize(x, axis=1):
    x = _t(x)
    return x * torch.rsqrt(torch.clamp((x * x).sum(axis, keepdim=True), min=1e-12))


def mj_smoothL1(y_true, y_pred):
    x = (_t(y_true) - _t(y_pred)).abs()
    return torch.where(x < HUBER_DELTA, 0.5 * x ** 2, HUBER_DELTA * (x - 0.5 * HUBER_DELTA)).sum()


def mj_smoothL1bis(trash, y):
    return mj_smoothL1(y[0], y[1])


class _LossLayer:
    def __init__(self, alpha=0.5, **kwargs):
        self.alpha = alpha
        self.losses = []

    def add_loss(self, v):
        self.losses.append(v)

    def __call__(self, inputs):
        return self.call(inputs)

    def get_config(self):
        return {"alpha": self.alpha}


class PairLossLayer(_LossLayer):
    def pair_loss(self, inputs):
        x = (_t(inputs[0]) - _t(inputs[1])).abs()
        return torch.where(x < self.alpha, 0.5 * x ** 2, self.alpha * (x - 0.5 * self.alpha)).sum()

    def call(self, inputs):
        loss = self.pair_loss(inputs)
        self.add_loss(loss)
        return loss


class VerifLossLayer(_LossLayer):
    def pair_loss(self, inputs):
        a, b, labels = _t(inputs[0]), _t(inputs[1]), _t(inputs[2]).reshape(-1)
        res2 = (a - b) ** 2
        xpos = 0.5 * res2[labels == 1].sum()
        xneg = 0.5 * torch.clamp(self.alpha - torch.sqrt(res2[labels == 0].sum()), min=0.0) ** 2
        return xpos + xneg

    def call(self, inputs):
        loss = self.pair_loss(inputs)
        self.add_loss(loss)
        return loss


class TripletLossLayer(_LossLayer):
    def __init__(self, alpha, **kwargs):
        super().__init__(alpha, **kwargs)

    def triplet_loss(self, inputs):
        a, p, n = (_t(t) for t in inputs)
        return torch.clamp(((a - p) ** 2).sum(-1) - ((a - n) ** 2).sum(-1) + self.alpha, min=0).sum(0)

    def call(self, inputs):
        loss = self.triplet_loss(inputs)
        self.add_loss(loss)
        return loss
